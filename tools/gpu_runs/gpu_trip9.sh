#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/memcheck_smoke.log 2>&1; echo "memcheck exit $?" >> gpurun_out/memcheck_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?" >> gpurun_out/bench_n1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_r1c.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_vote -s 44 -c 2 -o gpurun_out/prof_vote_r1c \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_bench.log 2>&1
