#!/bin/bash
# First GPU trip: microbench, parity tests, smoke, bench, ncu launch list + full capture of k_vote.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 tools/red_microbench 64 > gpurun_out/red_microbench.csv 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; echo "bench exit $?" >> gpurun_out/bench_a.err
timeout 600 python bench.py --steps 5 --warmup 3 --kind uniform --no-cpu-baseline > gpurun_out/bench_uniform.json 2> gpurun_out/bench_uniform.err
for s in 4 8 16 32 64; do EMVS_SLAB=$s timeout 300 python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/bench_slab$s.json 2>> gpurun_out/bench_slab.err; done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --events-per-cam 2000000 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_vote -s 8 -c 2 -o gpurun_out/prof_vote \
    python bench.py --steps 1 --warmup 1 --events-per-cam 2000000 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_bench.log 2>&1
ls -la gpurun_out
