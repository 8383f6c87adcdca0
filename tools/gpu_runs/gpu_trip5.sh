#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 120 dvs_mcemvs_b200/host/example_process1 > gpurun_out/host_example.log 2>&1; echo "example exit $?" >> gpurun_out/host_example.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?" >> gpurun_out/bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench2 exit $?" >> gpurun_out/bench_n2.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_r1b.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_vote -s 40 -c 2 -o gpurun_out/prof_vote_r1b \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_fuse_collapse|k_merge" -s 70 -c 3 -o gpurun_out/prof_other_r1b \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_other.log 2>&1
