#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 120 dvs_mcemvs_b200/host/example_process1 > gpurun_out/host_example.log 2>&1; echo "example exit $?" >> gpurun_out/host_example.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?" >> gpurun_out/bench_n1.err
timeout 900 python bench.py --steps 10 --warmup 3 --events-per-cam 10000000 --no-cpu-baseline > gpurun_out/bench_n1_10M.json 2> gpurun_out/bench_n1_10M.err
