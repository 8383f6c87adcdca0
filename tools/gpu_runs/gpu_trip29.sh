#!/bin/bash
# trip 29 (1 GPU): final validation after the slab chooser change (full group of 8 planes for large planes)
set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 100 python tools/sweep.py --sizes 1024x1024x256,1024x1024x512,640x480x256 --counts 10000000 --no-cpu > gpurun_out/sweep_after.md 2>> gpurun_out/sweep_after.err
timeout 100 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_final.json 2> gpurun_out/bench_n1_final.err
