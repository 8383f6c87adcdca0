#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
for z in 0 1 2 4; do EMVS_FC_ZGROUPS=$z timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_fc$z.json 2>> gpurun_out/bench_fc.err; done
timeout 1500 python tools/sweep.py > gpurun_out/sweep.md 2> gpurun_out/sweep.err; echo "sweep exit $?" >> gpurun_out/sweep.err
