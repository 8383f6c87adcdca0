#!/bin/bash
set -x
mkdir -p gpurun_out
for g in 1 2 4; do
  EMVS_VOTE_GROUP=$g timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_g$g.json 2>> gpurun_out/bench_g.err
done
EMVS_VOTE_GROUP=4 timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_g4.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_g4.log
