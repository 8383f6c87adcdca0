#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
for v in 0 1; do
  EMVS_VOTE_PAIR=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_pair$v.json 2>> gpurun_out/bench_pair.err
  EMVS_VOTE_PAIR=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --kind uniform > gpurun_out/bench_pair${v}_uniform.json 2>> gpurun_out/bench_pair.err
done
