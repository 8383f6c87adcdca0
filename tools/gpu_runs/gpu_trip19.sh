#!/bin/bash
# trip 19 (1 GPU): full validation of the final configuration (k_vote_grouped<2>, split upload), e2e A/B of the split
# upload, plane-group A/B, ncu launch list of the bench command.
set -x
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?" >> gpurun_out/bench_n1.err
for s in 0 15 35 50; do
  EMVS_UPLOAD_SPLIT=$s timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_split$s.json 2>> gpurun_out/bench_split.err
done
for g in 1 4; do
  EMVS_VOTE_GROUP=$g timeout 120 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_g$g.json 2>> gpurun_out/bench_g.err
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
EMVS_VOTE_GROUP=4 timeout 240 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -x -q > gpurun_out/pytest_gpu_g4.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_g4.log
