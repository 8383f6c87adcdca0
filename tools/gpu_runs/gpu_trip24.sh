#!/bin/bash
# trip 24 (1 GPU): prepared division — device self-test against __fdiv_rn, parity with it enabled, A/B
set -x
mkdir -p gpurun_out
timeout 120 python -c "
from dvs_mcemvs_b200 import api
c = api.Context(0)
import time
for seed in range(1, 9):
    t = time.time(); bad = c.selftest_division(1 << 32, seed); print('seed', seed, 'pairs 2^32 mismatches', bad, round(time.time() - t, 2), 's', flush=True)
" > gpurun_out/selftest_division.log 2>&1
EMVS_VOTE_FASTDIV=1 timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_fastdiv.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_fastdiv.log
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
for f in 0 1 0 1; do
  EMVS_VOTE_FASTDIV=$f timeout 120 $B >> gpurun_out/bench_fastdiv$f.json 2>> gpurun_out/bench_fastdiv.err
done
EMVS_VOTE_FASTDIV=1 timeout 120 $B --kind uniform > gpurun_out/bench_fastdiv1_uniform.json 2>> gpurun_out/bench_fastdiv.err
EMVS_VOTE_FASTDIV=1 timeout 200 ncu --set full --clock-control none -k regex:k_vote_grouped -s 40 -c 1 -o gpurun_out/vote_g8_fastdiv -f $B --steps 2 --warmup 1 > gpurun_out/ncu_vote_fastdiv.log 2>&1
ncu -i gpurun_out/vote_g8_fastdiv.ncu-rep --page raw --csv > gpurun_out/vote_g8_fastdiv_raw.csv 2>/dev/null
