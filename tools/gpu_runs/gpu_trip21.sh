#!/bin/bash
# trip 21 (1 GPU): grouped merge kernel A/B, events-per-CTA (wave tail) A/B, parity for the new variants
set -x
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
run() { local name=$1; shift; env "$@" timeout 120 $B > gpurun_out/sweep2_$name.json 2>> gpurun_out/sweep2.err; }
run g8_s16_oldmerge EMVS_VOTE_GROUP=8 EMVS_SLAB=16 EMVS_MERGE_GROUPED=0
run g8_s16 EMVS_VOTE_GROUP=8 EMVS_SLAB=16
run g8_s8 EMVS_VOTE_GROUP=8 EMVS_SLAB=8
run g8_s16_nov EMVS_VOTE_GROUP=8 EMVS_SLAB=16 EMVS_OVERLAP=0
run g8_s16_e512 EMVS_VOTE_GROUP=8 EMVS_SLAB=16 EMVS_VOTE_EPC=512
run g8_s16_e256 EMVS_VOTE_GROUP=8 EMVS_SLAB=16 EMVS_VOTE_EPC=256
run g16_s16_e512 EMVS_VOTE_GROUP=16 EMVS_SLAB=16 EMVS_VOTE_EPC=512
run g4_s12 EMVS_VOTE_GROUP=4
run g4_s12_e512 EMVS_VOTE_GROUP=4 EMVS_VOTE_EPC=512
EMVS_VOTE_GROUP=8 EMVS_SLAB=16 timeout 120 $B --events-per-cam 10000000 > gpurun_out/sweep2_g8_s16_10M.json 2>> gpurun_out/sweep2.err
for e in 512 256; do
  EMVS_VOTE_GROUP=8 EMVS_VOTE_EPC=$e timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -x -q > gpurun_out/pytest_gpu_g8_e$e.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_g8_e$e.log
done
EMVS_VOTE_GROUP=2 timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -x -q > gpurun_out/pytest_gpu_g2_mg.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_g2_mg.log
EMVS_VOTE_GROUP=8 EMVS_SLAB=16 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_g8_s16.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu2.log 2>&1
