#!/bin/bash
# trip 20 (1 GPU): plane-group size x slab sweep of the smem-staged k_vote_grouped<G>, parity for the new group sizes
set -x
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 120 $B > gpurun_out/sweep_$name.json 2>> gpurun_out/sweep.err
}
run g2_s12 EMVS_VOTE_GROUP=2
run g4_s12 EMVS_VOTE_GROUP=4
run g4_s8 EMVS_VOTE_GROUP=4 EMVS_SLAB=8
run g4_s16 EMVS_VOTE_GROUP=4 EMVS_SLAB=16
run g8_s8 EMVS_VOTE_GROUP=8 EMVS_SLAB=8
run g8_s16 EMVS_VOTE_GROUP=8 EMVS_SLAB=16
run g8_s16_nov EMVS_VOTE_GROUP=8 EMVS_SLAB=16 EMVS_OVERLAP=0
run g16_s16 EMVS_VOTE_GROUP=16 EMVS_SLAB=16
run g16_s16_nov EMVS_VOTE_GROUP=16 EMVS_SLAB=16 EMVS_OVERLAP=0
run g32_s32_nov EMVS_VOTE_GROUP=32 EMVS_SLAB=32 EMVS_OVERLAP=0
EMVS_VOTE_GROUP=4 timeout 120 $B --kind uniform > gpurun_out/sweep_g4_s12_uniform.json 2>> gpurun_out/sweep.err
EMVS_VOTE_GROUP=8 EMVS_SLAB=8 timeout 120 $B --kind uniform > gpurun_out/sweep_g8_s8_uniform.json 2>> gpurun_out/sweep.err
for g in 8 16 32; do
  EMVS_VOTE_GROUP=$g timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -x -q > gpurun_out/pytest_gpu_g$g.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_g$g.log
done
