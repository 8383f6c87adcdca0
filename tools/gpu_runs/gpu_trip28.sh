#!/bin/bash
# trip 28 (1 GPU): plane group / slab for DSIs whose planes are larger than the 80 MiB budget allows 8 of (1024^2)
set -x
mkdir -p gpurun_out
S="python tools/sweep.py --sizes 1024x1024x256 --counts 10000000 --no-cpu"
for v in "EMVS_VOTE_GROUP=4 EMVS_SLAB=4" "EMVS_VOTE_GROUP=8 EMVS_SLAB=8" "EMVS_VOTE_GROUP=4 EMVS_SLAB=8" "EMVS_VOTE_GROUP=2 EMVS_SLAB=4" "EMVS_VOTE_GROUP=8 EMVS_SLAB=8 EMVS_OVERLAP=0" "EMVS_VOTE_GROUP=4 EMVS_SLAB=4 EMVS_OVERLAP=0"; do
  echo "## $v" >> gpurun_out/sweep_1024.md
  env $v timeout 60 $S 2>> gpurun_out/sweep_1024.err | grep "^| 1024" >> gpurun_out/sweep_1024.md
done
