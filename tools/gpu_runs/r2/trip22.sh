#!/bin/bash
# Round 2, trip 22 (1 GPU): the single-launch build on the large DSIs of configs[4] (scratch for all slabs: 4.3 / 8.6 GB) against
# the per-slab fallback they use under the default 4 GB budget: identical counts, equal volumes, build time.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out/r2
( timeout 105 python tools/multislab_large_check.py --events 5000000 --deadline-s 55 ) > gpurun_out/r2/t22_large.jsonl 2> gpurun_out/r2/t22_large.err
cat gpurun_out/r2/t22_large.jsonl; tail -n 3 gpurun_out/r2/t22_large.err
