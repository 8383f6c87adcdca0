#!/bin/bash
# Round 2, trip 24 (1 GPU, the round's last GPU seconds): the split-upload tests with the deferred merge as the default.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out/r2
( timeout 24 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "split" 2>&1 | tail -4 ) > gpurun_out/r2/t24_pytest_split.log; tail -2 gpurun_out/r2/t24_pytest_split.log
( timeout 9 python tools/e2e_ab.py --steps 6 --variants "default=" ) > gpurun_out/r2/t24_e2e.jsonl 2>/dev/null; cut -c1-200 gpurun_out/r2/t24_e2e.jsonl
