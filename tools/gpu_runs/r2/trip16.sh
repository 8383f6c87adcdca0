#!/bin/bash
# Round 2, trip 16 (1 GPU): one vote launch per build (EMVS_VOTE_MULTISLAB: all slabs in one persistent launch, merges started
# from per-slab completion counters) against one launch per slab: correctness of the variants, device-resident and e2e A/B.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/t16_smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -30 $O/t16_smoke.log; exit 1; }
( EMVS_VOTE_MULTISLAB=1 timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/t16_smoke_ms.log 2>&1 || { echo "MULTISLAB SMOKE FAILED"; tail -30 $O/t16_smoke_ms.log; exit 1; }
tail -1 $O/t16_smoke_ms.log
( timeout 600 python -m pytest tests/test_gpu_soa_kernels.py -m gpu -x -q 2>&1 | tail -8 ) > $O/t16_pytest_variants.log; tail -3 $O/t16_pytest_variants.log
V="default=;multislab=EMVS_VOTE_MULTISLAB:1;default_b=;multislab_b=EMVS_VOTE_MULTISLAB:1;multislab_vs0=EMVS_VOTE_MULTISLAB:1,EMVS_VOTE_SPLIT:0"
( timeout 600 python tools/ab_bench.py --variants "$V" ) > $O/t16_ab.jsonl 2> $O/t16_ab.err
( timeout 600 python tools/e2e_ab.py --steps 15 --variants "default=;multislab=EMVS_VOTE_MULTISLAB:1;default_b=;multislab_b=EMVS_VOTE_MULTISLAB:1" ) > $O/t16_e2e_ab.jsonl 2> $O/t16_e2e_ab.err
python - <<'PY'
import json
for ln in open("gpurun_out/r2/t16_ab.jsonl"):
    if ln.startswith("{"):
        d = json.loads(ln); print(f"dev  {d['variant']:14s} {d['ms_per_step']:.3f} ms  vote {d['vote_ms_per_launch']:.4f} x {d['vote_launches_per_step']:.0f}  votes {d['accepted_votes']}")
for ln in open("gpurun_out/r2/t16_e2e_ab.jsonl"):
    if ln.startswith("{"):
        d = json.loads(ln); print(f"e2e  {d['variant']:14s} stock {d['stock_ms']:.3f}  streaming {d['streaming_ms']:.3f}")
PY
tail -n 3 $O/t16_ab.err $O/t16_e2e_ab.err
