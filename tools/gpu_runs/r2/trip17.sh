#!/bin/bash
# Round 2, trip 17 (1 GPU): vote loop moved into an out-of-line device function (66 instructions per vote again, no spills);
# per-slab launches against the single multi-slab launch, correctness of both (full GPU suite under each), A/B timing.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/t17_smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -30 $O/t17_smoke.log; exit 1; }
( EMVS_VOTE_MULTISLAB=1 timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/t17_smoke_ms.log 2>&1 || { echo "MULTISLAB SMOKE FAILED"; tail -30 $O/t17_smoke_ms.log; exit 1; }
tail -1 $O/t17_smoke_ms.log
V="default=;multislab=EMVS_VOTE_MULTISLAB:1;default_b=;multislab_b=EMVS_VOTE_MULTISLAB:1"
( timeout 400 python tools/ab_bench.py --variants "$V" ) > $O/t17_ab.jsonl 2> $O/t17_ab.err
( timeout 400 python tools/e2e_ab.py --steps 15 --variants "default=;multislab=EMVS_VOTE_MULTISLAB:1" ) > $O/t17_e2e_ab.jsonl 2> $O/t17_e2e_ab.err
python - <<'PY'
import json
for ln in open("gpurun_out/r2/t17_ab.jsonl"):
    if ln.startswith("{"):
        d = json.loads(ln); print(f"dev  {d['variant']:14s} {d['ms_per_step']:.3f} ms  vote {d['vote_ms_per_launch']:.4f} x {d['vote_launches_per_step']:.0f}  votes {d['accepted_votes']}")
for ln in open("gpurun_out/r2/t17_e2e_ab.jsonl"):
    if ln.startswith("{"):
        d = json.loads(ln); print(f"e2e  {d['variant']:14s} stock {d['stock_ms']:.3f}  streaming {d['streaming_ms']:.3f}")
PY
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > $O/t17_pytest.log; tail -2 $O/t17_pytest.log
( EMVS_VOTE_MULTISLAB=1 timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > $O/t17_pytest_ms.log; tail -2 $O/t17_pytest_ms.log
tail -n 3 $O/t17_ab.err $O/t17_e2e_ab.err
