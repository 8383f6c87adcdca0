#!/bin/bash
# Round 2, trip 10 (1 GPU): three-piece split upload A/B (stock e2e), full GPU suite after the exchange / split changes,
# C++ mirror examples (SoA evaluateDSI, dsi_ assignment).
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/t10_smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -30 $O/t10_smoke.log; exit 1; }
( timeout 600 python tools/e2e_ab.py --steps 15 --variants "default=;pieces2_split15=EMVS_UPLOAD_PIECES:2,EMVS_UPLOAD_SPLIT:15;pieces3_split6=EMVS_UPLOAD_SPLIT:6;pieces3_split8=EMVS_UPLOAD_SPLIT:8;pieces3_split10=EMVS_UPLOAD_SPLIT:10;pieces3_split12=EMVS_UPLOAD_SPLIT:12;pieces3_split15=EMVS_UPLOAD_SPLIT:15;pieces3_split17=EMVS_UPLOAD_SPLIT:17" ) > $O/t10_e2e_ab.jsonl 2> $O/t10_e2e_ab.err
python - <<'PY'
import json
for ln in open("gpurun_out/r2/t10_e2e_ab.jsonl"):
    if ln.startswith("{"):
        d = json.loads(ln); print(d["variant"], "stock", d["stock_ms"], d["stock_calls_ms"], "| streaming", d["streaming_ms"])
PY
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $O/t10_pytest.log; tail -4 $O/t10_pytest.log
tail -n 3 $O/t10_e2e_ab.err
