#!/bin/bash
# Round 2, trip 21 (1 GPU): upload split re-tuned for the single-launch builds (a piece no longer costs 16 vote launches), then
# a sanity pass over the final tree (smoke, parity + variant tests).
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out/r2
O=gpurun_out/r2
V="default=;split10=EMVS_UPLOAD_SPLIT:10;split25=EMVS_UPLOAD_SPLIT:25;split35=EMVS_UPLOAD_SPLIT:35;p3_split10=EMVS_UPLOAD_PIECES:3,EMVS_UPLOAD_SPLIT:10;p3_split15=EMVS_UPLOAD_PIECES:3,EMVS_UPLOAD_SPLIT:15;default_b="
( timeout 100 python tools/e2e_ab.py --steps 12 --variants "$V" ) > $O/t21_e2e_ab.jsonl 2> $O/t21_e2e_ab.err
python - <<'PY'
import json
for ln in open("gpurun_out/r2/t21_e2e_ab.jsonl"):
    if ln.startswith("{"):
        d = json.loads(ln); print(f"e2e  {d['variant']:14s} stock {d['stock_ms']:.3f}  streaming {d['streaming_ms']:.3f}")
PY
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( timeout 100 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3 ) > $O/t21_pytest.log; tail -2 $O/t21_pytest.log
tail -n 2 $O/t21_e2e_ab.err
