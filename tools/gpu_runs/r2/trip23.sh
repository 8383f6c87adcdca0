#!/bin/bash
# Round 2, trip 23 (1 GPU, last of the round's budget): split upload whose early pieces only vote (EMVS_UPLOAD_DEFER_MERGE):
# correctness on the small case (launch counts say which path ran) and at full size against the oracle, then the e2e A/B.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out/r2
O=gpurun_out/r2
( timeout 40 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k multi_piece_split 2>&1 | tail -4 ) > $O/t23_pytest_pieces.log; tail -2 $O/t23_pytest_pieces.log
V="default=;defer_p2_s15=EMVS_UPLOAD_DEFER_MERGE:1;defer_p3_s8=EMVS_UPLOAD_DEFER_MERGE:1,EMVS_UPLOAD_PIECES:3,EMVS_UPLOAD_SPLIT:8;defer_p4_s6=EMVS_UPLOAD_DEFER_MERGE:1,EMVS_UPLOAD_PIECES:4,EMVS_UPLOAD_SPLIT:6;defer_p4_s4=EMVS_UPLOAD_DEFER_MERGE:1,EMVS_UPLOAD_PIECES:4,EMVS_UPLOAD_SPLIT:4"
( timeout 30 python tools/e2e_ab.py --steps 10 --variants "$V" ) > $O/t23_e2e_ab.jsonl 2> $O/t23_e2e_ab.err
python - <<'PY'
import json
for ln in open("gpurun_out/r2/t23_e2e_ab.jsonl"):
    if ln.startswith("{"):
        d = json.loads(ln); print(f"e2e  {d['variant']:14s} stock {d['stock_ms']:.3f}  streaming {d['streaming_ms']:.3f}")
PY
( EMVS_UPLOAD_DEFER_MERGE=1 EMVS_UPLOAD_PIECES=4 EMVS_UPLOAD_SPLIT=6 timeout 40 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -k configs1 2>&1 | tail -4 ) > $O/t23_pytest_fullsize.log; tail -2 $O/t23_pytest_fullsize.log
tail -n 2 $O/t23_e2e_ab.err
