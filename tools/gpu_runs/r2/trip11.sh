#!/bin/bash
# Round 2, trip 11 (1 GPU): L2 eviction hints (event-tile bulk copies, vote REDs, merged-plane stores, re-zero stores) x slab
# size: can the scratch be kept L2-resident between the re-zero and the next vote?  (the per-plane fixed cost of a vote
# launch equals ~16 MB of DRAM-speed traffic per 640x480 plane, profiles/r2_l2_hints.md)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/t11_smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -30 $O/t11_smoke.log; exit 1; }
V="default=;xy0_first=EMVS_HINT_XY0:1;xy0_last=EMVS_HINT_XY0:2;dsi_cs=EMVS_HINT_DSI:1;red_last=EMVS_HINT_RED:2"
V="$V;xy0f_dsi=EMVS_HINT_XY0:1,EMVS_HINT_DSI:1;xy0f_dsi_red=EMVS_HINT_XY0:1,EMVS_HINT_DSI:1,EMVS_HINT_RED:2"
V="$V;all_zero=EMVS_HINT_XY0:1,EMVS_HINT_DSI:1,EMVS_HINT_RED:2,EMVS_ZERO_CTAS:592,EMVS_HINT_ZERO:2"
V="$V;slab8=EMVS_SLAB:8;slab8_xy0f_dsi=EMVS_SLAB:8,EMVS_HINT_XY0:1,EMVS_HINT_DSI:1"
V="$V;slab8_xy0f_dsi_red=EMVS_SLAB:8,EMVS_HINT_XY0:1,EMVS_HINT_DSI:1,EMVS_HINT_RED:2"
V="$V;slab8_xy0l_dsi_red=EMVS_SLAB:8,EMVS_HINT_XY0:2,EMVS_HINT_DSI:1,EMVS_HINT_RED:2"
V="$V;slab8_all_zero=EMVS_SLAB:8,EMVS_HINT_XY0:1,EMVS_HINT_DSI:1,EMVS_HINT_RED:2,EMVS_ZERO_CTAS:592,EMVS_HINT_ZERO:2"
V="$V;default_b="
( timeout 600 python tools/ab_bench.py --variants "$V" ) > $O/t11_ab_hints.jsonl 2> $O/t11_ab_hints.err
python - <<'PY'
import json
for ln in open("gpurun_out/r2/t11_ab_hints.jsonl"):
    if ln.startswith("{"):
        d = json.loads(ln); print(f"{d['variant']:22s} {d['ms_per_step']:.3f} ms  {d['mevents_per_s']:.0f} Mev/s  vote {d['vote_ms_per_launch']:.4f} x {d['vote_launches_per_step']:.0f}  votes {d['accepted_votes']}")
PY
tail -n 3 $O/t11_ab_hints.err
