#!/bin/bash
# Round 2, trip 13 (1 GPU): what is the per-slab fixed cost of a vote launch?  One scratch buffer, nothing else running
# (EMVS_OVERLAP=0 + no merge + no re-zero: WRONG volumes, timing only) at slab sizes 4 (G=4) / 8 / 16 / 32, 5 M and 10 M events.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out/r2
O=gpurun_out/r2
S="EMVS_OVERLAP:0,EMVS_DEBUG_SKIP_MERGE:1,EMVS_DEBUG_SKIP_ZERO:1"
V="default=;one_buf_s16=$S;one_buf_s8=$S,EMVS_SLAB:8;one_buf_s32=$S,EMVS_SLAB:32;one_buf_s4_g4=$S,EMVS_SLAB:4,EMVS_VOTE_GROUP:4;two_buf_s4_g4=EMVS_SLAB:4,EMVS_VOTE_GROUP:4;two_buf_s8_g4=EMVS_SLAB:8,EMVS_VOTE_GROUP:4;one_buf_s8_g4=$S,EMVS_SLAB:8,EMVS_VOTE_GROUP:4;one_buf_s16_nohint=$S,EMVS_HINT_RED:0,EMVS_HINT_XY0:0"
for n in 5000000 10000000; do
  ( timeout 600 python tools/ab_bench.py --events-per-cam $n --steps 6 --variants "$V" ) > $O/t13_ab_$n.jsonl 2> $O/t13_ab_$n.err
done
python - <<'PY'
import json
for n in (5000000, 10000000):
    for ln in open(f"gpurun_out/r2/t13_ab_{n}.jsonl"):
        if ln.startswith("{"):
            d = json.loads(ln); print(f"{n//1000000:3d}M {d['variant']:20s} {d['ms_per_step']:.3f} ms  vote {d['vote_ms_per_launch']:.4f} x {d['vote_launches_per_step']:.0f}")
PY
