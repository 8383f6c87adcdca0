#!/bin/bash
# Round 2, trip 12 (2 GPUs): do the L2 hints (vote REDs evict_last, re-zero stores evict_last on 592 CTAs, event tiles evict_first,
# merged planes streamed) that gave -2 % on the device-resident step also hold end to end and at N = 2?
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out/r2
O=gpurun_out/r2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
H="all=EMVS_HINT_XY0:1,EMVS_HINT_DSI:1,EMVS_HINT_RED:2,EMVS_ZERO_CTAS:592,EMVS_HINT_ZERO:2"
V="default=;$H;red_only=EMVS_HINT_RED:2;default_b=;all_b=EMVS_HINT_XY0:1,EMVS_HINT_DSI:1,EMVS_HINT_RED:2,EMVS_ZERO_CTAS:592,EMVS_HINT_ZERO:2;red_zero=EMVS_HINT_RED:2,EMVS_ZERO_CTAS:592,EMVS_HINT_ZERO:2;red_zero296=EMVS_HINT_RED:2,EMVS_ZERO_CTAS:296,EMVS_HINT_ZERO:2"
( timeout 600 python tools/e2e_ab.py --steps 15 --variants "$V" ) > $O/t12_e2e_ab.jsonl 2> $O/t12_e2e_ab.err
( timeout 600 python tools/ab_bench.py --variants "$V" ) > $O/t12_ab.jsonl 2> $O/t12_ab.err
python - <<'PY'
import json
for ln in open("gpurun_out/r2/t12_e2e_ab.jsonl"):
    if ln.startswith("{"):
        d = json.loads(ln); print(f"e2e  {d['variant']:12s} stock {d['stock_ms']:.3f}  streaming {d['streaming_ms']:.3f}")
for ln in open("gpurun_out/r2/t12_ab.jsonl"):
    if ln.startswith("{"):
        d = json.loads(ln); print(f"dev  {d['variant']:12s} {d['ms_per_step']:.3f} ms  vote {d['vote_ms_per_launch']:.4f}")
PY
B="--gpus 2 --steps 10 --warmup 3 --no-e2e --no-parity"
( timeout 300 $TR --master-port 29901 bench.py $B ) > $O/t12_n2_default.json 2> $O/t12_n2_default.err
( EMVS_HINT_XY0=1 EMVS_HINT_DSI=1 EMVS_HINT_RED=2 EMVS_ZERO_CTAS=592 EMVS_HINT_ZERO=2 timeout 300 $TR --master-port 29902 bench.py $B ) > $O/t12_n2_hints.json 2> $O/t12_n2_hints.err
( EMVS_HINT_RED=2 timeout 300 $TR --master-port 29903 bench.py $B ) > $O/t12_n2_red.json 2> $O/t12_n2_red.err
for f in $O/t12_n2_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["ms_per_step"],3), "vote/launch", round(d["vote_ms_per_launch_max_over_ranks"],4), "depth", round(d["depth_map_ms"],3))
PY
done
