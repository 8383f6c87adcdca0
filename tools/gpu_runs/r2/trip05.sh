#!/bin/bash
# Round 2, trip 5 (2 GPUs): the adopted defaults (32-register persistent vote kernel at 6 CTAs per SM, peer reduce on 1024
# CTAs, float4 fuse+collapse) — multi-GPU tests, weak and strong scaling at N = 2 with e2e and parity, N = 1 beside it.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out/r2
O=gpurun_out/r2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15 ) > $O/t05_pytest.log
tail -3 $O/t05_pytest.log
( timeout 400 python bench.py --steps 10 --warmup 3 ) > $O/t05_n1.json 2> $O/t05_n1.err
( timeout 400 $TR --master-port 29641 bench.py --gpus 2 --steps 10 --warmup 3 ) > $O/t05_n2_weak.json 2> $O/t05_n2_weak.err
( EMVS_VOTE_CTAS_PER_SM=7 timeout 400 $TR --master-port 29642 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-parity ) > $O/t05_n2_weak_v7.json 2> $O/t05_n2_weak_v7.err
( EMVS_VOTE_CTAS_PER_SM=5 timeout 400 $TR --master-port 29643 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-parity ) > $O/t05_n2_weak_v5.json 2> $O/t05_n2_weak_v5.err
( timeout 600 $TR --master-port 29644 bench.py --gpus 2 --steps 8 --warmup 3 --scaling strong --events-per-cam 20000000 ) > $O/t05_n2_strong20M.json 2> $O/t05_n2_strong20M.err
for f in $O/t05_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e=lambda k: (round(d[k]["value"],1), round(d[k]["ms_per_step"],2)) if d.get(k) else None
    print(round(d["value"],1), round(d["ms_per_step"],3), "build", round(d["build_ms"],3), "depth", round(d["depth_map_ms"],3), "vote/launch", round(d["vote_ms_per_launch_max_over_ranks"],4), "e2e", e("e2e"), e("e2e_streaming"), e("e2e_soa"), "parity", (d.get("parity") or {}).get("ok"), "roof", round((d.get("roofline") or {}).get("frac") or 0,3))
except Exception as ex:
    print("unreadable:", ex)
PY
done
