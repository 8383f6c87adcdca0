#!/bin/bash
# Round 2, trip 7 (N GPUs, N = $1, default 8): multi-GPU correctness at N ranks with the final kernels (mgpu_check,
# Alg-2 sharded), weak scaling with e2e + parity, BASELINE configs[3] (20 M events per camera split over the GPUs, strong
# scaling) with e2e + parity, configs[4] sweep at N GPUs.
N=${1:-8}
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out/r2
O=gpurun_out/r2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
( timeout 600 $TR --master-port 29701 tests/mgpu_check.py ) > $O/t07_mgpu_check_n$N.log 2>&1; tail -6 $O/t07_mgpu_check_n$N.log
( timeout 600 $TR --master-port 29702 tests/mgpu_check_alg2.py ) > $O/t07_mgpu_alg2_n$N.log 2>&1; tail -5 $O/t07_mgpu_alg2_n$N.log
( timeout 600 $TR --master-port 29703 bench.py --gpus $N --steps 10 --warmup 3 ) > $O/t07_n${N}_weak.json 2> $O/t07_n${N}_weak.err
( timeout 900 $TR --master-port 29704 bench.py --gpus $N --steps 10 --warmup 3 --scaling strong --events-per-cam 20000000 ) > $O/t07_n${N}_strong20M.json 2> $O/t07_n${N}_strong20M.err
( timeout 900 $TR --master-port 29705 tools/sweep.py --no-cpu ) > $O/t07_sweep_n$N.md 2> $O/t07_sweep_n$N.err
for f in $O/t07_n${N}_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e=lambda k: (round(d[k]["value"],1), round(d[k]["ms_per_step"],2)) if d.get(k) else None
    print(round(d["value"],1), round(d["ms_per_step"],3), "build", round(d["build_ms"],3), "depth", round(d["depth_map_ms"],3), "vote/launch", round(d["vote_ms_per_launch_max_over_ranks"],4), "e2e", e("e2e"), e("e2e_streaming"), e("e2e_soa"), "parity", (d.get("parity") or {}).get("ok"))
except Exception as ex:
    print("unreadable:", ex)
PY
done
cat $O/t07_sweep_n$N.md
tail -n 3 $O/t07_n${N}_weak.err $O/t07_n${N}_strong20M.err $O/t07_sweep_n$N.err
