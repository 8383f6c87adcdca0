#!/bin/bash
# Round 2, trip 9 (N GPUs, default 2): camera x sub-interval sharding (one camera per GPU group) against sub-interval-only
# sharding — multi-GPU tests, weak and strong scaling with e2e and parity.
N=${1:-2}
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out/r2
O=gpurun_out/r2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" = 2 ]; then
  ( timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15 ) > $O/t09_pytest.log; tail -3 $O/t09_pytest.log
else
  ( timeout 600 $TR --master-port 29801 tests/mgpu_check.py ) > $O/t09_mgpu_check_n$N.log 2>&1; tail -6 $O/t09_mgpu_check_n$N.log
fi
( timeout 400 $TR --master-port 29802 bench.py --gpus $N --steps 10 --warmup 3 ) > $O/t09_n${N}_weak_2d.json 2> $O/t09_n${N}_weak_2d.err
( timeout 400 $TR --master-port 29803 bench.py --gpus $N --steps 10 --warmup 3 --sharding interval --no-e2e --no-parity ) > $O/t09_n${N}_weak_interval.json 2> $O/t09_n${N}_weak_interval.err
( timeout 600 $TR --master-port 29804 bench.py --gpus $N --steps 10 --warmup 3 --scaling strong --events-per-cam 20000000 ) > $O/t09_n${N}_strong20M_2d.json 2> $O/t09_n${N}_strong20M_2d.err
for f in $O/t09_n${N}_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e=lambda k: (round(d[k]["value"],1), round(d[k]["ms_per_step"],2)) if d.get(k) else None
    print(round(d["value"],1), round(d["ms_per_step"],3), "build", round(d["build_ms"],3), "depth", round(d["depth_map_ms"],3), "vote/launch", round(d["vote_ms_per_launch_max_over_ranks"],4), "e2e", e("e2e"), e("e2e_streaming"), e("e2e_soa"), "parity", (d.get("parity") or {}).get("ok"))
except Exception as ex:
    print("unreadable:", ex)
PY
done
tail -n 5 $O/t09_n${N}_weak_2d.err | grep -v "OMP_NUM\|\*\*\*\*"
