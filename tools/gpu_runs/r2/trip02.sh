#!/bin/bash
# Round 2, trip 2 (2 GPUs): multi-GPU correctness (mgpu_check incl. the persistent float4 peer reduce, Alg-2 sharded,
# bench parity weak + strong inside pytest), weak and strong scaling at N = 2, peer-reduce grid A/B, sweep at N = 2.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out/r2
O=gpurun_out/r2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( timeout 1200 python -m pytest tests/test_gpu_multi.py tests/test_gpu_soa_kernels.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -25 ) > $O/t02_pytest.log
tail -4 $O/t02_pytest.log
( timeout 600 $TR --master-port 29601 bench.py --gpus 2 --steps 10 --warmup 3 ) > $O/t02_bench_n2_weak.json 2> $O/t02_bench_n2_weak.err
for c in 0 16 256; do
  ( EMVS_PEER_REDUCE_CTAS=$c timeout 600 $TR --master-port 29602 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-parity ) > $O/t02_bench_n2_weak_ctas$c.json 2> $O/t02_bench_n2_weak_ctas$c.err
done
( timeout 600 $TR --master-port 29603 bench.py --gpus 2 --steps 10 --warmup 3 --exchange nccl --no-e2e ) > $O/t02_bench_n2_weak_nccl.json 2> $O/t02_bench_n2_weak_nccl.err
( timeout 900 $TR --master-port 29604 bench.py --gpus 2 --steps 8 --warmup 3 --scaling strong --events-per-cam 20000000 ) > $O/t02_bench_n2_strong20M.json 2> $O/t02_bench_n2_strong20M.err
( timeout 900 $TR --master-port 29605 tools/sweep.py --no-cpu ) > $O/t02_sweep_n2.md 2> $O/t02_sweep_n2.err
for f in $O/t02_bench_n2_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e=lambda k: (round(d[k]["value"],1), round(d[k]["ms_per_step"],2)) if d.get(k) else None
    print(round(d["value"],1), round(d["ms_per_step"],3), "build", round(d["build_ms"],3), "depth", round(d["depth_map_ms"],3), "vote/launch", round(d["vote_ms_per_launch_max_over_ranks"],4), "e2e", e("e2e"), e("e2e_streaming"), e("e2e_soa"), "parity", (d.get("parity") or {}).get("ok"))
except Exception as ex:
    print("unreadable:", ex)
PY
done
cat $O/t02_sweep_n2.md
tail -3 $O/*.err | tail -40
