#!/bin/bash
# Round 2, trip 18 (2 GPUs): the multi-slab vote launch (now the default) under the multi-GPU paths — the 2-rank tests
# (sharded build, Alg-2, bench parity weak / strong / interval), then the weak and strong bench lines at N = 2.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out/r2
O=gpurun_out/r2
( timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -8 ) > $O/t18_pytest_multi.log; tail -2 $O/t18_pytest_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
( timeout 300 $TR --master-port 29961 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline ) > $O/t18_bench_n2.json 2> $O/t18_bench_n2.err
( timeout 300 $TR --master-port 29962 bench.py --gpus 2 --steps 5 --warmup 3 --events-per-cam 20000000 --scaling strong --no-cpu-baseline ) > $O/t18_bench_n2_strong.json 2> $O/t18_bench_n2_strong.err
for f in $O/t18_bench_n2.json $O/t18_bench_n2_strong.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e=lambda k: (round(d[k]["value"],1), round(d[k].get("ms_per_step",0),2)) if d.get(k) else None
    print(round(d["value"],2), round(d["ms_per_step"],3), "e2e", e("e2e"), e("e2e_streaming"), e("e2e_soa"), "parity", (d.get("parity") or {}).get("ok"), "roof", (d.get("roofline") or {}).get("frac"))
except Exception as ex:
    print("unreadable:", ex)
PY
done
tail -n 3 $O/t18_bench_n2.err $O/t18_bench_n2_strong.err
