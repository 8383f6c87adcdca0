#!/bin/bash
# Round 2, trip 15 (2 GPUs): the north-star case on one GPU (10 M events x 2 cameras, 640x480x256) with its full-size parity
# block, and configs[4] at 2 GPUs again with the final kernels (the trip-2 column predates the reduce-grid sweep).
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out/r2
O=gpurun_out/r2
( CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --steps 10 --warmup 3 --events-per-cam 10000000 --cpu-sample-events 10000000 ) > $O/t15_bench_n1_10M.json 2> $O/t15_bench_n1_10M.err
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29951 tools/sweep.py --no-cpu ) > $O/t15_sweep_n2.md 2> $O/t15_sweep_n2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2/t15_bench_n1_10M.json").read().strip().splitlines()[-1])
e=lambda k: (round(d[k]["value"],1), round(d[k]["ms_per_step"],2))
print(round(d["value"],1), round(d["ms_per_step"],3), "e2e", e("e2e"), e("e2e_streaming"), e("e2e_soa"), "parity", d["parity"]["ok"], d["parity"]["counts_exact"], d["parity"]["conf_max_rel"], d["parity"]["idx_mismatches"], "roof", d["roofline"]["frac"], d["roofline"]["hbm_algorithmic"]["frac"], "cpu", d["cpu_baseline"]["value"])
PY
grep "^|" $O/t15_sweep_n2.md
