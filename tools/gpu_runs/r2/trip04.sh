#!/bin/bash
# Round 2, trip 4 (2 GPUs): grid of the slab-wise peer reduce x resident vote CTAs, at N = 2 (weak scaling), with the
# N = 1 figure of the same box beside it.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out/r2
O=gpurun_out/r2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
B="--steps 10 --warmup 3 --no-e2e --no-parity --no-cpu-baseline"
( timeout 300 python bench.py $B ) > $O/t04_n1.json 2> $O/t04_n1.err
( EMVS_VOTE_REGS32=1 timeout 300 python bench.py $B ) > $O/t04_n1_regs32.json 2> $O/t04_n1_regs32.err
i=0
for v in "0 6" "256 6" "1024 6" "4096 6" "1024 5" "1024 7"; do
  set -- $v; i=$((i+1))
  ( EMVS_PEER_REDUCE_CTAS=$1 EMVS_VOTE_CTAS_PER_SM=$2 timeout 300 $TR --master-port $((29610+i)) bench.py --gpus 2 $B ) > $O/t04_n2_r$1_v$2.json 2> $O/t04_n2_r$1_v$2.err
done
( EMVS_PEER_REDUCE_CTAS=1024 timeout 300 $TR --master-port 29630 bench.py --gpus 2 $B --exchange nccl ) > $O/t04_n2_nccl.json 2> $O/t04_n2_nccl.err
for f in $O/t04_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d["value"],1), round(d["ms_per_step"],3), "build", round(d["build_ms"],3), "depth", round(d["depth_map_ms"],3), "vote/launch", round(d["vote_ms_per_launch_max_over_ranks"],4))
except Exception as ex:
    print("unreadable:", ex)
PY
done
