#!/bin/bash
# trip 26 (4 GPUs): weak scaling of the final kernel with the slab-wise peer reduce, 4-rank sharded-build check
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541"
timeout 200 $TR tests/mgpu_check.py > gpurun_out/mgpu_check_n4.log 2>&1; echo "exit $?" >> gpurun_out/mgpu_check_n4.log
timeout 200 $TR bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/bench_n4_peer.json 2> gpurun_out/bench_n4_peer.err; echo "exit $?" >> gpurun_out/bench_n4_peer.err
