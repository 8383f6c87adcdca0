#!/bin/bash
# 8-GPU validation: multi-rank parity check at 8 ranks, bench at N=8 and N=4.
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus8.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tests/mgpu_check.py > gpurun_out/mgpu8.log 2>&1; echo "mgpu8 exit $?" >> gpurun_out/mgpu8.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench8 exit $?" >> gpurun_out/bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo "bench4 exit $?" >> gpurun_out/bench_n4.err
