#!/bin/bash
# trip 23 (2 GPUs): full test suite incl. the 2-rank sharded build, N=2 bench with both exchanges, N=1 re-check
set -x
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_2gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_2gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2_peer.json 2> gpurun_out/bench_n2_peer.err; echo "exit $?" >> gpurun_out/bench_n2_peer.err
timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 --exchange nccl > gpurun_out/bench_n2_nccl.json 2> gpurun_out/bench_n2_nccl.err; echo "exit $?" >> gpurun_out/bench_n2_nccl.err
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1_b.json 2> gpurun_out/bench_n1_b.err
