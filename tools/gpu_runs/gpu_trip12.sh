#!/bin/bash
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tests/mgpu_check.py > gpurun_out/mgpu8.log 2>&1; echo "mgpu8 exit $?" >> gpurun_out/mgpu8.log
for n in 8 4 2; do timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_n${n}_peer.json 2> gpurun_out/bench_n${n}_peer.err; echo "bench$n exit $?" >> gpurun_out/bench_n${n}_peer.err; done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 10 --warmup 3 --exchange nccl > gpurun_out/bench_n8_nccl.json 2> gpurun_out/bench_n8_nccl.err
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1_onbox8.json 2> gpurun_out/bench_n1_onbox8.err
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu8.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu8.log
