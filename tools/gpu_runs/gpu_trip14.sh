#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/mgpu_check.py > gpurun_out/mgpu2.log 2>&1; echo "mgpu2 exit $?" >> gpurun_out/mgpu2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2_band.json 2> gpurun_out/bench_n2_band.err; echo "bench2 exit $?" >> gpurun_out/bench_n2_band.err
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
