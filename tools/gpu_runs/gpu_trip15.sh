#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tests/mgpu_check.py > gpurun_out/mgpu8.log 2>&1; echo "mgpu8 exit $?" >> gpurun_out/mgpu8.log
for n in 8 4; do timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_n${n}_band.json 2> gpurun_out/bench_n${n}_band.err; echo "bench$n exit $?" >> gpurun_out/bench_n${n}_band.err; done
