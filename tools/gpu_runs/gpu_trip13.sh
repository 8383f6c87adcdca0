#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
for z in 1 4 8 16; do EMVS_PEER_ZSPLIT=$z timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$z bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_n2_pz$z.json 2>> gpurun_out/bench_n2_pz.err; done
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/racecheck_smoke.log 2>&1; echo "racecheck exit $?" >> gpurun_out/racecheck_smoke.log
