#!/bin/bash
# trip 22 (1 GPU): final configuration (group 8, 2 x 16 planes, grouped merge, prefetch): full tests, full bench,
# e2e A/B of the prefetch, ncu launch list + full capture of the vote and merge kernels
set -x
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?" >> gpurun_out/bench_n1.err
timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-prefetch > gpurun_out/bench_noprefetch.json 2>> gpurun_out/bench_ab.err
EMVS_UPLOAD_SPLIT=0 timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-prefetch > gpurun_out/bench_noprefetch_nosplit.json 2>> gpurun_out/bench_ab.err
EMVS_HOST_THREADS=1 timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_prefetch_1thread.json 2>> gpurun_out/bench_ab.err
timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --kind uniform > gpurun_out/bench_uniform.json 2>> gpurun_out/bench_ab.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_vote_grouped -s 40 -c 2 -o gpurun_out/vote_g8 -f python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_vote.log 2>&1
ncu -i gpurun_out/vote_g8.ncu-rep --page raw --csv > gpurun_out/vote_g8_raw.csv 2>/dev/null
timeout 200 ncu --set full --clock-control none -k regex:k_merge_quads_grouped -s 40 -c 1 -o gpurun_out/merge_g8 -f python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_merge.log 2>&1
ncu -i gpurun_out/merge_g8.ncu-rep --page raw --csv > gpurun_out/merge_g8_raw.csv 2>/dev/null
