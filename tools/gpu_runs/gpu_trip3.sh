#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 tools/red_locality_bench 64 > gpurun_out/red_locality.csv 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_r1.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_vote -s 100 -c 2 -o gpurun_out/prof_vote_r1 \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_bench.log 2>&1
ls -la gpurun_out
