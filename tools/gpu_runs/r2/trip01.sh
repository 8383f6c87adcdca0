#!/bin/bash
# Round 2, trip 1 (1 GPU): first run of the TMA-staged persistent vote kernel, the SoA entry points, the new bench
# (parity block), the design micro-benchmarks and the merge / re-zero interference A/B.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out/r2
O=gpurun_out/r2
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/t01_smi.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/t01_smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -30 $O/t01_smoke.log; exit 1; }
tail -2 $O/t01_smoke.log
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > $O/t01_pytest.log
( timeout 120 tools/vote_variants_bench csv ) > $O/t01_variants.csv 2> $O/t01_variants.err
( timeout 600 python tools/ab_bench.py ) > $O/t01_ab.jsonl 2> $O/t01_ab.err
( timeout 600 python bench.py --steps 10 --warmup 3 ) > $O/t01_bench.json 2> $O/t01_bench.err
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 160 --csv --log-file $O/t01_launches.csv \
    python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu-baseline --no-parity ) > $O/t01_bench_under_ncu.log 2>&1
( timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_vote_tma -s 40 -c 2 -o $O/t01_vote_tma \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-parity ) > $O/t01_ncu_vote.log 2>&1
tail -5 $O/t01_pytest.log
cat $O/t01_variants.csv
cat $O/t01_ab.jsonl
cut -c1-1500 $O/t01_bench.json
