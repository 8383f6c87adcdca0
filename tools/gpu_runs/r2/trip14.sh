#!/bin/bash
# Round 2, trip 14 (1 GPU): the adopted configuration — full GPU suite, bench (configs[1]) with parity, reference arm,
# strong-scaling N = 1 baseline (20 M events per camera), configs[2] (bar4) with parity, ncu launch list + full captures (L2 hints adopted).

cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out/r2
O=gpurun_out/r2
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $O/t14_pytest.log; tail -3 $O/t14_pytest.log
( timeout 600 python bench.py --steps 10 --warmup 3 ) > $O/t14_bench_n1.json 2> $O/t14_bench_n1.err
( timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > $O/t14_bench_ref.json 2> $O/t14_bench_ref.err
( timeout 600 python bench.py --steps 5 --warmup 3 --events-per-cam 20000000 --scaling strong --no-cpu-baseline --no-parity ) > $O/t14_bench_n1_20M.json 2> $O/t14_bench_n1_20M.err
( timeout 900 python bench.py --steps 3 --warmup 3 --workload bar4 --cpu-sample-events 10000000 ) > $O/t14_bench_bar4.json 2> $O/t14_bench_bar4.err
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 170 --csv --log-file $O/t14_launches.csv \
    python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu-baseline --no-parity ) > $O/t14_bench_under_ncu.log 2>&1
( timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_vote_tma|k_merge_quads_grouped|k_fuse_collapse_zsplit_v4" -s 60 -c 5 -o $O/t14_kernels \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-parity ) > $O/t14_ncu_kernels.log 2>&1
for f in $O/t14_bench_n1.json $O/t14_bench_ref.json $O/t14_bench_n1_20M.json $O/t14_bench_bar4.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e=lambda k: (round(d[k]["value"],1), round(d[k].get("ms_per_step",0),2)) if d.get(k) else None
    print(round(d["value"],2), round(d["ms_per_step"],3), "e2e", e("e2e"), e("e2e_streaming"), e("e2e_soa"), "parity", (d.get("parity") or {}).get("ok"), "roof", (d.get("roofline") or {}).get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
except Exception as ex:
    print("unreadable:", ex)
PY
done
