#!/bin/bash
# Round 2, trip 19 (1 GPU): final numbers of the adopted configuration (multi-slab vote launch): bench (configs[1]) with
# parity and CPU baseline, the 20 M strong-scaling baseline, ncu launch list, one full capture of the vote kernel.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out/r2
O=gpurun_out/r2
( timeout 200 python bench.py --steps 10 --warmup 3 ) > $O/t19_bench_n1.json 2> $O/t19_bench_n1.err
( timeout 100 python bench.py --steps 5 --warmup 3 --events-per-cam 20000000 --scaling strong --no-cpu-baseline --no-parity ) > $O/t19_bench_n1_20M.json 2> $O/t19_bench_n1_20M.err
( timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 330 --csv --log-file $O/t19_launches.csv \
    python bench.py --steps 2 --warmup 2 --no-e2e --no-cpu-baseline --no-parity ) > $O/t19_bench_under_ncu.log 2>&1
( timeout 150 ncu --set full --clock-control none --import-source on -k regex:"k_vote_tma" -s 4 -c 1 -o $O/t19_vote \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-parity ) > $O/t19_ncu_vote.log 2>&1
for f in $O/t19_bench_n1.json $O/t19_bench_n1_20M.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e=lambda k: (round(d[k]["value"],1), round(d[k].get("ms_per_step",0),2)) if d.get(k) else None
    print(round(d["value"],2), round(d["ms_per_step"],3), "e2e", e("e2e"), e("e2e_streaming"), e("e2e_soa"), "parity", (d.get("parity") or {}).get("ok"), "roof", (d.get("roofline") or {}).get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"), "launches", d.get("gpu_launches"))
except Exception as ex:
    print("unreadable:", ex)
PY
done
tail -n 2 $O/t19_bench_n1.err $O/t19_ncu_vote.log
