#!/bin/bash
# Round 2, trip 20 (1 GPU): with one vote launch per build the slab count no longer costs launches: slab size and CTA-count A/B,
# then BASELINE configs[2] (bar4) with the adopted default.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out/r2
O=gpurun_out/r2
V="default=;slab8=EMVS_SLAB:8;slab24=EMVS_SLAB:24;slab32=EMVS_SLAB:32;ctas8=EMVS_VOTE_CTAS_PER_SM:8;ctas6=EMVS_VOTE_CTAS_PER_SM:6;default_b="
( timeout 130 python tools/ab_bench.py --variants "$V" ) > $O/t20_ab.jsonl 2> $O/t20_ab.err
python - <<'PY'
import json
for ln in open("gpurun_out/r2/t20_ab.jsonl"):
    if ln.startswith("{"):
        d = json.loads(ln); print(f"dev  {d['variant']:14s} {d['ms_per_step']:.3f} ms  vote {d['vote_ms_per_launch']:.4f} x {d['vote_launches_per_step']:.0f}  votes {d['accepted_votes']}")
PY
( timeout 80 python bench.py --steps 3 --warmup 3 --workload bar4 --no-cpu-baseline ) > $O/t20_bench_bar4.json 2> $O/t20_bench_bar4.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2/t20_bench_bar4.json").read().strip().splitlines()[-1])
    e=lambda k: (round(d[k]["value"],1), round(d[k].get("ms_per_step",0),2)) if d.get(k) else None
    print("bar4", round(d["value"],2), round(d["ms_per_step"],3), "e2e", e("e2e"), e("e2e_streaming"), e("e2e_soa"), "parity", (d.get("parity") or {}).get("ok"), "roof", (d.get("roofline") or {}).get("frac"))
except Exception as ex:
    print("unreadable:", ex)
PY
tail -n 2 $O/t20_ab.err $O/t20_bench_bar4.err
