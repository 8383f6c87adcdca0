#!/bin/bash
# Round 2, trip 6 (1 GPU): why is the streaming e2e step 0.3-0.9 ms above the device-resident step with the persistent
# vote kernel but only 0.2 ms with the classic one?  Alternating variants, 20 steps, per-call host times.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out/r2
O=gpurun_out/r2
V="default=;classic=EMVS_VOTE_KERNEL:classic;default_b=;classic_b=EMVS_VOTE_KERNEL:classic;cta7=EMVS_VOTE_CTAS_PER_SM:7;cta8=EMVS_VOTE_CTAS_PER_SM:8;cta8_b=EMVS_VOTE_CTAS_PER_SM:8"
( timeout 600 python tools/e2e_ab.py --steps 20 --variants "$V" ) > $O/t06_e2e_ab.jsonl 2> $O/t06_e2e_ab.err
cat $O/t06_e2e_ab.jsonl; tail -n 5 $O/t06_e2e_ab.err
