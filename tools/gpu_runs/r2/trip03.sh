#!/bin/bash
# Round 2, trip 3 (1 GPU): end-to-end A/B (stock vs streaming; classic vs TMA kernel; upload split; work-item split),
# device-resident A/B of the work-item split and the fuse+collapse variants, validation of the kernel variants.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out/r2
O=gpurun_out/r2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/t03_smoke.log 2>&1 || { echo "SMOKE FAILED"; tail -30 $O/t03_smoke.log; exit 1; }
( timeout 900 python -m pytest tests/test_gpu_soa_kernels.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -15 ) > $O/t03_pytest.log
tail -3 $O/t03_pytest.log
( timeout 600 python tools/e2e_ab.py ) > $O/t03_e2e_ab.jsonl 2> $O/t03_e2e_ab.err
( timeout 600 python tools/ab_bench.py --variants "default=;vsplit0=EMVS_VOTE_SPLIT:0;vsplit1=EMVS_VOTE_SPLIT:1;vsplit2=EMVS_VOTE_SPLIT:2;cta6=EMVS_VOTE_CTAS_PER_SM:6;cta6_vs1=EMVS_VOTE_CTAS_PER_SM:6,EMVS_VOTE_SPLIT:1;fc_v4_z8=EMVS_FC_V4:1,EMVS_FC_ZSPLIT:8;fc_v4_z16=EMVS_FC_V4:1,EMVS_FC_ZSPLIT:16;fc_z16=EMVS_FC_ZSPLIT:16;classic=EMVS_VOTE_KERNEL:classic" ) > $O/t03_ab.jsonl 2> $O/t03_ab.err
( timeout 600 python tools/ab_bench.py --events-per-cam 1250000 --variants "head_default=;head_vsplit0=EMVS_VOTE_SPLIT:0;head_vsplit2=EMVS_VOTE_SPLIT:2;head_classic=EMVS_VOTE_KERNEL:classic" ) > $O/t03_ab_head.jsonl 2> $O/t03_ab_head.err
cat $O/t03_e2e_ab.jsonl $O/t03_ab.jsonl $O/t03_ab_head.jsonl
tail -n 3 $O/t03_e2e_ab.err $O/t03_ab.err
