#!/bin/bash
# round 2, call 1: per-block errors of the wide fixtures on both conv routes + the tx128 diagnostic
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python tools/diag.py 1 aishell3_long baker_v1_cli > gpurun_out/r2_diag_tc1.log 2>&1; echo "diag tc1 rc=$?"
timeout 300 python tools/diag.py 0 aishell3_long baker_v1_cli > gpurun_out/r2_diag_tc0.log 2>&1; echo "diag tc0 rc=$?"
timeout 300 python -m pytest tests/test_zz_widecases_gpu.py -q -s -rA -m gpu > gpurun_out/r2_wide.log 2>&1; echo "wide rc=$?"
cat gpurun_out/r2_diag_tc1.log gpurun_out/r2_diag_tc0.log
grep -E "text encoder|passed|failed|xfail|XPASS" gpurun_out/r2_wide.log
timeout 120 tools/ubench/mma_numerics > gpurun_out/r2_mma_numerics.txt 2>&1; echo "numerics rc=$?"
cat gpurun_out/r2_mma_numerics.txt
