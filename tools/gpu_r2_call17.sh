#!/bin/bash
# round 2, call 17: batch-composition invariance diagnostic; staging branch + two-CTA policy: tests and bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
one() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('[$1] ms/step', round(d['ms_per_step'],2), 'gen', round(d['roofline']['ms'],2), 'value', round(d['value'],1))"; }
timeout 300 python tools/diag_batch_invariance.py > gpurun_out/r2r_diag.log 2>&1; echo "diag rc=$?"; cat gpurun_out/r2r_diag.log | tail -60
timeout 900 python -m pytest tests -q -x -m gpu > gpurun_out/r2r_tests.log 2>&1; echo "tests rc=$? $(tail -1 gpurun_out/r2r_tests.log)"
for wl in multilingual_v3_b256x128 baker_v1_gen_b64x640 aishell3_v1_b32x512; do
  for v in "X=1" "WETTS_TC16_TWO_CTAS=0"; do
    env $v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --workload $wl 2>/dev/null | one "$wl $v"
  done
done
