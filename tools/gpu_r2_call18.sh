#!/bin/bash
# round 2, call 18: same-box A/B of the opt-in fused-stage variants (two repeats each), memcheck of the smoke case
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
one() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('[$1] ms/step', round(d['ms_per_step'],2), 'gen', round(d['roofline']['ms'],2), 'value', round(d['value'],1))"; }
for rep in 1 2; do
  for v in "X=1" "WETTS_MRF16_C64_CTAS=2" "WETTS_MRF16_C128=1" "WETTS_MRF16_C128=1 WETTS_MRF16_C64_CTAS=2" "WETTS_MRF16_CTAS=2"; do
    env $v timeout 200 python bench.py --steps 8 --warmup 3 --no-cpu 2>/dev/null | one "rep$rep $v"
  done
done
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python __graft_entry__.py smoke > gpurun_out/r2t_memcheck_smoke.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r2t_memcheck_smoke.log
