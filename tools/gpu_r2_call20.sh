#!/bin/bash
# round 2, call 20: tc16r + C=64 two-CTA fused stage as defaults: full GPU suite, bench, clock64 profile of tc16r
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
one() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('[$1] ms/step', round(d['ms_per_step'],2), 'gen', round(d['roofline']['ms'],2), 'value', round(d['value'],1))"; }
timeout 1200 python -m pytest tests -q -x -m gpu > gpurun_out/r2v_tests.log 2>&1; echo "tests rc=$? $(tail -1 gpurun_out/r2v_tests.log)"
for v in "X=1" "WETTS_TC16R=0" "WETTS_MRF16_C64_CTAS=1"; do
  env $v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | one "$v"
done
env WETTS_TC16R_PROFILE=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu --batch 64 > /dev/null 2> gpurun_out/r2v_tc16r_profile.txt
grep "tc16r profile" gpurun_out/r2v_tc16r_profile.txt | tail -64 | awk '!seen[$0]++' | cut -c17- | sort | uniq -c | sort -rn | head -24
