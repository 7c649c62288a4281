#!/bin/bash
# round 2, call 22: 384-sample items (C=32), L2 prefetch in the unprefetched C=64 variant: parity + same-box A/B
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
one() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('[$1] ms/step', round(d['ms_per_step'],2), 'gen', round(d['roofline']['ms'],2), 'value', round(d['value'],1))"; }
T="tests/test_mrf16_gpu.py tests/test_fused_gpu.py tests/test_parity_gpu.py tests/test_zz_widecases_gpu.py"
env WETTS_MRF16_ITEM_C32=384 WETTS_MRF16_L2PF=1 timeout 600 python -m pytest $T -q -x -m gpu > gpurun_out/r2x_tests_item384.log 2>&1; echo "tests item384+l2pf rc=$? $(tail -1 gpurun_out/r2x_tests_item384.log)"
for rep in 1 2; do
for v in "X=1" "WETTS_MRF16_ITEM_C32=256" "WETTS_MRF16_ITEM_C32=384" "WETTS_MRF16_L2PF=1" "WETTS_MRF16_ITEM_C32=256 WETTS_MRF16_L2PF=1"; do
  env $v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | one "rep$rep $v"
done
done
for v in "WETTS_TC16_TWO_CTAS=1" "WETTS_TC16_TWO_CTAS=1 WETTS_TC16R_MIN_TILES=1"; do
  env $v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | one "$v"
done
env WETTS_MRF16_ITEM_C32=384 WETTS_MRF16_L2PF=1 WETTS_FUSED_RB_PROFILE=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu --batch 64 > /dev/null 2> gpurun_out/r2x_mrf16_item384_profile.txt
grep -A2 "fused_mrf16 profile" gpurun_out/r2x_mrf16_item384_profile.txt | tail -6
