#!/bin/bash
# round 2, call 21: 256-sample work items of the fused MRF stage kernel (opt-in): parity on hardware, same-box A/B, phase profile
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
one() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('[$1] ms/step', round(d['ms_per_step'],2), 'gen', round(d['roofline']['ms'],2), 'value', round(d['value'],1))"; }
T="tests/test_mrf16_gpu.py tests/test_fused_gpu.py tests/test_parity_gpu.py tests/test_zz_widecases_gpu.py"
env WETTS_MRF16_ITEM_C32=256 WETTS_MRF16_ITEM_C64=256 WETTS_MRF16_ITEM_RB1=256 timeout 600 python -m pytest $T -q -x -m gpu > gpurun_out/r2w_tests_item256.log 2>&1; echo "tests item256 rc=$? $(tail -1 gpurun_out/r2w_tests_item256.log)"
for rep in 1 2; do
for v in "X=1" "WETTS_MRF16_ITEM_C32=256" "WETTS_MRF16_ITEM_C32=256 WETTS_MRF16_ITEM_C64=256" "WETTS_MRF16_ITEM_C64=256"; do
  env $v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | one "rep$rep $v"
done
done
for v in "X=1" "WETTS_MRF16_ITEM_RB1=256"; do
  env $v timeout 300 python bench.py --workload baker_v1_gen_b64x640 --steps 3 --warmup 3 --no-cpu 2>/dev/null | one "v1gen $v"
done
env WETTS_MRF16_ITEM_C32=256 WETTS_MRF16_ITEM_C64=256 WETTS_FUSED_RB_PROFILE=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu --batch 64 > /dev/null 2> gpurun_out/r2w_mrf16_item256_profile.txt
grep -A2 "fused_mrf16 profile" gpurun_out/r2w_mrf16_item256_profile.txt | tail -6
env WETTS_FUSED_RB_PROFILE=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu --batch 64 > /dev/null 2> gpurun_out/r2w_mrf16_item128_profile.txt
grep -A2 "fused_mrf16 profile" gpurun_out/r2w_mrf16_item128_profile.txt | tail -6
