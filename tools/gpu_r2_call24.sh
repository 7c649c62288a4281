#!/bin/bash
# round 2, call 24: batched staging loads in the C=64 fused stage variant (opt-in), new work-item-size test
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
one() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('[$1] ms/step', round(d['ms_per_step'],2), 'gen', round(d['roofline']['ms'],2), 'value', round(d['value'],1))"; }
timeout 600 python -m pytest tests/test_mrf16_gpu.py -q -x -m gpu > gpurun_out/r2z_tests_mrf16.log 2>&1; echo "tests mrf16 rc=$? $(tail -1 gpurun_out/r2z_tests_mrf16.log)"
env WETTS_MRF16_STAGE_BATCH=1 timeout 600 python -m pytest tests/test_mrf16_gpu.py tests/test_fused_gpu.py tests/test_parity_gpu.py -q -x -m gpu > gpurun_out/r2z_tests_sb.log 2>&1; echo "tests stage_batch rc=$? $(tail -1 gpurun_out/r2z_tests_sb.log)"
for rep in 1 2; do
for v in "X=1" "WETTS_MRF16_STAGE_BATCH=1"; do
  env $v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | one "rep$rep $v"
done
done
env WETTS_MRF16_STAGE_BATCH=1 WETTS_FUSED_RB_PROFILE=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu --batch 64 > /dev/null 2> gpurun_out/r2z_profile.txt
grep -A2 "fused_mrf16 profile" gpurun_out/r2z_profile.txt | tail -6
