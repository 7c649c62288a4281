#!/bin/bash
# round 2, call 8: pipelined kernel with all worker warps staging + draining (WETTS_TC16P=1), fast gate epilogue
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
one() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('[$1] ms/step', round(d['ms_per_step'],2), 'gen', round(d['roofline']['ms'],2), 'value', round(d['value'],1))"; }
timeout 600 python -m pytest tests/test_zz_widecases_gpu.py tests/test_mrf16_gpu.py tests/test_fused_gpu.py tests/test_parity_gpu.py tests/test_vits2_vocos_gpu.py -q -x -m gpu > gpurun_out/r2h_tests_default.log 2>&1; echo "[default] tests rc=$? $(tail -1 gpurun_out/r2h_tests_default.log)"
env WETTS_TC16P=1 timeout 600 python -m pytest tests/test_zz_widecases_gpu.py tests/test_mrf16_gpu.py tests/test_fused_gpu.py tests/test_parity_gpu.py tests/test_vits2_vocos_gpu.py tests/test_fullsize_gpu.py -q -x -m gpu > gpurun_out/r2h_tests_tc16p.log 2>&1; echo "[tc16p] tests rc=$? $(tail -1 gpurun_out/r2h_tests_tc16p.log)"
for wl in multilingual_v3_b256x128 baker_v1_gen_b64x640 baker_v3_gen_b64x640 aishell3_v1_b32x512; do
  for v in "X=1" "WETTS_TC16P=1" "WETTS_TC16P=1 WETTS_TC16P_ALLWARPS=0"; do
    env $v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --workload $wl 2>/dev/null | one "$wl $v"
  done
done
env WETTS_TC16P=1 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2h_launches_tc16p_aw.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu --profile-range > gpurun_out/r2h_ncu.log 2>&1; echo "ncu rc=$?"
python tools/launches.py gpurun_out/r2h_launches_tc16p_aw.csv 2>&1 | head -8
