#!/bin/bash
# round 2, call 5: the pipelined per-layer kernel (WETTS_TC16P=1) at weight-ring depths 2/3/4: parity tests, bench, launch list
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
one() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('[$1] ms/step', round(d['ms_per_step'],2), 'gen', round(d['roofline']['ms'],2), 'value', round(d['value'],1), 'eager', d.get('gpu_eager_baseline'))"; }
timeout 240 python bench.py --steps 5 --warmup 3 2> gpurun_out/r2e_default.err | tee gpurun_out/r2e_default.json | one default
tail -3 gpurun_out/r2e_default.err
env WETTS_TC16P=1 timeout 400 python -m pytest tests/test_zz_widecases_gpu.py tests/test_mrf16_gpu.py tests/test_fused_gpu.py tests/test_golden_gpu.py tests/test_vits2_vocos_gpu.py -q -x -m gpu > gpurun_out/r2e_tc16p_tests.log 2>&1; echo "[tc16p] tests rc=$? $(tail -1 gpurun_out/r2e_tc16p_tests.log)"
for v in "WETTS_TC16P=1 WETTS_TC16P_NB=2" "WETTS_TC16P=1 WETTS_TC16P_NB=3" "WETTS_TC16P=1" "WETTS_TC16P=1 WETTS_MRF16_C128=1 WETTS_MRF16_C64_CTAS=2"; do
  env $v timeout 240 python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | one "$v"
done
env WETTS_TC16P=1 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2e_launches_tc16p.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu --profile-range > gpurun_out/r2e_ncu.log 2>&1; echo "ncu rc=$?"
python tools/launches.py gpurun_out/r2e_launches_tc16p.csv 2>&1 | head -24
for wl in baker_v1_gen_b64x640 aishell3_v1_b32x512 baker_v1_cli_b1; do
  env WETTS_TC16P=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --workload $wl 2>/dev/null | one "tc16p $wl"
done
