#!/bin/bash
# round 2, call 19: row-block-resident kernel (WETTS_TC16R=1) for the multi-tile layers: parity tests, bench, launch list
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
one() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('[$1] ms/step', round(d['ms_per_step'],2), 'gen', round(d['roofline']['ms'],2), 'value', round(d['value'],1))"; }
env WETTS_TC16R=1 timeout 900 python -m pytest tests/test_zz_widecases_gpu.py tests/test_mrf16_gpu.py tests/test_fused_gpu.py tests/test_parity_gpu.py tests/test_vits2_vocos_gpu.py tests/test_fullsize_gpu.py -q -x -m gpu > gpurun_out/r2u_tests.log 2>&1; echo "[tc16r] tests rc=$? $(tail -1 gpurun_out/r2u_tests.log)"
for wl in multilingual_v3_b256x128 baker_v1_gen_b64x640 aishell3_v1_b32x512; do
  for v in "X=1" "WETTS_TC16R=1" "WETTS_TC16R=1 WETTS_TC16R_MIN_TILES=3"; do
    env $v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --workload $wl 2>/dev/null | one "$wl $v"
  done
done
env WETTS_TC16R=1 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2u_launches.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu --profile-range > gpurun_out/r2u_ncu.log 2>&1; echo "ncu rc=$?"
python tools/launches.py gpurun_out/r2u_launches.csv 2>&1 | head -22
