#!/bin/bash
# round 2, call 6: N-tile-minor item order + L2 prefetch in the per-layer kernel: parity tests, bench variants, launch list,
# then ncu --set full captures of the dominant kernels (batch 64)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
one() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('[$1] ms/step', round(d['ms_per_step'],2), 'gen', round(d['roofline']['ms'],2), 'value', round(d['value'],1))"; }
timeout 600 python -m pytest tests/test_zz_widecases_gpu.py tests/test_mrf16_gpu.py tests/test_fused_gpu.py tests/test_parity_gpu.py tests/test_fullsize_gpu.py tests/test_vits2_vocos_gpu.py -q -x -m gpu > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$? $(tail -1 gpurun_out/r2f_tests.log)"
for v in "X=1" "WETTS_TC16_PREFETCH=0" "WETTS_TC16_NTMINOR=0" "WETTS_TC16_PREFETCH=0 WETTS_TC16_NTMINOR=0" "WETTS_MRF16_C128=1 WETTS_MRF16_C64_CTAS=2"; do
  env $v timeout 240 python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | one "$v"
done
timeout 500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2f_dram_traffic.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu --profile-range > gpurun_out/r2f_ncu.log 2>&1; echo "ncu rc=$?"
python tools/launches.py gpurun_out/r2f_dram_traffic.csv 2>&1 | head -12
for wl in baker_v1_gen_b64x640 aishell3_v1_b32x512; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --workload $wl 2>/dev/null | one "$wl"
done
prof() { # name, kernel regex (function base name), skip
  timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$2 -s $3 -c 1 -o gpurun_out/r2_$1 -f \
     python bench.py --steps 1 --warmup 2 --no-cpu --batch 64 --profile-range > gpurun_out/r2_ncu_$1.log 2>&1
  echo "$1 rc=$? $(tail -n 1 gpurun_out/r2_ncu_$1.log)"
}
prof mrf16_c64 "fused_mrf16_kernel" 0
prof mrf16_c32 "fused_mrf16_kernel" 1
prof tc16_flow_in "conv1d_tc16_kernel" 29
prof tc16_flow_rs "conv1d_tc16_kernel" 30
prof tc16_rb128 "conv1d_tc16_kernel" 70
prof tc16_up2 "conv1d_tc16_kernel" 77
ls -la gpurun_out/*.ncu-rep
