#!/bin/bash
# round 2, call 9: 64-wide N tiles (WETTS_TC16_NMAX=64) + two accumulator slots in TMEM: the pipelined kernel drains item i
# while the MMAs of item i + 1 run
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
one() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('[$1] ms/step', round(d['ms_per_step'],2), 'gen', round(d['roofline']['ms'],2), 'value', round(d['value'],1))"; }
env WETTS_TC16_NMAX=64 WETTS_TC16P=1 timeout 600 python -m pytest tests/test_zz_widecases_gpu.py tests/test_mrf16_gpu.py tests/test_fused_gpu.py tests/test_parity_gpu.py tests/test_vits2_vocos_gpu.py tests/test_fullsize_gpu.py -q -x -m gpu > gpurun_out/r2i_tests_pp.log 2>&1; echo "[nmax64 tc16p] tests rc=$? $(tail -1 gpurun_out/r2i_tests_pp.log)"
for wl in multilingual_v3_b256x128 baker_v1_gen_b64x640 baker_v3_gen_b64x640; do
  for v in "X=1" "WETTS_TC16_NMAX=64" "WETTS_TC16_NMAX=64 WETTS_TC16P=1" "WETTS_TC16_NMAX=64 WETTS_TC16P=1 WETTS_TC16P_ALLWARPS=1" "WETTS_TC16_NMAX=64 WETTS_TC16P=1 WETTS_TC16P_PINGPONG=0 WETTS_TC16P_ALLWARPS=0"; do
    env $v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --workload $wl 2>/dev/null | one "$wl $v"
  done
done
env WETTS_TC16_NMAX=64 WETTS_TC16P=1 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2i_launches_pp.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu --profile-range > gpurun_out/r2i_ncu.log 2>&1; echo "ncu rc=$?"
python tools/launches.py gpurun_out/r2i_launches_pp.csv 2>&1 | head -8
env WETTS_TC16_NMAX=64 WETTS_TC16P=1 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2i_launches_pp_v1gen.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu --profile-range --workload baker_v1_gen_b64x640 > gpurun_out/r2i_ncu2.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.DictReader(l for l in open("gpurun_out/r2i_launches_pp_v1gen.csv") if l.startswith('"'))]
tc=[round(float(r["Metric Value"].replace(",",""))/1e3) for r in rows if "tc16" in r["Kernel Name"]]
print("v1 gen, nmax64 + tc16p ping-pong: tc16 n=%d sum %.2f ms: %s" % (len(tc), sum(tc)/1e3, tc))
PY
