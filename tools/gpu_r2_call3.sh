#!/bin/bash
# round 2, call 3: every BASELINE.json config as a bench workload on one GPU (with the CPU / GPU-eager legs), smoke
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke.log
for wl in multilingual_v3_b256x128 baker_v1_cli_b1 baker_v1_gen_b64x640 baker_v3_gen_b64x640 aishell3_v1_b32x512 multilingual_v3_b1024x128 baker_v1_b64x128; do
  timeout 420 python bench.py --steps 5 --warmup 3 --workload $wl > gpurun_out/r2c_$wl.json 2> gpurun_out/r2c_$wl.err; rc=$?
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2c_$wl.json"))
    print("$wl rc=$rc: ms/step", round(d["ms_per_step"],3), "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "gen ms", round(d["roofline"]["ms"],2), "frac", round(d["roofline"]["frac"],3),
          "| cpu", d["cpu_baseline"] and (round(d["cpu_baseline"]["value"],2), d["cpu_baseline"]["kind"], d["cpu_baseline"]["cores"]), "| eager", d.get("gpu_eager_baseline"), "| dur", d.get("duration_check"))
except Exception as e:
    print("$wl rc=$rc failed", e); print(open("gpurun_out/r2c_$wl.err").read()[-1200:])
PY
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c_reference_arm.json 2> gpurun_out/r2c_reference_arm.err; echo "reference arm rc=$?"; head -c 600 gpurun_out/r2c_reference_arm.json
# experiments (opt-in kernel variants): 4-deep activation ring of the per-layer kernel, C = 128 stage in the fused kernel
for v in "WETTS_TC16_ABUF=4" "WETTS_MRF16_C128=1" "WETTS_TC16P=1" "WETTS_TC16P=1 WETTS_MRF16_C128=1"; do
  env $v timeout 300 python -m pytest tests/test_zz_widecases_gpu.py tests/test_mrf16_gpu.py tests/test_fused_gpu.py -q -x -m gpu > gpurun_out/r2c_exp_tests.log 2>&1; echo "[$v] tests rc=$? $(tail -1 gpurun_out/r2c_exp_tests.log)"
  env $v timeout 240 python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('[$v] ms/step', round(d['ms_per_step'],2), 'gen', round(d['roofline']['ms'],2))"
done
