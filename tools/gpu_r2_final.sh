#!/bin/bash
# round 2, final 1-GPU call: smoke, full GPU suite, every BASELINE config as a bench line (with the CPU / GPU-eager legs),
# reference arm, launch list + DRAM traffic of one step
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/r2s_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2s_smoke.log
timeout 1200 python -m pytest tests -q -x -m gpu > gpurun_out/r2s_tests.log 2>&1; echo "tests rc=$? $(tail -1 gpurun_out/r2s_tests.log)"
for wl in multilingual_v3_b256x128 baker_v1_cli_b1 baker_v1_gen_b64x640 baker_v3_gen_b64x640 aishell3_v1_b32x512 multilingual_v3_b1024x128 baker_v1_b64x128; do
  timeout 420 python bench.py --steps 5 --warmup 3 --workload $wl > gpurun_out/r2s_$wl.json 2> gpurun_out/r2s_$wl.err; rc=$?
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2s_$wl.json"))
    print("$wl rc=$rc: ms/step", round(d["ms_per_step"],3), "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "gen ms", round(d["roofline"]["ms"],2), "frac", round(d["roofline"]["frac"],3),
          "| cpu", d["cpu_baseline"] and (round(d["cpu_baseline"]["value"],2), d["cpu_baseline"]["kind"], d["cpu_baseline"]["cores"]), "| eager", d.get("gpu_eager_baseline") and round(d["gpu_eager_baseline"].get("value",0),1), "| dur", d.get("duration_check"), "| clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print("$wl rc=$rc failed", e); print(open("gpurun_out/r2s_$wl.err").read()[-1200:])
PY
done
timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu --length-aware 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('length-aware: ms/step', round(d['ms_per_step'],2), 'value', round(d['value'],1))"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2s_reference_arm.json 2> gpurun_out/r2s_reference_arm.err; echo "reference arm rc=$?"; head -c 400 gpurun_out/r2s_reference_arm.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off \
    --csv --log-file gpurun_out/r2s_dram_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu --profile-range > gpurun_out/r2s_ncu.log 2>&1; echo "ncu rc=$?"
python tools/launches.py gpurun_out/r2s_dram_traffic.csv 2>&1 | head -16
python tools/traffic_from_ncu.py gpurun_out/r2s_dram_traffic.csv --quiet
