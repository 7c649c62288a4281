#!/bin/bash
# round 2, call 2: full GPU suite on the f16 default build, bench variants (attention on the tensor pipe, start-up stagger,
# two CTAs per SM for the C = 64 stage, length-aware), launch list of one step
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -s -rA -m gpu > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED|ERROR" gpurun_out/r2_pytest_gpu.log | tail -15
grep -E "err/rms|text encoder|transformer flow|e2e z_p|logw|flow " gpurun_out/r2_pytest_gpu.log | head -60
run() { # name, env..., -- args
  name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 240 python bench.py --steps 5 --warmup 3 --no-cpu "$@" > gpurun_out/r2b_$name.json 2> gpurun_out/r2b_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2b_$name.json"))
    print("$name: ms/step", round(d["ms_per_step"],2), "value", round(d["value"],1), "gen ms", round(d["roofline"]["ms"],2), "frac", round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"],1))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/r2b_$name.err").read()[-800:])
PY
}
run default X=1 --
run noattn X=1 -- --attention-tc 0
run stagger10k WETTS_MRF16_STAGGER=10000 --
run stagger16k WETTS_MRF16_STAGGER=16000 --
run c64x2 WETTS_MRF16_C64_CTAS=2 --
run c64x2_stagger WETTS_MRF16_C64_CTAS=2 WETTS_MRF16_STAGGER=12000 --
run c32x2 WETTS_MRF16_CTAS=2 --
run lenaware X=1 -- --length-aware 1
WETTS_MRF16_C64_CTAS=2 WETTS_FUSED_RB_PROFILE=1 timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu > /dev/null 2> gpurun_out/r2_prof16_c64x2.log; grep -A2 "fused_mrf16 profile" gpurun_out/r2_prof16_c64x2.log | tail -6
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_fmt16.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu --profile-range > gpurun_out/r2_ncu_bench.log 2>&1
python tools/launches.py gpurun_out/r2_launches_fmt16.csv | head -30
