#!/bin/bash
# round 2, first landing call: wide-fixture diagnostics on both conv routes, MMA numerics/timing probe, the f16 fused
# stage kernel (parity + speed + phase profile), bench on both operand formats
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 120 tools/ubench/mma_numerics > gpurun_out/r2_mma_numerics.txt 2>&1; echo "numerics rc=$?"
cat gpurun_out/r2_mma_numerics.txt
timeout 300 python -m pytest tests/test_mrf16_gpu.py -x -q -s -m gpu > gpurun_out/r2_mrf16_tests.log 2>&1; echo "mrf16 tests rc=$?"
grep -E "generator err|passed|failed|Error|error" gpurun_out/r2_mrf16_tests.log | head -20
timeout 200 python tools/diag.py 1 aishell3_long v3_tx128 > gpurun_out/r2_diag_tc1.log 2>&1; echo "diag tc1 rc=$?"
timeout 200 python tools/diag.py 0 aishell3_long v3_tx128 > gpurun_out/r2_diag_tc0.log 2>&1; echo "diag tc0 rc=$?"
cat gpurun_out/r2_diag_tc1.log gpurun_out/r2_diag_tc0.log
timeout 300 python -m pytest tests/test_zz_widecases_gpu.py -q -s -rA -m gpu > gpurun_out/r2_wide.log 2>&1; echo "wide rc=$?"
grep -E "text encoder|passed|failed|xfail|XPASS|assert" gpurun_out/r2_wide.log | head
for fmt in 32 16; do
  timeout 240 python bench.py --steps 5 --warmup 3 --no-cpu --tensor-format $fmt > gpurun_out/r2_bench_fmt$fmt.json 2> gpurun_out/r2_bench_fmt$fmt.err; echo "bench fmt$fmt rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_bench_fmt$fmt.json"))
    print("fmt$fmt", "ms/step", round(d["ms_per_step"],2), "value", round(d["value"],1), "gen ms", round(d["roofline"]["ms"],2), "frac", round(d["roofline"]["frac"],3), "launches/step", d["gpu_launches_per_step"])
except Exception as e:
    print("bench fmt$fmt parse failed", e); print(open("gpurun_out/r2_bench_fmt$fmt.err").read()[-1500:])
PY
done
WETTS_FUSED_RB_PROFILE=1 timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu --tensor-format 16 > /dev/null 2> gpurun_out/r2_prof16.log; grep -A2 "fused_mrf16 profile" gpurun_out/r2_prof16.log | tail -9
for ctas in 2; do
  WETTS_MRF16_CTAS=$ctas timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu --tensor-format 16 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('mrf16 ctas/sm(C=32)=$ctas ms/step', round(d['ms_per_step'],2), 'gen', round(d['roofline']['ms'],2))"
done
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu --tensor-format 16 --attention-tc 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('fmt16 + tensor-pipe attention: ms/step', round(d['ms_per_step'],2), 'gen', round(d['roofline']['ms'],2))"
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu --tensor-format 16 --attention-tc 1 --length-aware 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('fmt16 + attn + length-aware: ms/step', round(d['ms_per_step'],2), 'gen', round(d['roofline']['ms'],2))"
