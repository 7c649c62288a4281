#!/bin/bash
# round 2, call 23: 384-sample items as the default of the C=32 ResBlock2 stage (large launches): full GPU suite, A/B against
# the 128-sample items, the bench line as the driver runs it (pipelined D2H in the end-to-end leg), smoke
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
one() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('[$1] ms/step', round(d['ms_per_step'],2), 'gen', round(d['roofline']['ms'],2), 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],2))"; }
python __graft_entry__.py smoke > gpurun_out/r2y_smoke.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/r2y_smoke.log)"
timeout 1200 python -m pytest tests -q -x -m gpu > gpurun_out/r2y_tests.log 2>&1; echo "tests rc=$? $(tail -1 gpurun_out/r2y_tests.log)"
for rep in 1 2; do
for v in "X=1" "WETTS_MRF16_ITEM_C32=128"; do
  env $v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | one "rep$rep $v"
done
done
for v in "WETTS_TC16_TWO_CTAS=1" "WETTS_TC16_TWO_CTAS=1 WETTS_TC16R_MIN_TILES=1"; do
  env $v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | one "$v"
done
timeout 600 python bench.py > gpurun_out/r2y_bench_default.json 2> gpurun_out/r2y_bench_default.err; echo "bench default rc=$?"; cat gpurun_out/r2y_bench_default.json | one default
