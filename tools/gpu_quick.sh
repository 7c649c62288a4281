#!/bin/bash
# quick correctness + headline bench + launch list of one timed step
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python tools/diag.py 1 > gpurun_out/diag_tc.log 2>&1; echo "diag rc=$?"
grep -E "case|dec\(|flow\(|e2e o|Error|error" gpurun_out/diag_tc.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu --profile-range > gpurun_out/ncu_bench.log 2>&1
cat gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
