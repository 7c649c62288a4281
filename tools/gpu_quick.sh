#!/bin/bash
# quick correctness + headline bench + launch list of one timed step; stops at the first failure
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 200 python tools/diag.py 1 v3_ragged v3_single > gpurun_out/diag_tc.log 2>&1; rc=$?; echo "diag rc=$rc"
grep -E "case|dec\(|flow\(|e2e o|Error|error|timed out" gpurun_out/diag_tc.log | head -30
if [ $rc -ne 0 ]; then tail -n 5 gpurun_out/diag_tc.log; exit 1; fi
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; rc=$?; echo "bench rc=$rc"
if [ $rc -ne 0 ]; then tail -n 5 gpurun_out/bench.err; exit 1; fi
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu --profile-range > gpurun_out/ncu_bench.log 2>&1
python tools/launches.py gpurun_out/launches.csv | head -20
cat gpurun_out/bench.json
WETTS_FUSED_RB_PROFILE=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu > /dev/null 2> gpurun_out/prof.log; grep -A2 "fused_rb profile" gpurun_out/prof.log | tail -6
