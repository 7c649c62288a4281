#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (headline with the CPU baseline leg), launch list of the timed
# step, memcheck of the smoke case.
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu --profile-range > gpurun_out/ncu_bench.log 2>&1
timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python __graft_entry__.py smoke > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/memcheck.log
tail -n 3 gpurun_out/pytest_gpu.log gpurun_out/smoke.log; tail -n 6 gpurun_out/memcheck.log; tail -n 2 gpurun_out/bench.err
python tools/launches.py gpurun_out/launches.csv > gpurun_out/launches.txt 2>&1; head -6 gpurun_out/launches.txt
