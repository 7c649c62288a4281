#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (small first, then headline), ncu launch list.
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
python bench.py --workload multilingual_v3_b8x32 --steps 3 --warmup 3 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
python bench.py --workload baker_v1_b64x128 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_v1.json 2> gpurun_out/bench_v1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --batch 64 > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/pytest_gpu.log gpurun_out/smoke.log; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
