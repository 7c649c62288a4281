#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (headline + variants), ncu launch list of the timed step.
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 python bench.py --steps 5 --warmup 3 --fused-resblock 0 --no-cpu > gpurun_out/bench_layer.json 2> gpurun_out/bench_layer.err
timeout 600 python bench.py --workload baker_v1_b64x128 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_v1.json 2> gpurun_out/bench_v1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu --profile-range > gpurun_out/ncu_bench.log 2>&1
tail -n 3 gpurun_out/pytest_gpu.log gpurun_out/smoke.log; cat gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
python tools/launches.py gpurun_out/launches.csv | head -12
