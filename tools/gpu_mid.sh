#!/bin/bash
# parity tests + headline bench + launch list
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu --profile-range > gpurun_out/ncu_bench.log 2>&1
python tools/launches.py gpurun_out/launches.csv > gpurun_out/launches.txt 2>&1; head -8 gpurun_out/launches.txt
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'gen_ms', d['roofline']['ms'])"
