#!/bin/bash
# Profiling visit: (a) DRAM bytes + duration of every launch of one timed step (full batch), (b) ncu --set full
# capture of the two fused MRF stage kernels (batch 64 to keep the replay short).
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --profile-from-start off --csv --log-file gpurun_out/dram.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu --profile-range > gpurun_out/ncu_dram.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:fused_resblock2 -c 2 -o gpurun_out/fused_rb -f \
    python bench.py --steps 1 --warmup 3 --no-cpu --profile-range --batch 64 > gpurun_out/ncu_full.log 2>&1
tail -n 2 gpurun_out/ncu_full.log
ls -la gpurun_out/*.ncu-rep
