#!/bin/bash
# ncu --set full capture of selected tcgen05 conv launches inside the timed step
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
SKIP=${1:-81}; COUNT=${2:-2}; NAME=${3:-prof_tc}
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off \
   -k regex:conv1d_tc_kernel -s $SKIP -c $COUNT -o gpurun_out/$NAME -f \
   python bench.py --steps 1 --warmup 3 --no-cpu --profile-range --batch 64 > gpurun_out/ncu_full.log 2>&1
tail -n 5 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
