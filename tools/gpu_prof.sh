#!/bin/bash
# ncu --set full capture of selected tcgen05 conv launches inside the timed step: pairs of (skip, name)
set -x
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
while [ $# -ge 2 ]; do
  SKIP=$1; NAME=$2; shift 2
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
     -k regex:conv1d_tc_kernel -s $SKIP -c 1 -o gpurun_out/$NAME -f \
     python bench.py --steps 1 --warmup 3 --no-cpu --profile-range --batch 64 > gpurun_out/ncu_full_$NAME.log 2>&1
  tail -n 2 gpurun_out/ncu_full_$NAME.log
done
ls -la gpurun_out/*.ncu-rep
