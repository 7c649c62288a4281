#!/bin/bash
# ncu --set full of the three dominant kernels (batch 64, inside the bench's profiler range).  -k matches the function
# name without template arguments: the launch is picked by its position in the step (fused stages: C=64 first, then C=32;
# row-block-resident per-layer kernel: 20 text-encoder / DP launches and one flow `pre` come before the first in_layer)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
prof() { # name, kernel regex, skip
  timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$2 -s $3 -c 1 -o gpurun_out/r3b_$1 -f \
     python bench.py --steps 1 --warmup 2 --no-cpu --batch 64 --profile-range > gpurun_out/r3b_ncu_$1.log 2>&1
  echo "$1 rc=$? $(grep -c PROF gpurun_out/r3b_ncu_$1.log) $(tail -n 1 gpurun_out/r3b_ncu_$1.log | cut -c1-120)"
}
prof mrf16_c32 "fused_mrf16_kernel" 1
prof mrf16_c64 "fused_mrf16_kernel" 0
prof tc16r_flow_in "conv1d_tc16r_kernel" 21
ls -la gpurun_out/*.ncu-rep
