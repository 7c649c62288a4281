#!/bin/bash
# ncu --set full of two per-layer tcgen05 launches inside the timed step: the text encoder's FFN conv_1
# (192 -> 768, k3: the dense "attention/FFN GEMM" the north_star asks a tensor-pipe figure for) and a flow WaveNet
# in_layer (192 -> 384, k5).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for pair in "2 tc_ffn1" "29 tc_flow_in"; do
  set -- $pair
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
     -k regex:conv1d_tc_kernel -s $1 -c 1 -o gpurun_out/$2 -f \
     python bench.py --steps 1 --warmup 3 --no-cpu --profile-range > gpurun_out/ncu_$2.log 2>&1
  tail -n 1 gpurun_out/ncu_$2.log
done
ls -la gpurun_out/*.ncu-rep
