#!/bin/bash
# round 2 profiling call: (1) DRAM bytes of every launch of one step (generator traffic for bench.py's roofline.traffic),
# (2) ncu --set full of the dominant kernels at batch 64 (fused C=32 / C=64 stages, a flow in_layer, an FFN conv, attention)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off \
    --csv --log-file gpurun_out/r2_dram_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu --profile-range > gpurun_out/r2_ncu_dram.log 2>&1
echo "dram traffic rc=$?"
prof() { # name, kernel regex, skip
  timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$2 -s $3 -c 1 -o gpurun_out/r2_$1 -f \
     python bench.py --steps 1 --warmup 2 --no-cpu --batch 64 --profile-range > gpurun_out/r2_ncu_$1.log 2>&1
  echo "$1 rc=$? $(tail -n 1 gpurun_out/r2_ncu_$1.log)"
}
prof mrf16_c32 "fused_mrf16_kernel<32" 0
prof mrf16_c64 "fused_mrf16_kernel<64" 0
prof tc16_flow_in "conv1d_tc16_kernel<512" 29
prof tc16_ffn1 "conv1d_tc16_kernel<512" 2
prof attn_tc "rel_attention_tc_kernel" 0
ls -la gpurun_out/*.ncu-rep
