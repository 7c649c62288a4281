#!/bin/bash
# round 2, call 10: clock64 phase profile of the pipelined per-layer kernel (one line per launch, cycles per work item and role)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { # tag, env..., -- bench args
  tag=$1; shift
  env "$@" > /dev/null 2>&1
}
env WETTS_TC16P=1 WETTS_TC16P_PROFILE=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu --workload baker_v3_gen_b64x640 > /dev/null 2> gpurun_out/r2j_prof_v3gen_aw.txt
env WETTS_TC16P=1 WETTS_TC16P_PROFILE=1 WETTS_TC16P_ALLWARPS=0 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu --workload baker_v3_gen_b64x640 > /dev/null 2> gpurun_out/r2j_prof_v3gen_split.txt
env WETTS_TC16P=1 WETTS_TC16P_PROFILE=1 WETTS_TC16_NMAX=64 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu --workload baker_v3_gen_b64x640 > /dev/null 2> gpurun_out/r2j_prof_v3gen_n64pp.txt
env WETTS_TC16P=1 WETTS_TC16P_PROFILE=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu --workload baker_v1_gen_b64x640 > /dev/null 2> gpurun_out/r2j_prof_v1gen_aw.txt
env WETTS_TC16P=1 WETTS_TC16P_PROFILE=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu --batch 64 > /dev/null 2> gpurun_out/r2j_prof_v3_b64_aw.txt
env WETTS_TC16P=1 WETTS_TC16P_PROFILE=1 WETTS_TC16_DEBUG_SKIP=3 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu --workload baker_v3_gen_b64x640 > /dev/null 2> gpurun_out/r2j_prof_v3gen_aw_skip3.txt
for f in v3gen_aw v3gen_split v3gen_n64pp v3gen_aw_skip3; do echo "== $f"; grep "tc16p profile" gpurun_out/r2j_prof_$f.txt | tail -8 | cut -c16-; done
echo "== v1gen_aw (last 40)"; grep "tc16p profile" gpurun_out/r2j_prof_v1gen_aw.txt | tail -40 | cut -c16- | awk 'NR%3==1'
echo "== v3 b64 (flow part)"; grep "tc16p profile" gpurun_out/r2j_prof_v3_b64_aw.txt | tail -50 | head -12 | cut -c16-
