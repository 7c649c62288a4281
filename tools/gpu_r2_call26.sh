#!/bin/bash
# round 2, last 1-GPU call (3.5 GPU-minutes left): tensor-pipe attention as two CTAs per SM (WETTS_ATTN_TC_CTAS=2): parity on the
# Tx >= 64 fixtures, ncu --set full of the kernel at the bench shape, per-launch time of both layouts, bench A/B
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
env WETTS_ATTN_TC_CTAS=2 timeout 90 python -m pytest tests/test_zz_widecases_gpu.py -q -x -m gpu > gpurun_out/r3d_tests_attn2.log 2>&1; echo "tests attn 2 CTAs rc=$? $(tail -1 gpurun_out/r3d_tests_attn2.log)"
env WETTS_ATTN_TC_CTAS=2 timeout 120 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:rel_attention_tc_kernel -s 0 -c 1 -o gpurun_out/r3d_attn_tc -f \
   python bench.py --steps 1 --warmup 2 --no-cpu --profile-range > gpurun_out/r3d_ncu_attn.log 2>&1; echo "ncu full rc=$?"
for c in 1 2; do
  env WETTS_ATTN_TC_CTAS=$c timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:rel_attention_tc_kernel -c 6 --csv \
     --log-file gpurun_out/r3d_attn_time_ctas$c.csv python bench.py --steps 1 --warmup 2 --no-cpu --profile-range > /dev/null 2>&1
  echo "ctas=$c: $(grep -o '"[0-9.,]*"$' gpurun_out/r3d_attn_time_ctas$c.csv | tr -d '"' | tr '\n' ' ')"
done
one() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('[$1] ms/step', round(d['ms_per_step'],2), 'gen', round(d['roofline']['ms'],2), 'value', round(d['value'],1))"; }
for v in "WETTS_ATTN_TC_CTAS=1" "WETTS_ATTN_TC_CTAS=2"; do
  env $v timeout 60 python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | one "$v"
done
