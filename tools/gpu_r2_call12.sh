#!/bin/bash
# round 2, call 12: start-up stagger of the per-layer kernel's CTAs (desynchronise memory-bound and tensor-bound phases)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
one() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('[$1] ms/step', round(d['ms_per_step'],2), 'gen', round(d['roofline']['ms'],2), 'value', round(d['value'],1))"; }
for wl in multilingual_v3_b256x128 baker_v1_gen_b64x640; do
  for v in "X=1" "WETTS_TC16_STAGGER=16000" "WETTS_TC16_STAGGER=32000" "WETTS_TC16_STAGGER=64000" "WETTS_TC16_STAGGER=128000"; do
    env $v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --workload $wl 2>/dev/null | one "$wl $v"
  done
done
env WETTS_TC16_STAGGER=48000 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2l_launches_stagger.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu --profile-range > gpurun_out/r2l_ncu.log 2>&1; echo "ncu rc=$?"
python tools/launches.py gpurun_out/r2l_launches_stagger.csv 2>&1 | head -12
