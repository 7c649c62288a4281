#!/bin/bash
# planner experiments for the per-layer tcgen05 kernel: headline step time per variant
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for v in 0 1 2 3; do
  WETTS_TC_PLAN_VARIANT=$v timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_plan$v.json 2> gpurun_out/bench_plan$v.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_plan$v.json')); print('variant $v: ms/step', round(d['ms_per_step'],2), 'gen_ms', round(d['roofline']['ms'],2))" || tail -n 3 gpurun_out/bench_plan$v.err
done
