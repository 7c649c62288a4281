#!/bin/bash
# round 2, call 13: upper bound of a design whose activations arrive without SIMT staging: pipelined kernel with the staging
# work skipped (wrong values, fixed-frame generator workloads), N = 128 (14 drain warps) vs N = 64 with two accumulator slots
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for wl in baker_v3_gen_b64x640 baker_v1_gen_b64x640; do
for v in "X=1" "WETTS_TC16P=1 WETTS_TC16P_ALLWARPS=1 WETTS_TC16_DEBUG_SKIP=1" "WETTS_TC16_NMAX=64 WETTS_TC16P=1 WETTS_TC16_DEBUG_SKIP=1" "WETTS_TC16_NMAX=64 WETTS_TC16P=1 WETTS_TC16P_ALLWARPS=1 WETTS_TC16_DEBUG_SKIP=1"; do
  tag=${wl}_$(echo "$v" | tr ' =' '__')
  env $v timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2m_$tag.csv \
     python bench.py --steps 1 --warmup 2 --no-cpu --profile-range --workload $wl > gpurun_out/r2m_$tag.log 2>&1
  python - <<PY
import csv
try:
    rows=[r for r in csv.DictReader(l for l in open("gpurun_out/r2m_$tag.csv") if l.startswith('"'))]
    tc=[round(float(r["Metric Value"].replace(",",""))/1e3) for r in rows if "tc16" in r["Kernel Name"]]
    tot=sum(float(r["Metric Value"].replace(",","")) for r in rows)/1e6
    print("%-100s total %.2f ms; tc16 n=%d sum %.2f ms: %s" % ("$wl $v", tot, len(tc), sum(tc)/1e3, tc))
except Exception as e:
    print("$wl $v failed", e)
PY
done
done
