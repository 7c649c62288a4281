#!/bin/bash
# round 2, call 7: where does the per-layer kernel's time go?  Bench with staging and/or epilogue work skipped (results wrong
# on purpose) in the simple and the pipelined kernel; the launch list of each gives the per-layer times.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for v in "X=1" "WETTS_TC16_DEBUG_SKIP=1" "WETTS_TC16_DEBUG_SKIP=2" "WETTS_TC16_DEBUG_SKIP=3" "WETTS_TC16P=1" "WETTS_TC16P=1 WETTS_TC16_DEBUG_SKIP=1" "WETTS_TC16P=1 WETTS_TC16_DEBUG_SKIP=2" "WETTS_TC16P=1 WETTS_TC16_DEBUG_SKIP=3"; do
  tag=$(echo "$v" | tr ' =' '__')
  env $v timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2g_$tag.csv \
     python bench.py --steps 1 --warmup 2 --no-cpu --profile-range > gpurun_out/r2g_$tag.log 2>&1
  echo "== $v"
  python - <<PY
import csv
rows=[r for r in csv.DictReader(l for l in open("gpurun_out/r2g_$tag.csv") if l.startswith('"'))]
tc=[float(r["Metric Value"].replace(",",""))/1e3 for r in rows if "tc16" in r["Kernel Name"]]
tot=sum(float(r["Metric Value"].replace(",","")) for r in rows)/1e6
# order of tc16 launches: 0..25 encoder, 26..27 dp, 28 pre, 29 in, 30 rs, ... (see profiles/r02f)
print("total %.2f ms; tc16 launches %d sum %.2f ms; enc qkv %.0f o %.0f ffn1 %.0f ffn2 %.0f | flow pre %.0f in %.0f rs %.0f in %.0f rs(last) %.0f post %.0f | gen pre %.0f up0 %.0f rb128 %s up1 %.0f up2 %.0f us" % (
  tot, len(tc), sum(tc)/1e3, tc[0], tc[1], tc[2], tc[3], tc[28], tc[29], tc[30], tc[31], tc[36], tc[37], tc[68], tc[69], [round(x) for x in tc[70:76]], tc[76], tc[77]))
PY
done
