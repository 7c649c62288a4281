"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list."""
import csv, re, sys
from collections import defaultdict
path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
rows = list(csv.DictReader(l for l in open(path) if l.startswith('"')))
# a capture with several metrics has one row per (launch, metric): keep the durations
if any(r.get('Metric Name') != 'gpu__time_duration.sum' for r in rows):
    rows = [r for r in rows if r.get('Metric Name') == 'gpu__time_duration.sum']
tot = 0
agg = defaultdict(lambda: [0, 0.0])
for r in rows:
    name = re.sub(r'\(.*', '', r['Kernel Name']).replace('wetts::<unnamed>::', '').replace('void ', '')
    t = float(r['Metric Value'].replace(',', '')) * {'ns': 1e-6, 'nsecond': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1.0, 'msecond': 1.0}.get(r.get('Metric Unit', 'ns').lower(), 1e-6)
    agg[name][0] += 1; agg[name][1] += t; tot += t
print(f"{len(rows)} launches, {tot:.3f} ms total (serialised, cold-cache: compare shares)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:55s} n={v[0]:4d} {v[1]:9.3f} ms {100*v[1]/tot:5.1f}%")
if "-v" in sys.argv:
    for r in rows:
        print(r['ID'], re.sub(r'\(.*', '', r['Kernel Name'])[-36:], r['Grid Size'], r['Block Size'], r['Metric Value'], r.get('Metric Unit', ''))
