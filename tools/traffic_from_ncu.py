#!/usr/bin/env python
"""DRAM bytes per launch from an ncu CSV (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum) of one
bench step: prints a per-launch table and the generator total (launches from the conv_pre after the last flow cond_vec
launch to conv_post), and, with --json, updates profiles/generator_traffic.json for bench.py's roofline.traffic.

usage: traffic_from_ncu.py <csv> [--workload NAME --json profiles/generator_traffic.json --source TEXT]"""
import argparse, collections, csv, json, os, sys


def load(path):
    rows = [r for r in csv.DictReader(l for l in open(path) if l.startswith('"'))]
    by = collections.OrderedDict()
    for r in rows:
        m = by.setdefault(int(r["ID"]), {"name": r["Kernel Name"]})
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"].lower()
        scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}.get(unit, 1)
        m[r["Metric Name"]] = v * scale
    return list(by.values())


def short(name):
    n = name.replace("wetts::", "").replace("<unnamed>::", "").replace("void ", "")
    return n.split("(")[0][:44]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--workload")
    ap.add_argument("--json")
    ap.add_argument("--source")
    ap.add_argument("--quiet", action="store_true")
    a = ap.parse_args()
    L = load(a.csv)
    # the generator starts after the last cond_vec launch (conditioning of conv_pre) and ends with conv_post
    last_cond = max(i for i, m in enumerate(L) if "cond_vec" in m["name"])
    post = max(i for i, m in enumerate(L) if "conv_post" in m["name"])
    gen = L[last_cond + 1: post + 1]
    if not a.quiet:
        for i, m in enumerate(L):
            tag = "G" if last_cond < i <= post else " "
            print(f"{i:4d} {tag} {short(m['name']):46s} {m['gpu__time_duration.sum'] / 1e3:9.1f} us  read {m['dram__bytes_read.sum'] / 1e6:9.1f} MB  write {m['dram__bytes_write.sum'] / 1e6:9.1f} MB"
                  f"  {(m['dram__bytes_read.sum'] + m['dram__bytes_write.sum']) / m['gpu__time_duration.sum']:7.0f} GB/s")
    tot = sum(m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"] for m in gen)
    t = sum(m["gpu__time_duration.sum"] for m in gen)
    allb = sum(m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"] for m in L)
    print(f"generator: {len(gen)} launches, {tot / 1e9:.3f} GB DRAM traffic, {t / 1e6:.3f} ms (serialised) -> {tot / t:.0f} GB/s; whole step {allb / 1e9:.3f} GB")
    if a.json and a.workload:
        d = {}
        if os.path.exists(a.json):
            d = json.load(open(a.json))
            if "workload" in d:   # round-1 single-entry form
                d = {d["workload"]: {k: v for k, v in d.items() if k != "workload"}}
        d[a.workload] = {"generator_dram_bytes_per_step": tot, "generator_launches": len(gen), "step_dram_bytes": allb,
                         "source": a.source or os.path.basename(a.csv)}
        json.dump(d, open(a.json, "w"), indent=1)
        print("updated", a.json)


if __name__ == "__main__":
    sys.exit(main())
