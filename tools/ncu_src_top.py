#!/usr/bin/env python
"""Top source lines of an `ncu --page source --csv --print-source cuda,sass` export by warp-stall samples.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > x.csv; ncu_src_top.py x.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
fname, hdr, ix, out = "", None, {}, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Name":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr, ix = r, {}
        for i, h in enumerate(hdr):
            ix.setdefault(h, i)
        continue
    if hdr is None or len(r) < len(hdr) or r[0] == "":
        continue
    out.append((fname, r))
si, ei = ix["# Samples"], ix["Instructions Executed"]


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


out = [(f, [c if i < 2 else c for i, c in enumerate(r)]) for f, r in out]
tot = sum(num(r[si]) for _, r in out)
tote = sum(num(r[ei]) for _, r in out)
print("total samples", tot, "warp-instructions", tote)
stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
out.sort(key=lambda fr: -num(fr[1][si]))
for f, r in out[:N]:
    st = sorted(((h[6:], num(r[ix[h]])) for h in stall if num(r[ix[h]]) > 0), key=lambda kv: -kv[1])[:3]
    print(f"{f[:22]:22s} {r[0]:>4s} {100 * num(r[si]) / tot:5.1f}% smp {100 * num(r[ei]) / max(tote, 1):5.1f}% ins  {r[1].strip()[:80]:80s} {st}")
