#!/usr/bin/env python
"""Issue-path hygiene scan of the built library: for every kernel that issues tcgen05.mma, count the
UTCHMMA instructions, those under a PER-THREAD predicate (@Pn; the uniform @UPn form is fine), and the R2UR moves that
sit between the MMAs of an issue region (a vector -> uniform register move in front of an MMA means ptxas treated the
issue path as possibly divergent: every descriptor then takes a detour through the vector register file).

usage: sass_issue_scan.py [path/to/libwetts_b200.so]      (needs cuobjdump and c++filt; no GPU)"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def scan(lib):
    out = subprocess.run(f"cuobjdump -sass {lib} | c++filt", shell=True, capture_output=True, text=True, check=True).stdout
    cur, data = None, {}
    for line in out.split("\n"):
        m = re.search(r"Function : (.*)", line)
        if m:
            cur = m.group(1).replace("(anonymous namespace)::", "")
            cur = re.sub(r"\((wetts::|Tc|Fused|Attn).*$", "", cur).replace("void ", "").strip()
            data[cur] = []
            continue
        if cur and re.search(r"/\*[0-9a-f]{4}\*/", line):
            data[cur].append(line)
    rows = []
    for k, L in data.items():
        mm = [i for i, l in enumerate(L) if "UTCHMMA" in l]
        if not mm:
            continue
        regions, s, p = [], mm[0], mm[0]
        for i in mm[1:]:
            if i - p > 200:          # another issue loop (e.g. the second GEMM of the attention kernel)
                regions.append((s, p))
                s = i
            p = i
        regions.append((s, p))
        inside = sum(1 for a, b in regions for l in L[a:b + 1] if "R2UR" in l)
        pred = sum(1 for i in mm if re.search(r"@!?P\d", L[i]))
        total = sum(1 for l in L if "R2UR" in l)
        rows.append({"kernel": k, "mmas": len(mm), "regions": len(regions), "r2ur_between_mmas": inside,
                     "per_thread_predicated_mmas": pred, "r2ur_total": total})
    return rows


def is_profiling(name):
    return bool(re.search(r"fused_mrf16_kernel<.*, true, \d+>$", name) or re.search(r"conv1d_tc16r_kernel<\d+, true>$", name)
                or re.search(r"fused_resblock2_kernel<.*, true>$", name) or re.search(r"tc16p_kernel<true>$", name))


if __name__ == "__main__":
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "wetts_b200", "csrc", "libwetts_b200.so")
    for r in sorted(scan(lib), key=lambda r: (r["r2ur_between_mmas"], r["kernel"])):
        if is_profiling(r["kernel"]):
            continue
        print(f"{r['kernel'][:88]:88s} UTCHMMA={r['mmas']:3d} regions={r['regions']} R2UR_between_MMAs={r['r2ur_between_mmas']:3d} "
              f"per_thread_predicated_MMAs={r['per_thread_predicated_mmas']} R2UR_total={r['r2ur_total']}")
