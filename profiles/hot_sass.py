#!/usr/bin/env python
"""usage: python profiles/hot_sass.py <source.csv> [kernel-substring] [top]
Per kernel of an `ncu --page source --csv` export: total samples, instruction count, and the hottest SASS
instructions (stall samples) with their dominant stall reason."""
import csv, sys
path = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ""; top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
rows = list(csv.reader(open(path)))
kernels, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = dict(name=r[1], hdr=None, ins=[]); kernels.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None and len(r) > 5:
        cur["ins"].append(r)
for k in kernels:
    if want not in k["name"]:
        continue
    h = {n: i for i, n in enumerate(k["hdr"])}
    stall_cols = [n for n in k["hdr"] if n.startswith("stall_") and "Not Issued" not in n]
    tot = sum(int(r[h["# Samples"]] or 0) for r in k["ins"])
    execd = sum(int(r[h["Instructions Executed"]] or 0) for r in k["ins"])
    print("====", k["name"][:110]); print("  samples", tot, " warp-instr executed", execd, " static instr", len(k["ins"]))
    agg = {}
    for r in k["ins"]:
        for s in stall_cols:
            agg[s] = agg.get(s, 0) + int(r[h[s]] or 0)
    print("  stall mix:", ", ".join("%s=%.1f%%" % (s[6:], 100.0 * v / max(tot, 1)) for s, v in sorted(agg.items(), key=lambda x: -x[1])[:7]))
    order = sorted(range(len(k["ins"])), key=lambda i: -int(k["ins"][i][h["# Samples"]] or 0))[:top]
    for i in sorted(order):
        r = k["ins"][i]
        n = int(r[h["# Samples"]] or 0)
        best = max(stall_cols, key=lambda s: int(r[h[s]] or 0))
        print("  %5d  %5.1f%%  #%4d exec=%8s  %-14s %s" % (n, 100.0 * n / max(tot, 1), i, r[h["Instructions Executed"]], best[6:], r[h["Source"]].strip()[:90]))
