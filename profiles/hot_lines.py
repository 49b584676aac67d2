#!/usr/bin/env python
"""usage: python profiles/hot_lines.py <report.ncu-rep | source.csv> [kernel-substring] [top]
Warp-stall samples of an `ncu --set full --import-source on` capture summed per CUDA source line (the export is
`ncu -i rep --page source --csv --print-source cuda,sass`), hottest lines first, with the dominant stall reasons."""
import collections
import csv
import subprocess
import sys

src = sys.argv[1]
sub = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
if src.endswith(".ncu-rep"):
    text = subprocess.run(["ncu", "-i", src, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                          capture_output=True, text=True).stdout
    rows = list(csv.reader(text.splitlines()))
else:
    rows = list(csv.reader(open(src)))
func = fil = None
agg = collections.defaultdict(lambda: [0, 0, "", collections.Counter()])
tot = collections.Counter()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fil = r[1].split("/")[-1]
        continue
    if r[0] in ("Function Name", "Kernel Name"):
        func = r[1][:70]
        continue
    if r[0] == "Line No":
        hdr = r
        si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
        stall = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    try:
        ln = int(r[0])
        s, ins = int(r[si] or 0), int(r[ii] or 0)
    except (ValueError, IndexError):
        continue
    k = (func, fil, ln)
    agg[k][0] += s
    agg[k][1] += ins
    agg[k][2] = r[1][:110].strip()
    for i, h in stall:
        try:
            agg[k][3][h[6:]] += int(r[i] or 0)
        except ValueError:
            pass
    tot[func] += s
for f in tot:
    if sub not in f:
        continue
    print("=====", f, "samples", tot[f])
    items = sorted(((k, v) for k, v in agg.items() if k[0] == f), key=lambda kv: -kv[1][0])
    for k, v in items[:top]:
        why = ",".join(f"{n}:{c}" for n, c in v[3].most_common(2))
        print(f"{v[0]:6d} {100 * v[0] / max(tot[f], 1):5.1f}%  inst {v[1]:8d}  {k[1]}:{k[2]:<4d} [{why}]  {v[2]}")
