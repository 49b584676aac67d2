#!/usr/bin/env python
"""usage: python profiles/gather_roofline.py <full_layer_raw.csv.gz> <l2_gather_peak.json> [out.json]
Second roofline of the gather kernels: the bytes that actually cross the L2 -> SM fabric (ncu
`l1tex__m_xbar2l1tex_read_bytes.sum`) over the kernel's duration, against the L2-resident gather ceiling measured on the
same GPU by profiles/micro/l2_gather_peak.cu (random 1 KB rows, 2 x LDG.128 per lane: the access shape of these kernels).
The HBM roofline (bench.py `roofline`) is quoted on ALGORITHMIC bytes; this one says how far the gather is from the fabric
it really runs on (every distinct quad of a unit is fetched from L2, a feature row is re-read ~5x per layer)."""
import csv
import gzip
import json
import sys

raw, peakf = sys.argv[1], sys.argv[2]
peak = json.load(open(peakf))
ceil_gbs = max(peak["l2_resident_32MB_gbs"].values())
rows = list(csv.reader(gzip.open(raw, "rt")))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
out = {"l2_gather_ceiling_gbs": ceil_gbs, "hbm_random_row_gbs": max(peak["hbm_resident_4GB_gbs"].values()), "kernels": []}
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    try:
        xbar = float(r[col["l1tex__m_xbar2l1tex_read_bytes.sum"]])
        unit = rows[1][col["l1tex__m_xbar2l1tex_read_bytes.sum"]]
        dur = float(r[col["gpu__time_duration.sum"]])
        dunit = rows[1][col["gpu__time_duration.sum"]]
    except (KeyError, ValueError):
        continue
    xbar *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    dur *= {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "msecond": 1e-3}[dunit]
    gbs = xbar / dur / 1e9
    out["kernels"].append({"kernel": name[:70], "l2_to_sm_MB": round(xbar / 1e6, 1), "duration_us": round(dur * 1e6, 1),
                           "l2_to_sm_gbs": round(gbs, 0), "frac_of_l2_gather_ceiling": round(gbs / ceil_gbs, 3)})
text = json.dumps(out, indent=1)
print(text)
if len(sys.argv) > 3:
    open(sys.argv[3], "w").write(text + "\n")
