#!/usr/bin/env python
"""Per-CTA phase timelines of the layer's kernels (development build with -DHIPAD_DFA_TRACE):
   HIPAD_DFA_NVCC_EXTRA=-DHIPAD_DFA_TRACE python hip-ad_b200/build.py; python profiles/trace_kernels.py [bs] [f32|bf16]
Prints, per kernel: launch span, CTA count, per-phase share of the CTA-clocks, the longest CTAs and how busy the SMs were."""
import contextlib
import ctypes
import io
import json
import os
import sys

sys.argv = [sys.argv[0]] + (sys.argv[1:] or ["1"])
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "profiles"))
with contextlib.redirect_stdout(io.StringIO()):
    import run_group as R
import numpy as np
import torch

lib = R.lib
SLOTS, CTAS = 14, 1 << 14
for f in (lib.hipad_dfa_trace_read_gfeat, lib.hipad_dfa_trace_read_group):
    f.argtypes = [ctypes.c_void_p, ctypes.c_longlong]


def read(fn):
    host = np.zeros(CTAS * SLOTS, dtype=np.int64)
    fn(host.ctypes.data, host.size)
    return host.reshape(CTAS, SLOTS)


def run(fn, reader):
    read(reader)
    R.flush.view(torch.int64).sum()
    fn()
    torch.cuda.synchronize()
    return read(reader)


def summarize(name, t, phases, end_slot, info):
    live = np.nonzero((t[:, 0] != 0) & (t[:, end_slot] != 0))[0]
    t = t[live]
    t0 = t[:, 0].min()
    start = (t[:, 0] - t0) / 1e3
    end = (t[:, end_slot] - t0) / 1e3
    span = end.max()
    print(f"==== {name}: {len(live)} CTAs, span {span:.1f} us (first start to last end), SMs used {len(set(t[:, 1]))}")
    busy = (end - start).sum()
    print(f"  CTA-time sum {busy:.0f} us = {busy / span:.1f} CTAs resident on average; mean CTA {np.mean(end - start):.1f} us, "
          f"p50 {np.median(end - start):.1f}, p95 {np.percentile(end - start, 95):.1f}, max {np.max(end - start):.1f}")
    tot = {}
    for ph, fn in phases.items():
        tot[ph] = float(fn(t).sum())
    s = sum(tot.values())
    print("  phase share of CTA-clocks: " + ", ".join(f"{k} {v / s * 100:.0f}%" for k, v in tot.items()))
    # SM busy: fraction of the span during which the SM had at least one CTA of this kernel
    hist, _ = np.histogram(start, bins=10, range=(0, span))
    print("  starts per tenth of the span:", hist.tolist())
    hist, _ = np.histogram(end, bins=10, range=(0, span))
    print("  ends per tenth of the span:  ", hist.tolist())
    order = np.argsort(-(end - start))[:8]
    for i in order:
        print("   long CTA", int(live[i]), f"start {start[i]:.1f} end {end[i]:.1f}", {k: int(fn(t[i:i + 1])[0]) for k, fn in phases.items()},
              {k: int(t[i, s_]) for k, s_ in info.items()})
    return dict(name=name, ctas=len(live), span_us=float(span), cta_time_us=float(busy), phases=tot)


group_ph = {"setup(1)": lambda t: t[:, 3] - t[:, 2], "keys+sort+quads(2-4)": lambda t: t[:, 4] - t[:, 3], "rows+tables(5)": lambda t: t[:, 5],
            "gather(6)": lambda t: t[:, 6], "items(7)": lambda t: t[:, 7], "tail(8)": lambda t: t[:, 11] - t[:, 10],
            "other": lambda t: (t[:, 10] - t[:, 4]) - t[:, 5] - t[:, 6] - t[:, 7]}
out = []
out.append(summarize("group forward (4 calls)", run(lambda: R.fwd_group(R.calls, R.out_packed, R.wf_all), lib.hipad_dfa_trace_read_group),
                     group_ph, 12, {"items": 8, "quads": 9}))
out.append(summarize("group backward sample kernel", run(lambda: R.bwd_stage(1), lib.hipad_dfa_trace_read_group), group_ph, 12,
                     {"items": 8, "quads": 9}))
sort_ph = {"prefix": lambda t: t[:, 3] - t[:, 2], "scan": lambda t: t[:, 4] - t[:, 3], "sort": lambda t: t[:, 5] - t[:, 4],
           "emit": lambda t: t[:, 6] - t[:, 5], "seg": lambda t: t[:, 7] - t[:, 6]}
out.append(summarize("band sort", run(lambda: R.bwd_stage(2), lib.hipad_dfa_trace_read_gfeat), sort_ph, 9, {"n": 8}))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "trace_kernels.json"), "w"))
