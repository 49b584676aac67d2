#!/usr/bin/env python
"""BASELINE configs[1]: HiP-AD stage-2 unified decoder forward (det 900 + map 100 + plan 480 + ego queries, planning
deformable attention), bs=1 inference, through the UNMODIFIED reference decoder (harness/decoder.py) with
  reference    the reference's own ops package over its own CUDA op rebuilt for sm_100a (oracle/_ref)
  ours         hipad_b200.ops plugged in (5-argument op, reference module code around it)
  ours_module  hipad_b200.ops + hipad_b200.DeformableFeatureAggregation (fused projection + softmax + aggregation)
  ours_module_graph  the same, every module call replayed as one CUDA graph (module.graph_inference)
Times >= `frames` consecutive frames after 2 warm-up frames (temporal caches populated), CUDA events per frame, and
reports the DFA share of a forward (CUDA events around every aggregation call).  Prints one JSON object.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from harness import decoder as HD  # noqa: E402


def time_decoder(variant, hw=(352, 640), frames=30, bs=1, weights_from=None, device="cuda", seed=0):
    """ms per decoder forward (median and mean over `frames` frames), samples/s, DFA call count per forward."""
    dec = HD.build_decoder(variant, hw=hw, device=device, seed=seed)
    if weights_from is not None:
        HD.copy_weights(dec, weights_from)
    data = HD.make_frames(4, bs=bs, hw=hw, device=device)
    ops = dec._hipad_ops
    fms = [ops.feature_maps_format(levels) for levels, _ in data]
    img = data[0][0][0].new_zeros((bs, 6, 3, 8, 8))
    times = []
    with torch.no_grad():
        HD.reset(dec)
        for i in range(2):
            dec(img, fms[i % 4], data[i % 4][1])
        torch.cuda.synchronize()
        n0 = len(ops.calls)
        for i in range(frames):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            dec(img, fms[(i + 2) % 4], data[(i + 2) % 4][1])
            b.record()
            times.append((a, b))
        torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in times]
    med = float(np.median(ms))
    return dec, {"ms_per_forward_median": round(med, 3), "ms_per_forward_mean": round(float(np.mean(ms)), 3),
                 "samples_per_s": round(bs * 1e3 / med, 1), "frames": frames, "bs": bs,
                 "dfa_calls_per_forward": (len(ops.calls) - n0) / frames}


def dfa_share(dec, hw, bs=1, device="cuda", frames=5):
    """Fraction of the forward spent inside the aggregation modules (CUDA events around every *_deformable call)."""
    data = HD.make_frames(2, bs=bs, hw=hw, device=device)
    fms = [dec._hipad_ops.feature_maps_format(levels) for levels, _ in data]
    img = data[0][0][0].new_zeros((bs, 6, 3, 8, 8))
    spans = []
    hooks = []
    for name in ("det_deformable", "map_deformable", "plan_deformable", "ego_deformable"):
        for m in getattr(dec, name, []):
            def pre(mod, args, _s=spans):
                e = torch.cuda.Event(enable_timing=True); e.record(); mod._ev0 = e
            def post(mod, args, out, _s=spans):
                e = torch.cuda.Event(enable_timing=True); e.record(); _s.append((mod._ev0, e))
            hooks += [m.register_forward_pre_hook(pre), m.register_forward_hook(post)]
    tot = []
    with torch.no_grad():
        for i in range(frames + 1):
            if i == 1:
                spans.clear(); tot.clear()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); dec(img, fms[i % 2], data[i % 2][1]); b.record()
            tot.append((a, b))
        torch.cuda.synchronize()
    for h in hooks:
        h.remove()
    dfa_ms = sum(a.elapsed_time(b) for a, b in spans) / frames
    all_ms = sum(a.elapsed_time(b) for a, b in tot) / frames
    return {"dfa_module_ms_per_forward": round(dfa_ms, 3), "forward_ms": round(all_ms, 3),
            "dfa_share": round(dfa_ms / all_ms, 3)}


def run(hw_list=((352, 640), (256, 704)), frames=30, bs=1,
        variants=("reference", "ours", "ours_module", "ours_module_graph")):
    from oracle import build_ref
    res = {}
    for hw in hw_list:
        key = "%dx%d" % hw
        res[key] = {}
        base = None
        for v in variants:
            if v == "reference" and not build_ref.available():
                res[key][v] = {"unavailable": "oracle/_ref not built"}
                continue
            try:
                dec, r = time_decoder(v, hw=hw, frames=frames, bs=bs, weights_from=base)
                r.update(dfa_share(dec, hw, bs=bs))
                res[key][v] = r
                if base is None:
                    base = dec
                else:
                    del dec
            except Exception as e:  # keep the other variants
                res[key][v] = {"error": repr(e)[:300]}
            torch.cuda.empty_cache()
        del base
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=30)
    ap.add_argument("--bs", type=int, default=1)
    ap.add_argument("--hw", default="352x640,256x704")
    args = ap.parse_args()
    hws = [tuple(int(x) for x in s.split("x")) for s in args.hw.split(",")]
    print(json.dumps(run(hws, args.frames, args.bs)))
