#!/usr/bin/env python
"""One stage-2 decoder layer (det 900x13 + map 100x300 + plan 480x90 + ego 1x13, 352x640, C=256, G=8): CUDA-event
times of the forward / backward per call and grouped, L2 flushed before every timed launch sequence.
usage: python profiles/run_group.py [bs] [f32|bf16] [HxW] [plan anchors]   (env knobs of the kernels apply)
       HxW = 352x640 (stage-2, default) | 256x704 (BASELINE configs[0] geometry) | 512x1408 (configs[4], with 48 plan anchors)"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import helpers as H
import hipad_b200
from hipad_b200 import _lib

bs = int(sys.argv[1]) if len(sys.argv) > 1 else 1
bf16 = len(sys.argv) > 2 and sys.argv[2] == "bf16"
lib = _lib.get()
dev = torch.device("cuda")
HW = tuple(int(x) for x in sys.argv[3].split("x")) if len(sys.argv) > 3 else (352, 640)
PLAN_A = int(sys.argv[4]) if len(sys.argv) > 4 else 480
LV = [(HW[0] // s_, HW[1] // s_) for s_ in (4, 8, 16, 32)]
shapes, starts, F = H.level_tables(LV, 6)
C, G, L, CAMS = 256, 8, 4, 6
rng = np.random.default_rng(0)
feat = torch.from_numpy(rng.standard_normal((bs, F, C), dtype=np.float32)).to(dev)
if bf16:
    feat = feat.bfloat16()
sh, st = torch.from_numpy(shapes).to(dev), torch.from_numpy(starts).to(dev)
MODS = (("det", 900, 13), ("map", 100, 300), ("plan", PLAN_A, 90), ("ego", 1, 13))
calls = []
for i, (kind, A, P) in enumerate(MODS):
    c = H.make_geo_case(10 + i, "det" if kind == "ego" else kind, bs, LV, HW, A=A, P=P, with_feat=False)
    loc = c["loc"] if kind != "ego" else np.full_like(c["loc"], -0.5)
    d = dict(kind=kind, A=A, P=P, loc=torch.from_numpy(loc).to(dev), w=torch.from_numpy(c["weights"]).to(dev),
             go=torch.from_numpy(rng.standard_normal((bs, A, C), dtype=np.float32)).to(dev))
    d["out"] = torch.empty((bs, A, C), device=dev)
    d["g_loc"], d["g_w"] = torch.empty_like(d["loc"]), torch.empty_like(d["w"])
    calls.append(d)
a_total = sum(c["A"] for c in calls)
out_packed = torch.empty((bs, a_total, C), device=dev)
go_packed = torch.cat([c["go"] for c in calls], dim=1).contiguous()
g_feat = torch.empty_like(feat)
flush = torch.zeros(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
stream = torch.cuda.current_stream().cuda_stream


def table(cs, bwd):
    t = _lib.call_table([(c["loc"].data_ptr(), c["w"].data_ptr(), c["g_loc"].data_ptr() if bwd else None,
                          c["g_w"].data_ptr() if bwd else None, c["A"], c["P"]) for c in cs])
    return t, ctypes.cast(t, ctypes.c_void_p)


def ws_fwd(cs):
    t, tp = table(cs, False)
    n = lib.hipad_dfa_group_forward_workspace_bytes(tp, len(cs), bs, CAMS, C)
    return torch.empty(max(n, 256), dtype=torch.uint8, device=dev)


def ws_bwd(cs):
    t, tp = table(cs, True)
    n = lib.hipad_dfa_group_backward_workspace_bytes(tp, len(cs), bs, CAMS, F, C, L, G)
    return torch.empty(max(n, 256), dtype=torch.uint8, device=dev)


def fwd_legacy(c):
    fn = lib.hipad_dfa_forward_bf16 if bf16 else lib.hipad_dfa_forward_f32
    _lib.check(fn(c["out"].data_ptr(), feat.data_ptr(), sh.data_ptr(), st.data_ptr(), c["loc"].data_ptr(), c["w"].data_ptr(),
                  bs, CAMS, F, C, L, c["A"], c["P"], G, stream), "fwd")


def fwd_group(cs, out, work):
    t, tp = table(cs, False)
    _lib.check(lib.hipad_dfa_group_forward(int(bf16), out.data_ptr(), feat.data_ptr(), sh.data_ptr(), st.data_ptr(), tp, len(cs),
                                           bs, CAMS, F, C, L, G, work.data_ptr(), work.numel(), stream), "group fwd")


def bwd_group(cs, go, work, flags=0):
    t, tp = table(cs, True)
    _lib.check(lib.hipad_dfa_group_backward(int(bf16), flags, feat.data_ptr(), sh.data_ptr(), st.data_ptr(), tp, len(cs),
                                            go.data_ptr(), g_feat.data_ptr(), bs, CAMS, F, C, L, G, work.data_ptr(),
                                            work.numel(), stream), "group bwd")


def timed(fn, reps=7, prep=None):
    ts = []
    for _ in range(reps):
        if prep is not None:
            prep()
        flush.view(torch.int64).sum()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return round(float(np.median(ts[2:])), 1)


res = {"bs": bs, "dtype": "bf16" if bf16 else "f32", "input_hw": list(HW), "plan_anchors": PLAN_A, "feature_MB": round(feat.numel() * feat.element_size() / 1e6, 1), "env": {k: v for k, v in os.environ.items() if k.startswith("HIPAD_")}}
wf = {c["kind"]: ws_fwd([c]) for c in calls}
wb = {c["kind"]: ws_bwd([c]) for c in calls}
wf_all, wb_all = ws_fwd(calls), ws_bwd(calls)
for c in calls:
    k = c["kind"]
    res[k] = {"fwd_legacy_us": timed(lambda: fwd_legacy(c)),
              "fwd_group1_us": timed(lambda: fwd_group([c], c["out"], wf[k])),
              "bwd_group1_us": timed(lambda: bwd_group([c], c["go"], wb[k]))}
res["layer"] = {"fwd_sum_legacy_us": round(sum(res[c["kind"]]["fwd_legacy_us"] for c in calls), 1),
                "fwd_sum_group1_us": round(sum(res[c["kind"]]["fwd_group1_us"] for c in calls), 1),
                "bwd_sum_group1_us": round(sum(res[c["kind"]]["bwd_group1_us"] for c in calls), 1),
                "fwd_grouped_us": timed(lambda: fwd_group(calls, out_packed, wf_all)),
                "bwd_grouped_us": timed(lambda: bwd_group(calls, go_packed, wb_all)),
                "bwd_grouped_accumulate_us": timed(lambda: bwd_group(calls, go_packed, wb_all, flags=1))}
def bwd_stage(mask):
    t, tp = table(calls, True)
    _lib.check(lib.hipad_dfa_group_backward_stages(int(bf16), 0, mask, feat.data_ptr(), sh.data_ptr(), st.data_ptr(), tp, len(calls),
                                                   go_packed.data_ptr(), g_feat.data_ptr(), bs, CAMS, F, C, L, G, wb_all.data_ptr(),
                                                   wb_all.numel(), stream), "group bwd stage")


bwd_group(calls, go_packed, wb_all)      # workspace holds a complete chain result for the isolated stages
# (stage 4 consumes what stage 2 left in the workspace -- the compaction kernel resets the list counters and the reduce
# queue -- so every timed classify + reduce is preceded by an untimed compaction + sort)
res["layer"]["bwd_grouped_stage_us"] = {"sample+zero": timed(lambda: bwd_stage(1)), "compact+sort": timed(lambda: bwd_stage(2)),
                                        "classify+reduce": timed(lambda: bwd_stage(4), prep=lambda: bwd_stage(2)),
                                        "serial_all": timed(lambda: [bwd_stage(1), bwd_stage(2), bwd_stage(4)])}
# serial 4-call layer through one stream (what the public per-call API does)
res["layer"]["fwd_serial4_legacy_us"] = timed(lambda: [fwd_legacy(c) for c in calls])
res["layer"]["fwd_serial4_group1_us"] = timed(lambda: [fwd_group([c], c["out"], wf[c["kind"]]) for c in calls])
res["layer"]["bwd_serial4_group1_us"] = timed(lambda: [bwd_group([c], c["go"], wb[c["kind"]]) for c in calls])
print(json.dumps(res))
