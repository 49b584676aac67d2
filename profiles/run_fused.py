#!/usr/bin/env python
"""Inference forward of one stage-2 DFA call: fused kernel (projection + group softmax + aggregation) vs the reference
chain in torch ops (project_points, softmax, two permute copies; blocks.py:134-161) feeding the unfused kernel.
usage: python profiles/run_fused.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, bench, helpers as H, hipad_b200
ops = hipad_b200.ops; dev = torch.device("cuda")
shapes, starts, F = H.level_tables(H.LEVELS_352x640, 6)
fm = [torch.randn((1, F, 256), device=dev), torch.from_numpy(shapes).to(dev).long(), torch.from_numpy(starts).to(dev).long()]
hipad_b200.ops._attach_host_tables(fm[1], fm[2], shapes.tolist(), starts.tolist())
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timed(fn, n=8):
    ts = []
    for _ in range(n):
        flush.sum(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    return float(np.median(ts[1:]))
for kind, A, P in bench.MODALITIES[:3]:
    c = H.make_geo_case(5, kind, 1, H.LEVELS_352x640, (352, 640), A=A, P=P, with_feat=False)
    bs, cams, _, C, L, _, _, G = c["dims"]
    kp, pm, wh = (torch.from_numpy(np.ascontiguousarray(c[k], dtype=np.float32)).to(dev) for k in ("key_points", "projection_mat", "image_wh"))
    logits = torch.from_numpy(c["logits"]).to(dev).reshape(bs, A, cams, L * P * G)
    def chain():
        p2d = hipad_b200.DeformableFeatureAggregation.project_points(kp, pm, wh).permute(0, 2, 3, 1, 4).contiguous()
        w = logits.reshape(bs, A, -1, G).softmax(dim=-2).reshape(bs, A, cams, L, P, G).permute(0, 1, 4, 2, 3, 5).contiguous()
        return ops.deformable_aggregation_function(*fm, p2d, w)
    with torch.no_grad():
        t_f = timed(lambda: ops.fused_deformable_aggregation(fm, kp, pm, wh, logits))
        t_c = timed(chain)
        loc = hipad_b200.DeformableFeatureAggregation.project_points(kp, pm, wh).permute(0, 2, 3, 1, 4).contiguous()
        w = logits.reshape(bs, A, -1, G).softmax(dim=-2).reshape(bs, A, cams, L, P, G).permute(0, 1, 4, 2, 3, 5).contiguous()
        t_u = timed(lambda: ops.deformable_aggregation_function(*fm, loc, w))
    print("%s %dx%d: fused %.1f us | torch project+softmax+permutes + unfused kernel %.1f us (kernel alone %.1f us)" % (kind, A, P, t_f, t_c, t_u))
