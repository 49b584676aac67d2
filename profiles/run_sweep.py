#!/usr/bin/env python
"""BASELINE.json configs[3]: deformable-aggregation microbench sweep (anchors 900-3600, key points 7-32, f32 / bf16
feature maps), one call forward + backward through the C ABI, CUDA events, L2 flushed before every call.
usage: python profiles/run_sweep.py [bs] > sweep.md"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, bench, helpers as H
import hipad_b200
from hipad_b200 import _lib
bs = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = torch.device("cuda"); lib = _lib.get(); ops = hipad_b200.ops
shapes, starts, F = H.level_tables(H.LEVELS_352x640, 6)
sh, st = torch.from_numpy(shapes).to(dev), torch.from_numpy(starts).to(dev)
feat32 = torch.randn((bs, F, 256), device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
peak = 6549.8
print("| A | P | dtype | visible | fwd us | bwd us | B_fwd+B_bwd MB | GB/s | %% of %.0f GB/s |" % peak)
print("|---|---|---|---|---|---|---|---|---|")
for A in (900, 1800, 2700, 3600):
    for P in (7, 13, 20, 32):
        c = H.make_geo_case(A * 100 + P, "det", bs, H.LEVELS_352x640, (352, 640), A=A, P=P, with_feat=False)
        loc = torch.from_numpy(c["loc"]).to(dev); w = torch.from_numpy(c["weights"]).to(dev)
        go = torch.randn((bs, A, 256), device=dev)
        vis = float(((loc > 0) & (loc < 1)).all(-1).float().mean())
        for dt in ("f32", "bf16"):
            feat = feat32 if dt == "f32" else feat32.bfloat16()
            elem = 4 if dt == "f32" else 2
            by = bench.algorithmic_bytes(ops, sh, st, loc, bs, F, A, P, 4, elem)
            out = torch.empty((bs, A, 256), device=dev); g_feat = torch.empty_like(feat)
            g_loc, g_w = torch.empty_like(loc), torch.empty_like(w)
            dims = (bs, 6, F, 256, 4, A, P, 8)
            nb = lib.hipad_dfa_backward_workspace_bytes(*dims); ws = torch.empty(nb, dtype=torch.uint8, device=dev)
            s = torch.cuda.current_stream().cuda_stream
            import ctypes
            tab = _lib.call_table([(loc.data_ptr(), w.data_ptr(), None, None, A, P)]); tp = ctypes.cast(tab, ctypes.c_void_p)
            wf = torch.empty(max(256, lib.hipad_dfa_group_forward_workspace_bytes(tp, 1, bs, 6, 256)), dtype=torch.uint8, device=dev)
            tf, tb = [], []
            for rep in range(6):
                flush.sum()
                e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                e[0].record()
                _lib.check(lib.hipad_dfa_group_forward(int(dt == "bf16"), out.data_ptr(), feat.data_ptr(), sh.data_ptr(), st.data_ptr(), tp, 1,
                                                       bs, 6, F, 256, 4, 8, wf.data_ptr(), wf.numel(), s), "fwd")
                e[1].record()
                _lib.check(lib.hipad_dfa_backward_stages(int(dt == "bf16"), 7, feat.data_ptr(), sh.data_ptr(), st.data_ptr(), loc.data_ptr(),
                                                         w.data_ptr(), go.data_ptr(), g_feat.data_ptr(), g_loc.data_ptr(), g_w.data_ptr(),
                                                         *dims, ws.data_ptr(), nb, s), "bwd")
                e[2].record(); torch.cuda.synchronize()
                if rep:
                    tf.append(e[0].elapsed_time(e[1]) * 1e3); tb.append(e[1].elapsed_time(e[2]) * 1e3)
            f_us, b_us = float(np.median(tf)), float(np.median(tb))
            tot = by["fwd"] + by["bwd"]
            gbs = tot / ((f_us + b_us) * 1e-6) / 1e9
            print("| %d | %d | %s | %.3f | %.1f | %.1f | %.1f | %.0f | %.1f |" % (A, P, dt, vis, f_us, b_us, tot / 1e6, gbs, 100 * gbs / peak))
