#!/usr/bin/env python
"""Tiny driver for ncu: runs the forward and the three backward kernels of one stage-2 DFA call.

usage: python profiles/run_kernels.py [det|map|plan|ego] [reps] [bs] [f32|bf16]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import bench
import hipad_b200
from hipad_b200 import _lib

kind = sys.argv[1] if len(sys.argv) > 1 else "det"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
bs = int(sys.argv[3]) if len(sys.argv) > 3 else 1
bf16 = len(sys.argv) > 4 and sys.argv[4] == "bf16"
mods = [m for m in bench.MODALITIES if m[0] == kind]
calls, shapes, starts, F = bench.make_calls(bs, seed=0, layers=1, modalities=mods)
c = calls[0]
dev = torch.device("cuda")
lib = _lib.get()
feat = torch.from_numpy(np.random.default_rng(0).standard_normal((bs, F, bench.C), dtype=np.float32)).to(dev)
if bf16:
    feat = feat.bfloat16()
sh, st = torch.from_numpy(shapes).to(dev), torch.from_numpy(starts).to(dev)
loc, w, go = (torch.from_numpy(c[k]).to(dev) for k in ("loc", "weights", "grad_out"))
out = torch.empty((bs, c["A"], bench.C), device=dev)
g_feat, g_loc, g_w = torch.empty_like(feat), torch.empty_like(loc), torch.empty_like(w)
dims = (bs, bench.CAMS, F, bench.C, 4, c["A"], c["P"], bench.G)
nb = lib.hipad_dfa_backward_workspace_bytes(*dims)
ws = torch.empty(nb, dtype=torch.uint8, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
s = torch.cuda.current_stream().cuda_stream
fwd = lib.hipad_dfa_forward_bf16 if bf16 else lib.hipad_dfa_forward_f32
ev = []
for r in range(reps):
    flush.sum()   # read-only L2 flush: leaves clean lines, no write-back tail in the first timed kernel
    e = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    e[0].record()
    _lib.check(fwd(out.data_ptr(), feat.data_ptr(), sh.data_ptr(), st.data_ptr(), loc.data_ptr(), w.data_ptr(), *dims, s), "fwd")
    e[1].record()
    zmask = 8 if os.environ.get("HIPAD_SEPARATE_ZERO") else 0
    for i, m in enumerate((1 | zmask, 2, 4)):
        _lib.check(lib.hipad_dfa_backward_stages(int(bf16), m, feat.data_ptr(), sh.data_ptr(), st.data_ptr(), loc.data_ptr(),
                                                 w.data_ptr(), go.data_ptr(), g_feat.data_ptr(), g_loc.data_ptr(),
                                                 g_w.data_ptr(), *dims, ws.data_ptr(), nb, s), "bwd")
        e[2 + i].record()
    ev.append(e)
torch.cuda.synchronize()
for e in ev[1:]:
    print(kind, "bs", bs, "us: fwd %.1f  bwd_sample+zero %.1f  compact+sort %.1f  rows+heavy %.1f" %
          tuple(e[i].elapsed_time(e[i + 1]) * 1e3 for i in range(4)))
