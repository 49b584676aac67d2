#!/usr/bin/env python
"""One grouped stage-2 layer (det+map+plan+ego) forward + backward, for ncu.  usage: prof_group.py [reps] [bs]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import helpers as H
from hipad_b200 import _lib
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
bs = int(sys.argv[2]) if len(sys.argv) > 2 else 1
lib = _lib.get(); dev = torch.device("cuda")
LV = H.LEVELS_352x640
shapes, starts, F = H.level_tables(LV, 6)
C, G, L, CAMS = 256, 8, 4, 6
rng = np.random.default_rng(0)
feat = torch.from_numpy(rng.standard_normal((bs, F, C), dtype=np.float32)).to(dev)
sh, st = torch.from_numpy(shapes).to(dev), torch.from_numpy(starts).to(dev)
calls = []
for i, (kind, A, P) in enumerate((("det", 900, 13), ("map", 100, 300), ("plan", 480, 90), ("ego", 1, 13))):
    c = H.make_geo_case(10 + i, "det" if kind == "ego" else kind, bs, LV, (352, 640), A=A, P=P, with_feat=False)
    loc = c["loc"] if kind != "ego" else np.full_like(c["loc"], -0.5)
    d = dict(A=A, P=P, loc=torch.from_numpy(loc).to(dev), w=torch.from_numpy(c["weights"]).to(dev))
    d["g_loc"], d["g_w"] = torch.empty_like(d["loc"]), torch.empty_like(d["w"])
    calls.append(d)
a_total = sum(c["A"] for c in calls)
out = torch.empty((bs, a_total, C), device=dev)
go = torch.from_numpy(rng.standard_normal((bs, a_total, C), dtype=np.float32)).to(dev)
g_feat = torch.empty_like(feat)
t = _lib.call_table([(c["loc"].data_ptr(), c["w"].data_ptr(), c["g_loc"].data_ptr(), c["g_w"].data_ptr(), c["A"], c["P"]) for c in calls])
tp = ctypes.cast(t, ctypes.c_void_p)
wf = torch.empty(max(lib.hipad_dfa_group_forward_workspace_bytes(tp, 4, bs, CAMS, C), 256), dtype=torch.uint8, device=dev)
wb = torch.empty(max(lib.hipad_dfa_group_backward_workspace_bytes(tp, 4, bs, CAMS, F, C, L, G), 256), dtype=torch.uint8, device=dev)
s = torch.cuda.current_stream().cuda_stream
for _ in range(reps):
    _lib.check(lib.hipad_dfa_group_forward(0, out.data_ptr(), feat.data_ptr(), sh.data_ptr(), st.data_ptr(), tp, 4, bs, CAMS, F, C, L, G,
                                           wf.data_ptr(), wf.numel(), s), "fwd")
    _lib.check(lib.hipad_dfa_group_backward(0, 0, feat.data_ptr(), sh.data_ptr(), st.data_ptr(), tp, 4, go.data_ptr(), g_feat.data_ptr(),
                                            bs, CAMS, F, C, L, G, wb.data_ptr(), wb.numel(), s), "bwd")
torch.cuda.synchronize()
print("ok", float(out.abs().sum()), float(g_feat.abs().sum()))
