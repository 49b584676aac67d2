#!/usr/bin/env python
"""The four aggregation calls of one stage-2 decoder layer, ONE CALL AT A TIME through the C ABI (forward + backward),
for ncu: the per-launch DRAM traffic of every stage of the per-call step.  usage: prof_percall.py [reps]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import helpers as H
from hipad_b200 import _lib
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
lib = _lib.get(); dev = torch.device("cuda")
LV = H.LEVELS_352x640
shapes, starts, F = H.level_tables(LV, 6)
C, G, L, CAMS, bs = 256, 8, 4, 6, 1
rng = np.random.default_rng(0)
feat = torch.from_numpy(rng.standard_normal((bs, F, C), dtype=np.float32)).to(dev)
sh, st = torch.from_numpy(shapes).to(dev), torch.from_numpy(starts).to(dev)
g_feat = torch.empty_like(feat)
s = torch.cuda.current_stream().cuda_stream
calls = []
for i, (kind, A, P) in enumerate((("det", 900, 13), ("map", 100, 300), ("plan", 480, 90), ("ego", 1, 13))):
    c = H.make_geo_case(10 + i, "det" if kind == "ego" else kind, bs, LV, (352, 640), A=A, P=P, with_feat=False)
    loc = c["loc"] if kind != "ego" else np.full_like(c["loc"], -0.5)
    d = dict(A=A, P=P, loc=torch.from_numpy(loc).to(dev), w=torch.from_numpy(c["weights"]).to(dev),
             go=torch.from_numpy(rng.standard_normal((bs, A, C), dtype=np.float32)).to(dev), out=torch.empty((bs, A, C), device=dev))
    d["g_loc"], d["g_w"] = torch.empty_like(d["loc"]), torch.empty_like(d["w"])
    t = _lib.call_table([(d["loc"].data_ptr(), d["w"].data_ptr(), None, None, A, P)])
    d["tab"], d["tp"] = t, ctypes.cast(t, ctypes.c_void_p)
    d["wf"] = torch.empty(max(256, lib.hipad_dfa_group_forward_workspace_bytes(d["tp"], 1, bs, CAMS, C)), dtype=torch.uint8, device=dev)
    nb = lib.hipad_dfa_backward_workspace_bytes(bs, CAMS, F, C, L, A, P, G)
    d["wb"] = torch.empty(nb, dtype=torch.uint8, device=dev)
    calls.append(d)
for _ in range(reps):
    for d in calls:
        _lib.check(lib.hipad_dfa_group_forward(0, d["out"].data_ptr(), feat.data_ptr(), sh.data_ptr(), st.data_ptr(), d["tp"], 1,
                                               bs, CAMS, F, C, L, G, d["wf"].data_ptr(), d["wf"].numel(), s), "fwd")
        _lib.check(lib.hipad_dfa_backward_f32(feat.data_ptr(), sh.data_ptr(), st.data_ptr(), d["loc"].data_ptr(), d["w"].data_ptr(),
                                              d["go"].data_ptr(), g_feat.data_ptr(), d["g_loc"].data_ptr(), d["g_w"].data_ptr(),
                                              bs, CAMS, F, C, L, d["A"], d["P"], G, d["wb"].data_ptr(), d["wb"].numel(), s), "bwd")
torch.cuda.synchronize()
print("ok")
