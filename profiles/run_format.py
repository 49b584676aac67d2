#!/usr/bin/env python
"""feature_maps_format: one transposing kernel vs the reference's cat + permute + flatten (torch ops), GB/s of the
useful bytes (one read + one write of every feature map).  usage: python profiles/run_format.py [bs] [f32|bf16]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, hipad_b200
bs = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dt = torch.bfloat16 if len(sys.argv) > 2 and sys.argv[2] == "bf16" else torch.float32
lv = [(88, 160), (44, 80), (22, 40), (11, 20)]
fm = [torch.randn((bs, 6, 256, h, w), device="cuda").to(dt) for h, w in lv]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def ref(f):
    return torch.cat([x.reshape(bs, 6, 256, -1) for x in f], dim=-1).permute(0, 1, 3, 2).flatten(1, 2).contiguous()
def timed(fn, n=10):
    ts = []
    for _ in range(n):
        flush.sum()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    return sorted(ts)[len(ts) // 2]
nbytes = 2 * sum(x.numel() * x.element_size() for x in fm)
t_ours = timed(lambda: hipad_b200.ops.format_feature_levels(fm))
t_ref = timed(lambda: ref(fm))
assert torch.equal(hipad_b200.ops.format_feature_levels(fm), ref(fm))
print("feature_maps_format bs %d %s: kernel %.1f us (%.0f GB/s of %.1f MB read+write) | torch cat+permute+flatten %.1f us"
      % (bs, str(dt)[6:], t_ours, nbytes / t_ours / 1e3, nbytes / 1e6, t_ref))
