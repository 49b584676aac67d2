#!/usr/bin/env python
"""Decoder-level parity report (GPU): every tensor the unmodified reference decoder returns, over consecutive frames,
with `ours` / `ours_module` against the reference's own CUDA op — and the reference against ITSELF (its fp32 atomics
make it run-to-run non-deterministic, which bounds what "matching outputs" can mean).  Prints one JSON object."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from harness import decoder as HD  # noqa: E402


def run(dec, frames):
    HD.reset(dec)
    outs = []
    with torch.no_grad():
        for levels, metas in frames:
            outs.append({k: v.detach().clone() for k, v in HD.flatten_outputs(HD.run_frame(dec, levels, metas)).items()})
    torch.cuda.synchronize()
    return outs


def compare(a, b):
    """per frame: worst relative error over float tensors (max|a-b| / max|b|) and worst mismatch fraction over ints"""
    rep = []
    for fa, fb in zip(a, b):
        worst, worst_k, imis, imis_k = 0.0, None, 0.0, None
        for k in fb:
            if fa[k].dtype.is_floating_point:
                e = float((fa[k].double() - fb[k].double()).abs().max() / max(float(fb[k].double().abs().max()), 1e-30))
                if e > worst:
                    worst, worst_k = e, k
            else:
                m = float((fa[k] != fb[k]).float().mean())
                if m > imis:
                    imis, imis_k = m, k
        rep.append({"max_rel_err": worst, "at": worst_k, "int_mismatch_frac": imis, "int_at": imis_k})
    return rep


def report(hw=(352, 640), n_frames=3, sdpa=True):
    dev = "cuda"
    ref = HD.build_decoder("reference", hw=hw, device=dev)
    ours = HD.build_decoder("ours", hw=hw, device=dev)
    mod = HD.build_decoder("ours_module", hw=hw, device=dev)
    HD.copy_weights(ours, ref)
    HD.copy_weights(mod, ref)
    if sdpa:
        for d in (ref, ours, mod):
            HD.use_sdpa_attention(d)
    frames = HD.make_frames(n_frames, bs=1, hw=hw, device=dev)
    r1 = run(ref, frames)
    r2 = run(ref, frames)
    o = run(ours, frames)
    m = run(mod, frames)
    o2 = run(ours, frames)
    return {"hw": list(hw), "attention": "sdpa fp32" if sdpa else "flash-attn fp16",
            "reference_vs_itself": compare(r2, r1), "ours_vs_reference": compare(o, r1),
            "ours_module_vs_reference": compare(m, r1), "ours_vs_itself": compare(o2, o)}


if __name__ == "__main__":
    res = [report((352, 640), 3, True), report((352, 640), 3, False)]
    print(json.dumps(res))
