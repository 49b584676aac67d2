"""Drop-in proof for SURVEY.md §8 row a11: the UNMODIFIED reference ``SparseOneDecoder`` (vendored under
baseline/_ref/hipad, never committed) runs with ``sys.modules['projects.mmdet3d_plugin.ops']`` replaced.

CPU part (no GPU): the harness builds the 70.8 M-parameter stage-2 decoder through the mmcv stand-in and runs two
consecutive frames with the C oracle standing in for the op: exactly 24 aggregation calls per forward, in the order
det -> map -> plan -> ego, with the stage-2 shapes.
GPU part: the same decoder three times with identical weights — the reference's own ops package over its own CUDA
extension (oracle/_ref), ``hipad_b200.ops``, and ``hipad_b200.ops`` + ``hipad_b200.DeformableFeatureAggregation`` —
every returned tensor compared over three frames.
"""
import os
import sys
import types

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from harness import decoder as HD  # noqa: E402
from harness import vendor  # noqa: E402

needs_vendored = pytest.mark.skipif(vendor.vendor() is None, reason="baseline/_ref/hipad not vendored on this box")
STAGE2_CALLS = [(900, 13), (100, 300), (480, 90), (1, 13)] * 6
DECODER_TOL = 1e-4


def _oracle_ops():
    """ops package whose aggregation is the C oracle (CUDA-op semantics) — CPU stand-in for tests only."""
    import hipad_b200
    import oracle

    def daf(feat, shapes, starts, loc, w):
        out = oracle.forward(feat.detach().numpy(), shapes.int().numpy(), starts.int().numpy(),
                             loc.detach().numpy(), w.detach().numpy())
        return torch.from_numpy(out)

    mod = types.ModuleType("oracle_ops")
    mod.deformable_aggregation_function = daf
    mod.feature_maps_format = hipad_b200.ops.feature_maps_format
    return mod


@needs_vendored
def test_reference_decoder_runs_on_a_swapped_ops_package_cpu(oracle_mod):
    dec = HD.build_decoder(_oracle_ops(), hw=(352, 640))
    assert abs(sum(p.numel() for p in dec.parameters()) / 1e6 - 70.81) < 0.01      # 70.81 M parameters
    HD.use_sdpa_attention(dec)
    frames = HD.make_frames(2, bs=1, hw=(352, 640))
    with torch.no_grad():
        for levels, metas in frames:
            n0 = len(dec._hipad_ops.calls)
            out = HD.run_frame(dec, levels, metas)
            assert dec._hipad_ops.calls[n0:] == STAGE2_CALLS
            flat = HD.flatten_outputs(out)
            assert len(flat) > 30
            for name, t in flat.items():
                assert torch.isfinite(t.float()).all(), name


@needs_vendored
def test_vendored_files_are_byte_identical_to_the_reference():
    ref = "/root/reference"
    if not os.path.isdir(ref):
        pytest.skip("reference checkout absent")
    for rel in ("projects/mmdet3d_plugin/models/sparse_onedecoder.py", "projects/mmdet3d_plugin/models/blocks.py",
                "projects/mmdet3d_plugin/ops/__init__.py", "projects/configs/hipad_b2d_stage2.py"):
        assert open(os.path.join(ref, rel), "rb").read() == open(os.path.join(vendor.DST, rel), "rb").read(), rel


def _max_rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


@pytest.mark.gpu
@needs_vendored
@pytest.mark.parametrize("hw", [(352, 640), (256, 704)])
def test_decoder_outputs_match_reference_cuda_op(cuda_lib, hw):
    from oracle import build_ref
    if not build_ref.available():
        pytest.skip("oracle/_ref not built")
    dev = "cuda"
    ref = HD.build_decoder("reference", hw=hw, device=dev)
    variants = {"ours": HD.build_decoder("ours", hw=hw, device=dev),
                "ours_module": HD.build_decoder("ours_module", hw=hw, device=dev)}
    for d in variants.values():
        HD.copy_weights(d, ref)
    for d in [ref] + list(variants.values()):     # fp32 attention: flash-attn's fp16 would sit between the variants
        HD.use_sdpa_attention(d)
    frames = HD.make_frames(3, bs=1, hw=hw, device=dev)
    outs = {}
    with torch.no_grad():
        for name, d in dict(reference=ref, **variants).items():
            HD.reset(d)
            outs[name] = []
            for levels, metas in frames:
                n0 = len(d._hipad_ops.calls)
                outs[name].append(HD.flatten_outputs(HD.run_frame(d, levels, metas)))
                if name != "ours_module":        # the module variant calls the fused entry point, not the 5-arg op
                    assert d._hipad_ops.calls[n0:] == STAGE2_CALLS
    torch.cuda.synchronize()
    worst = {}
    for name in variants:
        for f, (a, b) in enumerate(zip(outs[name], outs["reference"])):
            assert a.keys() == b.keys()
            for k in a:
                assert a[k].shape == b[k].shape, (name, f, k)
                if a[k].dtype.is_floating_point:
                    worst[(name, f, k)] = _max_rel(a[k], b[k])
                else:
                    worst[(name, f, k)] = 0.0 if torch.equal(a[k], b[k]) else 1.0
    bad = {k: v for k, v in worst.items() if not v <= DECODER_TOL}
    assert not bad, sorted(bad.items(), key=lambda kv: -kv[1])[:10]


@pytest.mark.gpu
@needs_vendored
def test_decoder_runs_with_flash_attention(cuda_lib):
    """The reference's own attention path (flash-attn varlen kv-packed, fp16) with our op: runs and stays finite."""
    dec = HD.build_decoder("ours_module", hw=(352, 640), device="cuda")
    frames = HD.make_frames(2, bs=1, hw=(352, 640), device="cuda")
    with torch.no_grad():
        for levels, metas in frames:
            flat = HD.flatten_outputs(HD.run_frame(dec, levels, metas))
            for name, t in flat.items():
                assert torch.isfinite(t.float()).all(), name
