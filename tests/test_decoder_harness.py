"""Drop-in proof for SURVEY.md §8 row a11: the UNMODIFIED reference ``SparseOneDecoder`` (vendored under
baseline/_ref/hipad, never committed) runs with ``sys.modules['projects.mmdet3d_plugin.ops']`` replaced.

CPU part (no GPU): the harness builds the 70.8 M-parameter stage-2 decoder through the mmcv stand-in and runs two
consecutive frames with the C oracle standing in for the op: exactly 24 aggregation calls per forward, in the order
det -> map -> plan -> ego, with the stage-2 shapes.
GPU part: (i) every one of the 24 aggregation calls per frame of the reference decoder (running on its own CUDA op)
is replayed through ``hipad_b200`` on the same inputs, per call and grouped per layer: outputs within 1e-5; (ii) the
same decoder three times with identical weights — the reference's own ops package over its own CUDA extension
(oracle/_ref), ``hipad_b200.ops``, and ``hipad_b200.ops`` + ``hipad_b200.DeformableFeatureAggregation`` — whole-decoder
outputs compared within the reference's own run-to-run noise (its atomics make it non-deterministic).
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
def test_every_aggregation_call_of_the_decoder_matches_in_situ(cuda_lib, hw):
    """The reference decoder runs three consecutive frames on the reference's OWN CUDA op; the inputs of each of its
    24 aggregation calls per frame (model-generated key points, softmaxed weights, temporal caches populated) are then
    fed to hipad_b200's op: every output within 1e-5 of the reference op's.  This is the decoder-level parity that does
    not depend on how the rest of the network amplifies last-bit differences."""
    from oracle import build_ref
    import hipad_b200
    if not build_ref.available():
        pytest.skip("oracle/_ref not built")
    ref = HD.build_decoder("reference", hw=hw, device="cuda")
    HD.use_sdpa_attention(ref)
    frames = HD.make_frames(3, bs=1, hw=hw, device="cuda")
    worst = 0.0
    with torch.no_grad():
        for levels, metas in frames:
            ref._hipad_ops.record = []
            HD.run_frame(ref, levels, metas)
            rec, ref._hipad_ops.record = ref._hipad_ops.record, None
            assert [tuple(r[3].shape[1:3]) for r in rec] == STAGE2_CALLS
            for feat, shapes, starts, loc, w, out_ref in rec:
                out = hipad_b200.deformable_aggregation_function(feat, shapes, starts, loc, w)
                worst = max(worst, _max_rel(out, out_ref))
            # the four calls of each layer as ONE grouped launch: same outputs
            for layer in range(6):
                grp = rec[4 * layer:4 * layer + 4]
                outs = hipad_b200.deformable_aggregation_group(grp[0][0], grp[0][1], grp[0][2], [(r[3], r[4]) for r in grp])
                for o, r in zip(outs, grp):
                    worst = max(worst, _max_rel(o, r[5]))
    assert worst <= 1e-5, worst


@pytest.mark.gpu
@needs_vendored
def test_decoder_outputs_match_reference_cuda_op_within_its_own_noise(cuda_lib):
    """Whole-decoder outputs, three variants with identical weights.  The reference op accumulates with fp32 atomics,
    so the reference decoder does not reproduce ITSELF run to run (measured: ~1e-3 on the first frame; from the second
    frame on the temporal top-k caches amplify it to O(1) in a few rows).  Ours is bitwise reproducible, and on the
    first frame it is as close to the reference as the reference is to itself."""
    from oracle import build_ref
    if not build_ref.available():
        pytest.skip("oracle/_ref not built")
    hw, dev = (352, 640), "cuda"
    ref = HD.build_decoder("reference", hw=hw, device=dev)
    variants = {"ours": HD.build_decoder("ours", hw=hw, device=dev),
                "ours_module": HD.build_decoder("ours_module", hw=hw, device=dev)}
    for d in variants.values():
        HD.copy_weights(d, ref)
    for d in [ref] + list(variants.values()):     # fp32 attention: flash-attn's fp16 would sit between the variants
        HD.use_sdpa_attention(d)
    frames = HD.make_frames(2, bs=1, hw=hw, device=dev)

    def run(d, count_calls):
        HD.reset(d)
        outs = []
        with torch.no_grad():
            for levels, metas in frames:
                n0 = len(d._hipad_ops.calls)
                outs.append({k: v.detach().clone() for k, v in HD.flatten_outputs(HD.run_frame(d, levels, metas)).items()})
                if count_calls:
                    assert d._hipad_ops.calls[n0:] == STAGE2_CALLS
        torch.cuda.synchronize()
        return outs

    def worst(a, b):
        w = 0.0
        for k in b:
            assert a[k].shape == b[k].shape, k
            if a[k].dtype.is_floating_point:
                w = max(w, _max_rel(a[k], b[k]))
        return w

    r1, r2 = run(ref, True), run(ref, True)
    noise = worst(r2[0], r1[0])                               # reference vs itself, first frame
    for name, d in variants.items():
        o1 = run(d, name == "ours")
        o2 = run(d, False)
        for fa, fb in zip(o1, o2):                            # ours: bitwise reproducible, every frame
            for k in fa:
                assert torch.equal(fa[k], fb[k]), (name, k)
        assert worst(o1[0], r1[0]) <= max(DECODER_TOL, 3.0 * noise), (name, worst(o1[0], r1[0]), noise)
        for f in o1:
            for k, v in f.items():
                assert torch.isfinite(v.float()).all(), (name, k)


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


@pytest.mark.gpu
@needs_vendored
def test_decoder_with_graphed_modules_is_bitwise_the_eager_decoder(cuda_lib):
    """hipad_b200.DeformableFeatureAggregation.graph_inference: every aggregation module call of the unmodified
    reference decoder replayed as one CUDA graph (static input copies, one shared static feature buffer).  Same
    kernels, same arithmetic: every decoder output of three consecutive frames equals the eager module's bit for bit."""
    hw, dev = (352, 640), "cuda"
    eager = HD.build_decoder("ours_module", hw=hw, device=dev)
    graphed = HD.build_decoder("ours_module_graph", hw=hw, device=dev)
    HD.copy_weights(graphed, eager)
    for d in (eager, graphed):
        HD.use_sdpa_attention(d)
    frames = HD.make_frames(3, bs=1, hw=hw, device=dev)

    def run(d):
        HD.reset(d)
        outs = []
        with torch.no_grad():
            for levels, metas in frames:
                outs.append({k: v.detach().clone() for k, v in HD.flatten_outputs(HD.run_frame(d, levels, metas)).items()})
        torch.cuda.synchronize()
        return outs

    a, b = run(eager), run(graphed)
    import hipad_b200
    mods = [m for m in graphed.modules() if isinstance(m, hipad_b200.DeformableFeatureAggregation)]
    assert len(mods) == 24 and all(m.graph_inference and len(m._graphs) == 1 for m in mods)
    for fa, fb in zip(a, b):
        for k in fa:
            assert torch.equal(fa[k], fb[k]), k
    c = run(graphed)                       # replays only (no capture): still the same
    for fa, fc in zip(a, c):
        for k in fa:
            assert torch.equal(fa[k], fc[k]), k
