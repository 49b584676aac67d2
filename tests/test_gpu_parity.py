"""GPU parity tests (run on the B200 box: ``pytest -m gpu``).

Every test calls the product path — the ctypes C ABI of lib/libhipad_dfa.so, directly or through
``hipad_b200.ops`` — and checks it against (a) the CPU oracle on the same seeded inputs, (b) the
committed golden fixtures generated from the unmodified reference, (c) the reference's own CUDA op
built from its sources into oracle/_ref (when present), and (d) size-independent properties at
BASELINE.json's full sizes.

Tolerances (BASELINE.json north_star): fp32 1e-5 relative, bf16 2e-2 relative, both measured as
max|a-b| / max|ref| over the tensor; integer sampling indices and level offsets bit-exact.
"""
import os

import numpy as np
import pytest
import torch

import helpers as H

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-5
BF16_TOL = 2e-2
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def dev(x, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(x)).cuda()
    return t if dtype is None else t.to(dtype)


@pytest.fixture(scope="module")
def ops(cuda_lib):
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import hipad_b200
    return hipad_b200.ops


def run_fwd(ops, case, bf16=False):
    feat = dev(case["feat"], torch.bfloat16 if bf16 else None)
    out = ops.deformable_aggregation_function(feat, dev(case["shapes"]).long(), dev(case["starts"]).long(),
                                              dev(case["loc"]), dev(case["weights"]))
    torch.cuda.synchronize()
    return out.cpu().numpy()


def run_bwd(ops, case, bf16=False):
    feat = dev(case["feat"], torch.bfloat16 if bf16 else None).requires_grad_(True)
    loc = dev(case["loc"]).requires_grad_(True)
    w = dev(case["weights"]).requires_grad_(True)
    out = ops.deformable_aggregation_function(feat, dev(case["shapes"]).long(), dev(case["starts"]).long(), loc, w)
    out.backward(dev(case["grad_out"]))
    torch.cuda.synchronize()
    return (out.detach().cpu().numpy(), feat.grad.float().cpu().numpy(), loc.grad.cpu().numpy(), w.grad.cpu().numpy())


SMALL_CASES = {
    # name: (bs, cams, levels, C, G, A, P)
    "c256_g8_l4": (2, 6, [(16, 28), (8, 14), (4, 7), (2, 4)], 256, 8, 40, 13),
    "c128_g8_l4": (1, 6, [(16, 28), (8, 14), (4, 7), (2, 4)], 128, 8, 33, 7),
    "c64_g2_l3": (2, 3, [(12, 20), (6, 10), (3, 5)], 64, 2, 17, 5),
    "c512_g8_l2": (1, 2, [(9, 11), (5, 6)], 512, 8, 9, 4),
    "c48_g3_scalar": (3, 2, [(7, 9), (5, 3)], 48, 3, 7, 3),          # gd=16 vector path, L=2
    "c36_g3_gd12": (2, 2, [(7, 9), (5, 3), (2, 2)], 36, 3, 11, 6),   # gd=12 -> scalar path
    "c30_g5_odd": (1, 3, [(6, 5)], 30, 5, 5, 9),                     # C%4!=0 -> scalar path
    "c200_g8_scalar": (1, 2, [(6, 5), (3, 3)], 200, 8, 6, 5),        # gd=25 -> scalar path, NCH=8
    "map_like_slices": (1, 6, [(16, 28), (8, 14), (4, 7), (2, 4)], 256, 8, 10, 300),  # cluster S=8
    "plan_like_slices": (1, 6, [(16, 28), (8, 14), (4, 7), (2, 4)], 256, 8, 48, 90),
    "one_anchor": (1, 6, [(16, 28), (8, 14), (4, 7), (2, 4)], 256, 8, 1, 13),
    "tiny_maps": (2, 2, [(1, 1), (1, 2), (2, 1)], 32, 4, 6, 4),
}


def small_case(name, seed=0, **kw):
    bs, cams, lv, C, G, A, P = SMALL_CASES[name]
    return H.make_case(seed, bs, cams, lv, C, G, A, P, **kw)


@pytest.mark.parametrize("name", sorted(SMALL_CASES))
def test_forward_matches_oracle(ops, oracle_mod, name):
    case = small_case(name, seed=11)
    got = run_fwd(ops, case)
    ref = oracle_mod.forward(case["feat"], case["shapes"], case["starts"], case["loc"], case["weights"])
    assert got.shape == ref.shape
    assert rel_err(got, ref) <= FP32_TOL


@pytest.mark.parametrize("name", sorted(SMALL_CASES))
def test_backward_matches_oracle(ops, oracle_mod, name):
    case = small_case(name, seed=12, weight_kind="raw")
    out, g_feat, g_loc, g_w = run_bwd(ops, case)
    r_feat, r_loc, r_w = oracle_mod.backward(case["feat"], case["shapes"], case["starts"], case["loc"],
                                             case["weights"], case["grad_out"])
    assert rel_err(g_feat, r_feat) <= FP32_TOL
    assert rel_err(g_w, r_w) <= FP32_TOL
    assert rel_err(g_loc, r_loc) <= FP32_TOL
    # all gradients of invisible samples are exactly zero (cu:168-171, 232-235)
    vis = ((case["loc"] > 0) & (case["loc"] < 1)).all(-1)
    assert not g_loc[~vis].any() and not g_w[~vis].any()
    # rows never touched by a visible sample get an exact zero, not garbage
    assert not g_feat[r_feat == 0].any()


@pytest.mark.parametrize("name", ["c256_g8_l4", "c128_g8_l4", "c36_g3_gd12", "map_like_slices"])
def test_bf16_features(ops, oracle_mod, name):
    case = small_case(name, seed=13, weight_kind="raw")
    feat_bf = torch.tensor(case["feat"]).bfloat16().float().numpy()    # what the kernel actually sees
    got = run_fwd(ops, case, bf16=True)
    ref = oracle_mod.forward(feat_bf, case["shapes"], case["starts"], case["loc"], case["weights"])
    assert rel_err(got, ref) <= 1e-4                                   # same rounded inputs, fp32 accumulate
    ref32 = oracle_mod.forward(case["feat"], case["shapes"], case["starts"], case["loc"], case["weights"])
    assert rel_err(got, ref32) <= BF16_TOL                             # the stated bf16 contract
    _, g_feat, g_loc, g_w = run_bwd(ops, case, bf16=True)
    r_feat, r_loc, r_w = oracle_mod.backward(feat_bf, case["shapes"], case["starts"], case["loc"],
                                             case["weights"], case["grad_out"])
    assert rel_err(g_w, r_w) <= 1e-4 and rel_err(g_loc, r_loc) <= 1e-4
    assert rel_err(g_feat, r_feat) <= BF16_TOL                         # gradient stored as bf16


@pytest.mark.parametrize("name", ["c256_g8_l4", "c64_g2_l3", "tiny_maps", "plan_like_slices"])
def test_sampling_indices_bit_exact(ops, oracle_mod, name):
    case = small_case(name, seed=14)
    got = ops.sample_indices(dev(case["shapes"]), dev(case["starts"]), dev(case["loc"])).cpu().numpy()
    ref = oracle_mod.indices(case["shapes"], case["starts"], case["loc"])
    assert got.dtype == np.int32 and np.array_equal(got, ref)


def test_sampling_indices_bit_exact_full_size(ops, oracle_mod):
    case = H.make_geo_case(3, "det", 2, H.LEVELS_352x640, (352, 640), with_feat=False)
    # adversarial extras: locations that land exactly on pixel centres / edges at some level
    rng = np.random.default_rng(0)
    loc = case["loc"].copy()
    k = rng.integers(0, 160, size=loc[..., 0].shape)
    exact = rng.random(loc[..., 0].shape) < 0.3
    loc[..., 0] = np.where(exact, ((k + 0.5) / 160).astype(np.float32), loc[..., 0])
    got = ops.sample_indices(dev(case["shapes"]), dev(case["starts"]), dev(loc)).cpu().numpy()
    ref = oracle_mod.indices(case["shapes"], case["starts"], loc)
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("name", ["op_small", "op_c256", "op_odd"])
def test_golden_reference_torch_path(ops, name):
    """CUDA op vs outputs of the UNMODIFIED reference torch path (tests/golden/make_golden.py)."""
    from oracle import torch_path as tp
    d = np.load(os.path.join(GOLD, name + ".npz"))
    L = len([k for k in d.files if k.startswith("fmap")])
    fmaps = [torch.tensor(d[f"fmap{i}"]).cuda() for i in range(L)]
    col, shapes, starts = ops.feature_maps_format(fmaps)
    col = col.detach().requires_grad_(True)
    loc = dev(d["loc"]).requires_grad_(True)
    w = torch.tensor(d["weights"]).permute(0, 1, 4, 2, 3, 5).contiguous().cuda().requires_grad_(True)
    out = ops.deformable_aggregation_function(col, shapes, starts, loc, w)
    out.backward(dev(d["grad_out"]))
    assert rel_err(out.detach().cpu().numpy(), d["out"]) <= FP32_TOL
    g_col, _, _ = tp.flatten_feature_maps([torch.tensor(d[f"g_fmap{i}"]) for i in range(L)])
    assert rel_err(col.grad.cpu().numpy(), g_col.numpy()) <= FP32_TOL
    assert rel_err(loc.grad.cpu().numpy(), d["g_loc"]) <= FP32_TOL
    g_w_ref = torch.tensor(d["g_weights"]).permute(0, 1, 4, 2, 3, 5).contiguous().numpy()
    assert rel_err(w.grad.cpu().numpy(), g_w_ref) <= FP32_TOL


def _reference_ext():
    from oracle import build_ref
    if not build_ref.available():
        pytest.skip("oracle/_ref/deformable_aggregation_ext.so not built (reference tree absent at build time)")
    return build_ref.load()


@pytest.mark.parametrize("name", ["c256_g8_l4", "c128_g8_l4", "c64_g2_l3", "map_like_slices", "tiny_maps"])
def test_matches_reference_cuda_op(ops, name):
    """Ours vs the reference's own CUDA kernels (deformable_aggregation_cuda.cu rebuilt for sm_100a),
    arbitrary locations including the half-pixel band where the torch path differs (N1)."""
    ext = _reference_ext()
    case = small_case(name, seed=15, weight_kind="raw")
    feat, loc, w, go = dev(case["feat"]), dev(case["loc"]), dev(case["weights"]), dev(case["grad_out"])
    shapes, starts = dev(case["shapes"]), dev(case["starts"])
    ref_out = ext.deformable_aggregation_forward(feat, shapes, starts, loc, w)
    r_feat, r_loc, r_w = torch.zeros_like(feat), torch.zeros_like(loc), torch.zeros_like(w)
    ext.deformable_aggregation_backward(feat, shapes, starts, loc, w, go, r_feat, r_loc, r_w)
    out, g_feat, g_loc, g_w = run_bwd(ops, case)
    assert rel_err(out, ref_out.cpu().numpy()) <= FP32_TOL
    assert rel_err(g_feat, r_feat.cpu().numpy()) <= FP32_TOL
    assert rel_err(g_loc, r_loc.cpu().numpy()) <= FP32_TOL
    assert rel_err(g_w, r_w.cpu().numpy()) <= FP32_TOL


def test_matches_reference_cuda_op_stage2_det(ops):
    ext = _reference_ext()
    case = H.make_geo_case(5, "det", 1, H.LEVELS_352x640, (352, 640))
    feat, loc, w, go = dev(case["feat"]), dev(case["loc"]), dev(case["weights"]), dev(case["grad_out"])
    shapes, starts = dev(case["shapes"]), dev(case["starts"])
    ref_out = ext.deformable_aggregation_forward(feat, shapes, starts, loc, w)
    r_feat, r_loc, r_w = torch.zeros_like(feat), torch.zeros_like(loc), torch.zeros_like(w)
    ext.deformable_aggregation_backward(feat, shapes, starts, loc, w, go, r_feat, r_loc, r_w)
    out, g_feat, g_loc, g_w = run_bwd(ops, case)
    assert rel_err(out, ref_out.cpu().numpy()) <= FP32_TOL
    assert rel_err(g_feat, r_feat.cpu().numpy()) <= FP32_TOL
    assert rel_err(g_loc, r_loc.cpu().numpy()) <= FP32_TOL
    assert rel_err(g_w, r_w.cpu().numpy()) <= FP32_TOL


@pytest.mark.parametrize("kind,A,P", [("det", 900, 13), ("map", 100, 300), ("plan", 480, 90), ("det", 1, 13)])
def test_stage2_shapes_properties(ops, kind, A, P):
    """BASELINE.json stage-2 shapes at 352x640, checked through size-independent properties:
    bitwise determinism, linearity in the weights, and the adjoint identities
    <grad_out, out> = <g_w, w> = <g_feat, feat> (the op is bilinear in (feat, w))."""
    case = H.make_geo_case(7, kind, 2, H.LEVELS_352x640, (352, 640), A=A, P=P)
    out1, g_feat1, g_loc1, g_w1 = run_bwd(ops, case)
    out2, g_feat2, g_loc2, g_w2 = run_bwd(ops, case)
    for a, b in ((out1, out2), (g_feat1, g_feat2), (g_loc1, g_loc2), (g_w1, g_w2)):
        assert np.array_equal(a, b), "results must be bitwise reproducible"
    go = case["grad_out"].astype(np.float64)
    lhs = float((go * out1).sum())
    assert abs(float((g_w1.astype(np.float64) * case["weights"]).sum()) - lhs) <= 1e-4 * max(1.0, abs(lhs))
    assert abs(float((g_feat1.astype(np.float64) * case["feat"]).sum()) - lhs) <= 1e-4 * max(1.0, abs(lhs))
    rng = np.random.default_rng(1)
    w2 = rng.standard_normal(case["weights"].shape, dtype=np.float32)
    c2 = dict(case, weights=w2)
    c3 = dict(case, weights=case["weights"] + w2)
    assert rel_err(run_fwd(ops, c3), out1.astype(np.float64) + run_fwd(ops, c2)) <= 2e-5
    vis = ((case["loc"] > 0) & (case["loc"] < 1)).all(-1)
    assert not g_loc1[~vis].any() and not g_w1[~vis].any()


def test_stage2_det_vs_oracle_full_size(ops, oracle_mod):
    case = H.make_geo_case(8, "det", 1, H.LEVELS_352x640, (352, 640))
    out, g_feat, g_loc, g_w = run_bwd(ops, case)
    ref = oracle_mod.forward(case["feat"], case["shapes"], case["starts"], case["loc"], case["weights"])
    r_feat, r_loc, r_w = oracle_mod.backward(case["feat"], case["shapes"], case["starts"], case["loc"],
                                             case["weights"], case["grad_out"])
    assert rel_err(out, ref) <= FP32_TOL
    assert rel_err(g_feat, r_feat) <= FP32_TOL
    assert rel_err(g_loc, r_loc) <= FP32_TOL
    assert rel_err(g_w, r_w) <= FP32_TOL


def test_all_samples_invisible_gives_zeros(ops):
    """The ego query's key points sit inside the ego box and are seen by no camera:
    the op must return exact zeros and zero gradients, with no pre-zeroed buffers to lean on."""
    case = small_case("c256_g8_l4", seed=16)
    case["loc"] = np.full_like(case["loc"], -0.25)
    out, g_feat, g_loc, g_w = run_bwd(ops, case)
    assert not out.any() and not g_feat.any() and not g_loc.any() and not g_w.any()


def test_dense_locations_all_visible(ops, oracle_mod):
    case = small_case("c128_g8_l4", seed=17, frac_inside=1.0)
    case["loc"] = np.random.default_rng(3).uniform(0.001, 0.999, case["loc"].shape).astype(np.float32)
    out, g_feat, g_loc, g_w = run_bwd(ops, case)
    ref = oracle_mod.forward(case["feat"], case["shapes"], case["starts"], case["loc"], case["weights"])
    r_feat, r_loc, r_w = oracle_mod.backward(case["feat"], case["shapes"], case["starts"], case["loc"],
                                             case["weights"], case["grad_out"])
    assert rel_err(out, ref) <= FP32_TOL and rel_err(g_feat, r_feat) <= FP32_TOL
    assert rel_err(g_loc, r_loc) <= FP32_TOL and rel_err(g_w, r_w) <= FP32_TOL


def test_large_bucket_uses_global_sort_path(ops, oracle_mod):
    """> 24576 visible samples in one (b, cam, level) bucket: the sort spills to the workspace."""
    case = H.make_case(21, 1, 1, [(24, 40), (6, 10)], 32, 4, 900, 32, frac_inside=1.0, weight_kind="raw")
    case["loc"] = np.random.default_rng(4).uniform(0.001, 0.999, case["loc"].shape).astype(np.float32)
    out, g_feat, g_loc, g_w = run_bwd(ops, case)
    r_feat, r_loc, r_w = oracle_mod.backward(case["feat"], case["shapes"], case["starts"], case["loc"],
                                             case["weights"], case["grad_out"])
    assert rel_err(g_feat, r_feat) <= FP32_TOL and rel_err(g_w, r_w) <= FP32_TOL and rel_err(g_loc, r_loc) <= FP32_TOL


def test_fused_forward_matches_unfused(ops, oracle_mod):
    """Fused projection + softmax + aggregation vs the reference op chain (blocks.py:134-161)."""
    import hipad_b200
    for kind, A, P in (("det", 64, 13), ("plan", 16, 90), ("map", 5, 300)):
        case = H.make_geo_case(9, kind, 2, [(16, 28), (8, 14), (4, 7), (2, 4)], (64, 112), A=A, P=P)
        bs, cams, F, C, L, A_, P_, G = case["dims"]
        fm = [dev(case["feat"]), dev(case["shapes"]).long(), dev(case["starts"]).long()]
        logits = dev(case["logits"]).reshape(bs, A, cams, L * P * G)
        kp, pm, wh = dev(case["key_points"]), dev(case["projection_mat"]), dev(case["image_wh"])
        out, loc = ops.fused_deformable_aggregation(fm, kp, pm, wh, logits, return_locations=True)
        # reference chain in torch on the GPU + our unfused op
        p2d = hipad_b200.DeformableFeatureAggregation.project_points(kp, pm, wh).permute(0, 2, 3, 1, 4).contiguous()
        w = logits.reshape(bs, A, -1, G).softmax(dim=-2).reshape(bs, A, cams, L, P, G)
        w = w.permute(0, 1, 4, 2, 3, 5).contiguous()
        # sampling locations: not bit-identical to cuBLAS bmm, but within a few ulp
        assert rel_err(loc.cpu().numpy(), p2d.cpu().numpy()) <= 1e-5
        vis_f = ((loc > 0) & (loc < 1)).all(-1)
        vis_r = ((p2d > 0) & (p2d < 1)).all(-1)
        mism = float((vis_f != vis_r).float().mean())
        assert mism <= 1e-3
        # aggregate on OUR locations so both paths see identical samples
        ref = ops.deformable_aggregation_function(*fm, loc, w)
        assert rel_err(out.cpu().numpy(), ref.cpu().numpy()) <= FP32_TOL
        cpu = oracle_mod.forward(case["feat"], case["shapes"], case["starts"], loc.cpu().numpy(), w.cpu().numpy())
        assert rel_err(out.cpu().numpy(), cpu) <= FP32_TOL


def _load_module_golden(name):
    import hipad_b200
    d = np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False)
    kind = str(d["kind"])
    embed = d["instance_feature"].shape[-1]
    G = int(d["num_groups"])
    L = len([k for k in d.files if k.startswith("fmap")])
    if kind == "det":
        kps = dict(type="SparseBox3DKeyPointsGenerator", num_learnable_pts=6,
                   fix_scale=[[0, 0, 0], [0.45, 0, 0], [-0.45, 0, 0], [0, 0.45, 0],
                              [0, -0.45, 0], [0, 0, 0.45], [0, 0, -0.45]])
    else:
        kps = dict(type="SparsePoint3DKeyPointsGenerator", embed_dims=embed, num_sample=6,
                   num_learnable_pts=3, fix_height=(0, 0.5, -0.5, 1, -1), ground_height=-1.84023)
    m = hipad_b200.DeformableFeatureAggregation(
        embed_dims=embed, num_groups=G, num_levels=L, num_cams=6, attn_drop=0.15,
        use_deformable_func=True, use_camera_embed=True, residual_mode="cat", kps_generator=kps)
    sd = {k[3:]: torch.tensor(d[k]) for k in d.files if k.startswith("sd.")}
    missing, unexpected = m.load_state_dict(sd, strict=True)
    return d, m.cuda().eval(), L


@pytest.mark.parametrize("name", ["module_det", "module_plan", "module_det_daf", "module_det_c256"])
@pytest.mark.parametrize("fused", [True, False])
def test_module_matches_reference_module(ops, name, fused):
    """Whole DeformableFeatureAggregation.forward (key points -> weights -> projection -> op ->
    output_proj -> cat) against the reference module run unmodified in the build container, with
    the reference's own state_dict loaded strictly."""
    d, m, L = _load_module_golden(name)
    m.fused_inference = fused
    fmaps = [torch.tensor(d[f"fmap{i}"]).float().cuda() for i in range(L)]
    fm = ops.feature_maps_format(fmaps)
    metas = dict(projection_mat=dev(d["projection_mat"]), image_wh=dev(d["image_wh"]))
    with torch.no_grad():
        out = m(dev(d["instance_feature"]), dev(d["anchor"]), dev(d["anchor_embed"]), fm, metas)
    assert rel_err(out.cpu().numpy(), d["out"]) <= 2e-5


@pytest.mark.parametrize("name", ["module_det", "module_det_c256"])
def test_module_graph_inference_equals_eager(ops, name):
    """graph_inference: the module's inference forward replayed as one CUDA graph.  Bitwise the eager result, across
    changed inputs, a NEW feature tensor (new frame), an in-place update of the same tensor, and a second signature."""
    d, m, L = _load_module_golden(name)
    fmaps = [torch.tensor(d[f"fmap{i}"]).float().cuda() for i in range(L)]
    metas = dict(projection_mat=dev(d["projection_mat"]), image_wh=dev(d["image_wh"]))
    inst, anchor, emb = dev(d["instance_feature"]), dev(d["anchor"]), dev(d["anchor_embed"])
    g = torch.Generator(device="cuda").manual_seed(3)

    def both(inst_, anchor_, emb_, fm_):
        with torch.no_grad():
            m.graph_inference = False
            want = m(inst_, anchor_, emb_, fm_, metas)
            m.graph_inference = True
            got = m(inst_, anchor_, emb_, fm_, metas)
        assert torch.equal(want, got)
        return got

    fm = ops.feature_maps_format(fmaps)
    first = both(inst, anchor, emb, fm)
    assert rel_err(first.cpu().numpy(), d["out"]) <= 2e-5
    assert len(m._graphs) == 1
    both(inst + 0.25 * torch.randn(inst.shape, device="cuda", generator=g), anchor, emb, fm)          # new inputs, replay
    fm2 = ops.feature_maps_format([f * 0.5 + 1.0 for f in fmaps])                                     # a new frame
    second = both(inst, anchor, emb, fm2)
    assert not torch.equal(first, second)
    fm2[0].mul_(2.0)                                                                                  # in-place update
    third = both(inst, anchor, emb, fm2)
    assert not torch.equal(second, third)
    both(inst[:, :5].contiguous(), anchor[:, :5].contiguous(), emb[:, :5].contiguous(), fm2)           # second signature
    assert len(m._graphs) == 2
    both(inst, anchor, emb, fm)                                                                       # back to frame one
    m.train()                                                                                         # training: eager path
    out = m(inst, anchor, emb, fm, metas)
    assert out.requires_grad and len(m._graphs) == 2
    m.eval()
    m.float()                                     # moving / casting the module drops the captured graphs
    assert len(m._graphs) == 0
    both(inst, anchor, emb, fm)
    assert len(m._graphs) == 1
    import copy
    assert len(copy.deepcopy(m)._graphs) == 0     # and they do not travel with a copy of the module


def test_module_training_path_backward(ops):
    d, m, L = _load_module_golden("module_det")
    m.train()
    fmaps = [torch.tensor(d[f"fmap{i}"]).cuda().requires_grad_(True) for i in range(L)]
    fm = ops.feature_maps_format(fmaps)
    metas = dict(projection_mat=dev(d["projection_mat"]), image_wh=dev(d["image_wh"]))
    inst = dev(d["instance_feature"]).requires_grad_(True)
    out = m(inst, dev(d["anchor"]), dev(d["anchor_embed"]), fm, metas)
    out.square().mean().backward()
    assert all(f.grad is not None and torch.isfinite(f.grad).all() for f in fmaps)
    assert inst.grad is not None and m.weights_fc.weight.grad is not None


def test_non_default_stream_and_graph_capture(ops, oracle_mod):
    """Launches follow torch's current stream (the reference uses the legacy default stream) and
    are CUDA-graph capturable."""
    case = small_case("c256_g8_l4", seed=18)
    feat, loc, w = dev(case["feat"]), dev(case["loc"]), dev(case["weights"])
    shapes, starts = dev(case["shapes"]).long(), dev(case["starts"]).long()
    ref = oracle_mod.forward(case["feat"], case["shapes"], case["starts"], case["loc"], case["weights"])
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            out = ops.deformable_aggregation_function(feat, shapes, starts, loc, w)
    s.synchronize()
    assert rel_err(out.cpu().numpy(), ref) <= FP32_TOL
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out_g = ops.deformable_aggregation_function(feat, shapes, starts, loc, w)
    out_g.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert rel_err(out_g.cpu().numpy(), ref) <= FP32_TOL


def test_frozen_features_skip_feature_gradient(ops, oracle_mod):
    case = small_case("c128_g8_l4", seed=19, weight_kind="raw")
    feat = dev(case["feat"])                                   # requires_grad False
    loc = dev(case["loc"]).requires_grad_(True)
    w = dev(case["weights"]).requires_grad_(True)
    out = ops.deformable_aggregation_function(feat, dev(case["shapes"]).long(), dev(case["starts"]).long(), loc, w)
    out.backward(dev(case["grad_out"]))
    _, r_loc, r_w = oracle_mod.backward(case["feat"], case["shapes"], case["starts"], case["loc"],
                                        case["weights"], case["grad_out"])
    assert rel_err(loc.grad.cpu().numpy(), r_loc) <= FP32_TOL and rel_err(w.grad.cpu().numpy(), r_w) <= FP32_TOL


def test_error_behaviour(ops, cuda_lib):
    case = small_case("tiny_maps", seed=20)
    with pytest.raises(RuntimeError):                          # CPU tensors: no CPU path exists
        ops.deformable_aggregation_function(torch.tensor(case["feat"]), torch.tensor(case["shapes"]),
                                            torch.tensor(case["starts"]), torch.tensor(case["loc"]),
                                            torch.tensor(case["weights"]))
    with pytest.raises(ValueError):
        ops.deformable_aggregation_function(dev(case["feat"]), dev(case["shapes"]), dev(case["starts"]),
                                            dev(case["loc"])[:, :, :, :1], dev(case["weights"]))
    assert cuda_lib.hipad_dfa_forward_f32(None, None, None, None, None, None, 1, 1, 1, 1, 1, 1, 1, 1, None) == -1
    assert cuda_lib.hipad_dfa_backward_workspace_bytes(1, 6, 112200, 256, 4, 900, 13, 8) > 0


# ------------------------------------------------------------------------------------- feature_maps_format kernel
def _torch_format(fmaps):
    """The reference's own layout ops (ops/__init__.py:74-103): cat over levels, permute, flatten."""
    bs, cams = fmaps[0].shape[:2]
    return torch.cat([f.reshape(bs, cams, f.shape[2], -1) for f in fmaps], dim=-1).permute(0, 1, 3, 2).flatten(1, 2)


@pytest.mark.parametrize("shape", [
    (2, 6, 256, [(16, 28), (8, 14), (4, 7), (2, 4)]),
    (1, 3, 48, [(7, 9), (5, 3)]),                       # C and H*W not multiples of the 64-wide tile
    (1, 2, 130, [(65, 1), (1, 1), (3, 67)]),
])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_format_kernel_is_a_pure_copy(ops, shape, dtype):
    bs, cams, C, levels = shape
    g = torch.Generator(device="cuda").manual_seed(0)
    fmaps = [torch.randn((bs, cams, C, h, w), device="cuda", generator=g).to(dtype).requires_grad_(True) for h, w in levels]
    col, shapes_t, starts_t = ops.feature_maps_format(fmaps)
    ref = _torch_format([f.detach() for f in fmaps])
    assert col.dtype == dtype and torch.equal(col, ref)                       # bit-exact: layout only
    assert shapes_t.tolist() == [[list(hw) for hw in levels]] * cams
    # autograd of the format step = the same kernel run backwards
    gcol = torch.randn(col.shape, device="cuda", generator=g).to(dtype)
    col.backward(gcol)
    ref_in = [f.detach().clone().requires_grad_(True) for f in fmaps]
    _torch_format(ref_in).backward(gcol)
    for f, r in zip(fmaps, ref_in):
        assert torch.equal(f.grad, r.grad)
    # zero-copy inverse views still see the same data
    back = ops.feature_maps_format([col, shapes_t, starts_t], inverse=True)
    for f, v in zip(fmaps, back[0]):
        assert torch.equal(f.detach(), v)


def test_format_kernel_narrows_to_bf16(ops):
    g = torch.Generator(device="cuda").manual_seed(1)
    fmaps = [torch.randn((1, 6, 256, h, w), device="cuda", generator=g) for h, w in [(16, 28), (8, 14)]]
    col = ops.format_feature_levels(fmaps, out_dtype=torch.bfloat16)
    assert col.dtype == torch.bfloat16 and torch.equal(col, _torch_format(fmaps).bfloat16())


def test_format_feeds_the_op_end_to_end(ops, oracle_mod):
    """levels -> format kernel -> aggregation -> backward through both: gradients reach the NCHW pyramid."""
    case = small_case("c256_g8_l4", seed=5)
    bs, cams, lv, C = 2, 6, [(16, 28), (8, 14), (4, 7), (2, 4)], 256
    feat = torch.as_tensor(case["feat"]).cuda()                              # [bs, F, C]
    per_cam = sum(h * w for h, w in lv)
    blocks = feat.view(bs, cams, per_cam, C).split([h * w for h, w in lv], dim=2)
    fmaps = [b.permute(0, 1, 3, 2).reshape(bs, cams, C, h, w).contiguous().requires_grad_(True)
             for b, (h, w) in zip(blocks, lv)]
    fm = ops.feature_maps_format(fmaps)
    assert torch.equal(fm[0], feat)
    out = ops.deformable_aggregation_function(*fm, dev(case["loc"]), dev(case["weights"]))
    out.backward(dev(case["grad_out"]))
    r_feat, _, _ = oracle_mod.backward(case["feat"], case["shapes"], case["starts"], case["loc"], case["weights"],
                                       case["grad_out"])
    got = torch.cat([f.grad.reshape(bs, cams, C, -1) for f in fmaps], dim=-1).permute(0, 1, 3, 2).flatten(1, 2)
    assert rel_err(got.cpu().numpy(), r_feat) <= FP32_TOL


# ------------------------------------------------------------------------------------- shared feature-gradient buffer
@pytest.mark.parametrize("bf16", [False, True])
def test_shared_feature_gradient_matches_per_call_gradients(ops, bf16):
    """Three aggregation calls on one feature tensor: with share_feature_gradient() their feature gradients are
    accumulated in ONE buffer by the kernels (hipad_dfa_backward_accumulate_*); the result must equal autograd's sum of
    the three per-call dense gradients, and the other gradients must be bit-identical."""
    names = ["c256_g8_l4", "map_like_slices", "plan_like_slices"]
    cases = [small_case(n, seed=20 + i) for i, n in enumerate(names)]
    base = cases[0]
    cases[1] = H.make_case(21, 2, 6, [(16, 28), (8, 14), (4, 7), (2, 4)], 256, 8, 10, 300)   # same bs / feature geometry
    cases[2] = H.make_case(22, 2, 6, [(16, 28), (8, 14), (4, 7), (2, 4)], 256, 8, 48, 90)
    dt = torch.bfloat16 if bf16 else None

    def run(shared):
        feat = dev(base["feat"], dt).requires_grad_(True)
        f = ops.share_feature_gradient(feat) if shared else feat
        leaves, outs = [], []
        for c in cases:
            loc, w = dev(c["loc"]).requires_grad_(True), dev(c["weights"]).requires_grad_(True)
            outs.append(ops.deformable_aggregation_function(f, dev(base["shapes"]).long(), dev(base["starts"]).long(), loc, w))
            leaves.append((loc, w))
        extra = (f.float() * 0.5).sum()                      # a consumer that is not an aggregation call
        torch.autograd.backward(outs + [extra], [dev(c["grad_out"]) for c in cases] + [torch.ones((), device="cuda")])
        torch.cuda.synchronize()
        return feat.grad.float().cpu().numpy(), [(l.grad.cpu().numpy(), w.grad.cpu().numpy()) for l, w in leaves]

    g_ref, leaves_ref = run(False)
    g_sh, leaves_sh = run(True)
    assert rel_err(g_sh, g_ref) <= (BF16_TOL if bf16 else FP32_TOL)
    for (l0, w0), (l1, w1) in zip(leaves_ref, leaves_sh):
        assert np.array_equal(l0, l1) and np.array_equal(w0, w1)
    g_sh2, _ = run(True)
    assert np.array_equal(g_sh, g_sh2)                       # deterministic


def test_backward_accumulate_abi_adds_to_the_buffer(ops, cuda_lib, oracle_mod):
    case = small_case("c256_g8_l4", seed=31)
    feat, loc, w, go = dev(case["feat"]), dev(case["loc"]), dev(case["weights"]), dev(case["grad_out"])
    sh, st = dev(case["shapes"]).int(), dev(case["starts"]).int()
    bs, F, C = feat.shape
    A, P, cams = loc.shape[1:4]
    dims = (bs, cams, F, C, sh.shape[1], A, P, w.shape[-1])
    nb = cuda_lib.hipad_dfa_backward_workspace_bytes(*dims)
    ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
    g_feat = torch.full_like(feat, 0.25)
    g_loc, g_w = torch.empty_like(loc), torch.empty_like(w)
    rc = cuda_lib.hipad_dfa_backward_accumulate_f32(feat.data_ptr(), sh.data_ptr(), st.data_ptr(), loc.data_ptr(), w.data_ptr(),
                                                    go.data_ptr(), g_feat.data_ptr(), g_loc.data_ptr(), g_w.data_ptr(), *dims,
                                                    ws.data_ptr(), nb, torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    r_feat, r_loc, r_w = oracle_mod.backward(case["feat"], case["shapes"], case["starts"], case["loc"], case["weights"],
                                             case["grad_out"])
    assert rel_err(g_feat.cpu().numpy(), r_feat + 0.25) <= FP32_TOL
    assert rel_err(g_loc.cpu().numpy(), r_loc) <= FP32_TOL and rel_err(g_w.cpu().numpy(), r_w) <= FP32_TOL


# ------------------------------------------------------------------------------------- the other BASELINE.json configs
def _check_vs_oracle(ops, oracle_mod, case, bf16=False):
    out, g_feat, g_loc, g_w = run_bwd(ops, case, bf16=bf16)
    feat = case["feat"]
    if bf16:   # the oracle sees the same rounded feature values the kernel reads
        feat = torch.as_tensor(feat).bfloat16().float().numpy()
    ref = oracle_mod.forward(feat, case["shapes"], case["starts"], case["loc"], case["weights"])
    r_feat, r_loc, r_w = oracle_mod.backward(feat, case["shapes"], case["starts"], case["loc"], case["weights"],
                                             case["grad_out"])
    tol = BF16_TOL if bf16 else FP32_TOL
    assert rel_err(out, ref) <= FP32_TOL          # forward accumulates in fp32 either way
    assert rel_err(g_feat, r_feat) <= tol         # bf16: the gradient is stored as bf16
    assert rel_err(g_loc, r_loc) <= FP32_TOL and rel_err(g_w, r_w) <= FP32_TOL


def test_config0_256x704_det_vs_oracle(ops, oracle_mod):
    """BASELINE.json configs[0]: bs=1, 6 cams, 4 levels of a 256x704 input, 900 det anchors x 13 key points, 8 groups."""
    _check_vs_oracle(ops, oracle_mod, H.make_geo_case(40, "det", 1, H.LEVELS_256x704, (256, 704)))


def test_config4_hires_512x1408_vs_oracle_and_indices(ops, oracle_mod):
    """BASELINE.json configs[4]: 512x1408 input (F = 359 040 rows, 368 MB of fp32 features per sample)."""
    case = H.make_geo_case(41, "det", 1, H.LEVELS_512x1408, (512, 1408))
    _check_vs_oracle(ops, oracle_mod, case)
    idx = ops.sample_indices(dev(case["shapes"]).long(), dev(case["starts"]).long(), dev(case["loc"])).cpu().numpy()
    assert np.array_equal(idx, oracle_mod.indices(case["shapes"], case["starts"], case["loc"]))


@pytest.mark.parametrize("A,P,bf16", [(1800, 20, False), (2700, 7, True), (3600, 32, False)])
def test_config3_sweep_points(ops, oracle_mod, A, P, bf16):
    """BASELINE.json configs[3] (anchors 900-3600, key points 7-32, bf16 / fp32): mid-size points against the oracle,
    the largest one (A*P = 115 200 samples per camera) through size-independent properties."""
    case = H.make_geo_case(42 + A, "det", 1, H.LEVELS_352x640, (352, 640), A=A, P=P)
    if A * P <= 40000:
        _check_vs_oracle(ops, oracle_mod, case, bf16=bf16)
        return
    out1, g_feat1, g_loc1, g_w1 = run_bwd(ops, case, bf16=bf16)
    out2, g_feat2, g_loc2, g_w2 = run_bwd(ops, case, bf16=bf16)
    for a, b in ((out1, out2), (g_feat1, g_feat2), (g_loc1, g_loc2), (g_w1, g_w2)):
        assert np.array_equal(a, b), "results must be bitwise reproducible"
    lhs = float((case["grad_out"].astype(np.float64) * out1).sum())
    assert abs(float((g_w1.astype(np.float64) * case["weights"]).sum()) - lhs) <= 1e-4 * max(1.0, abs(lhs))
    assert abs(float((g_feat1.astype(np.float64) * case["feat"]).sum()) - lhs) <= 1e-4 * max(1.0, abs(lhs))
    # and a random subset of output rows against the oracle
    sel = np.random.default_rng(0).choice(A, 64, replace=False)
    sub = dict(case, loc=case["loc"][:, sel], weights=case["weights"][:, sel], grad_out=case["grad_out"][:, sel])
    assert rel_err(out1[:, sel], oracle_mod.forward(sub["feat"], sub["shapes"], sub["starts"], sub["loc"], sub["weights"])) <= FP32_TOL


def test_pile_up_on_one_pixel_goes_through_many_parts(ops, oracle_mod):
    """Every sample of every anchor hits the same few pixels: feature rows with thousands of contributions are split
    into 64-contribution parts whose partial sums are combined in part order (dfa_gfeat_reduce_kernel)."""
    case = small_case("c256_g8_l4", seed=50)
    bs, A, P, cams, _ = case["loc"].shape
    rng = np.random.default_rng(5)
    case["loc"] = (0.5 + 0.01 * rng.standard_normal(case["loc"].shape)).astype(np.float32)   # all visible, same spot
    _check_vs_oracle(ops, oracle_mod, case)
    out1 = run_bwd(ops, case)
    out2 = run_bwd(ops, case)
    assert all(np.array_equal(a, b) for a, b in zip(out1, out2))


def test_batch_of_eight(ops, oracle_mod):
    case = H.make_case(60, 8, 6, [(16, 28), (8, 14), (4, 7), (2, 4)], 256, 8, 23, 13)
    _check_vs_oracle(ops, oracle_mod, case)


def test_partial_slot_exhaustion_falls_back_to_single_warp_rows(ops, oracle_mod, monkeypatch):
    """Rows with more than 64 contributions normally go through partial-sum slots; when the pool is exhausted
    (forced here with HIPAD_DFA_PARTIAL_CAP) such a row is summed by one warp instead, with the same result."""
    case = small_case("c256_g8_l4", seed=51)
    case["loc"] = (0.5 + 0.02 * np.random.default_rng(6).standard_normal(case["loc"].shape)).astype(np.float32)
    ref = run_bwd(ops, case)
    monkeypatch.setenv("HIPAD_DFA_PARTIAL_CAP", "5")
    got = run_bwd(ops, case)
    monkeypatch.delenv("HIPAD_DFA_PARTIAL_CAP")
    _check_vs_oracle(ops, oracle_mod, case)
    for a, b in zip(ref, got):     # a different summation split for the overflowing rows: equal to fp32 accuracy
        assert rel_err(b, a) <= FP32_TOL


def test_backward_under_graph_capture_with_forked_sort_chain(ops, cuda_lib, oracle_mod):
    """The backward forks its compaction/sort/classify chain onto a helper stream.  Captured in a CUDA graph (helper
    stream created by the eager warm-up call, so the capture contains the fork and the join) it must reproduce the eager
    result bit for bit; captured on a stream the library has never seen it stays serial and must do the same."""
    case = small_case("c256_g8_l4", seed=70)
    feat, loc, w, go = dev(case["feat"]), dev(case["loc"]), dev(case["weights"]), dev(case["grad_out"])
    sh, st = dev(case["shapes"]).int(), dev(case["starts"]).int()
    bs, F, C = feat.shape
    A, P, cams = loc.shape[1:4]
    dims = (bs, cams, F, C, sh.shape[1], A, P, w.shape[-1])
    nb = cuda_lib.hipad_dfa_backward_workspace_bytes(*dims)
    ws = torch.empty(nb, dtype=torch.uint8, device="cuda")

    def call(stream, g_feat, g_loc, g_w):
        rc = cuda_lib.hipad_dfa_backward_f32(feat.data_ptr(), sh.data_ptr(), st.data_ptr(), loc.data_ptr(), w.data_ptr(),
                                             go.data_ptr(), g_feat.data_ptr(), g_loc.data_ptr(), g_w.data_ptr(), *dims,
                                             ws.data_ptr(), nb, stream.cuda_stream)
        assert rc == 0

    def buffers():
        return torch.full_like(feat, 7.0), torch.full_like(loc, 7.0), torch.full_like(w, 7.0)

    torch.cuda.synchronize()
    results = []
    for warm in (True, False):
        s = torch.cuda.Stream()
        bufs = buffers()
        with torch.cuda.stream(s):
            if warm:
                call(s, *bufs)                      # eager: creates the helper stream of `s`
                s.synchronize()
                for b in bufs:
                    b.fill_(7.0)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                call(torch.cuda.current_stream(), *bufs)
        g.replay()
        torch.cuda.synchronize()
        results.append([b.cpu().numpy() for b in bufs])
    eager = buffers()
    call(torch.cuda.current_stream(), *eager)
    torch.cuda.synchronize()
    r_feat, r_loc, r_w = oracle_mod.backward(case["feat"], case["shapes"], case["starts"], case["loc"], case["weights"],
                                             case["grad_out"])
    for got in results:
        for a, b in zip(got, eager):
            assert np.array_equal(a, b.cpu().numpy())
    assert rel_err(results[0][0], r_feat) <= FP32_TOL and rel_err(results[0][1], r_loc) <= FP32_TOL
