"""GPU tests of the grouped launches and of the parity-breadth cases VERDICT round 1 asked for:
full-size map / plan / ego calls against the C oracle and against the reference's own CUDA op, fp32 and bf16.

A group = the aggregation calls of one decoder layer (they read the same feature maps): one forward launch, one backward
chain, one feature gradient.  Checked against the per-call oracle results (outputs, location / weight gradients per
call, feature gradient = sum over calls), bitwise run-to-run, and through the public Python API.
"""
import numpy as np
import pytest
import torch

import helpers as H

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-5
BF16_TOL = 2e-2
SMALL_LV = [(16, 28), (8, 14), (4, 7), (2, 4)]


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def dev(x, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(x)).cuda()
    return t if dtype is None else t.to(dtype)


@pytest.fixture(scope="module")
def ops(cuda_lib):
    assert torch.cuda.is_available()
    import hipad_b200
    return hipad_b200.ops


def _reference_ext():
    from oracle import build_ref
    if not build_ref.available():
        pytest.skip("oracle/_ref not built")
    return build_ref.load()


def make_group(seed, bs, level_hw, final_hw, shapes_ap, C=256, G=8, geo=True):
    """calls of one layer on ONE feature tensor: list of dicts (loc, weights, grad_out) + shared feat/shapes/starts."""
    rng = np.random.default_rng(seed)
    shapes, starts, F = H.level_tables(level_hw, 6)
    feat = rng.standard_normal((bs, F, C), dtype=np.float32)
    calls = []
    for i, (kind, A, P) in enumerate(shapes_ap):
        if geo:
            c = H.make_geo_case(seed * 100 + i, "det" if kind == "ego" else kind, bs, level_hw, final_hw, C=C, G=G, A=A, P=P,
                                with_feat=False)
            loc, w = c["loc"], c["weights"]
            if kind == "ego":          # visible to no camera
                loc = np.full_like(loc, -0.5)
        else:
            c = H.make_case(seed * 100 + i, bs, 6, level_hw, C, G, A, P)
            loc, w = c["loc"], c["weights"]
        calls.append(dict(loc=loc, weights=w, grad_out=rng.standard_normal((bs, A, C), dtype=np.float32)))
    return dict(feat=feat, shapes=shapes, starts=starts, calls=calls)


def run_group(ops, g, bf16=False, shared=False, grouped=True):
    feat = dev(g["feat"], torch.bfloat16 if bf16 else None).requires_grad_(True)
    f = ops.share_feature_gradient(feat) if shared else feat
    sh, st = dev(g["shapes"]).long(), dev(g["starts"]).long()
    leaves = [(dev(c["loc"]).requires_grad_(True), dev(c["weights"]).requires_grad_(True)) for c in g["calls"]]
    if grouped:
        outs = ops.deformable_aggregation_group(f, sh, st, leaves)
    else:
        outs = [ops.deformable_aggregation_function(f, sh, st, l, w) for l, w in leaves]
    torch.autograd.backward(outs, [dev(c["grad_out"]) for c in g["calls"]])
    torch.cuda.synchronize()
    return ([o.detach().cpu().numpy() for o in outs], feat.grad.float().cpu().numpy(),
            [(l.grad.cpu().numpy(), w.grad.cpu().numpy()) for l, w in leaves])


def oracle_group(oracle_mod, g, feat=None):
    feat = g["feat"] if feat is None else feat
    outs, leaves, g_feat = [], [], np.zeros(g["feat"].shape, np.float64)
    for c in g["calls"]:
        outs.append(oracle_mod.forward(feat, g["shapes"], g["starts"], c["loc"], c["weights"]))
        gf, gl, gw = oracle_mod.backward(feat, g["shapes"], g["starts"], c["loc"], c["weights"], c["grad_out"])
        g_feat += gf
        leaves.append((gl, gw))
    return outs, g_feat, leaves


SMALL_GROUP = [("det", 40, 13), ("map", 10, 300), ("plan", 48, 90), ("ego", 1, 13)]


@pytest.mark.parametrize("bs", [1, 2])
@pytest.mark.parametrize("geo", [True, False])
def test_group_matches_oracle_small(ops, oracle_mod, bs, geo):
    g = make_group(3, bs, SMALL_LV, (64, 112), SMALL_GROUP, geo=geo)
    outs, g_feat, leaves = run_group(ops, g)
    r_outs, r_feat, r_leaves = oracle_group(oracle_mod, g)
    for o, r in zip(outs, r_outs):
        assert o.shape == r.shape and rel_err(o, r) <= FP32_TOL
    assert rel_err(g_feat, r_feat) <= FP32_TOL
    for (gl, gw), (rl, rw), c in zip(leaves, r_leaves, g["calls"]):
        assert rel_err(gl, rl) <= FP32_TOL and rel_err(gw, rw) <= FP32_TOL
        vis = ((c["loc"] > 0) & (c["loc"] < 1)).all(-1)
        assert not gl[~vis].any() and not gw[~vis].any()
    assert not g_feat[r_feat == 0].any()          # untouched rows are exact zeros


def test_group_equals_per_call_results_and_is_reproducible(ops):
    g = make_group(4, 2, SMALL_LV, (64, 112), SMALL_GROUP)
    a = run_group(ops, g, grouped=True)
    b = run_group(ops, g, grouped=True)
    c = run_group(ops, g, grouped=False)
    for x, y in zip(a[0], b[0]):
        assert np.array_equal(x, y)
    assert np.array_equal(a[1], b[1])
    for (l0, w0), (l1, w1) in zip(a[2], b[2]):
        assert np.array_equal(l0, l1) and np.array_equal(w0, w1)
    # grouped vs one call at a time: the same kernels on the same work units -> identical outputs and per-call
    # gradients; the feature gradient differs only in fp32 summation order (union of the calls vs autograd's sum)
    for x, y in zip(a[0], c[0]):
        assert np.array_equal(x, y)
    for (l0, w0), (l1, w1) in zip(a[2], c[2]):
        assert np.array_equal(l0, l1) and np.array_equal(w0, w1)
    assert rel_err(a[1], c[1]) <= FP32_TOL


@pytest.mark.parametrize("bf16", [False, True])
def test_group_stage2_layer_full_size(ops, oracle_mod, bf16):
    """One stage-2 decoder layer at 352x640: det 900x13 + map 100x300 + plan 480x90 + ego 1x13 in one group."""
    g = make_group(5, 1, H.LEVELS_352x640, (352, 640), [("det", 900, 13), ("map", 100, 300), ("plan", 480, 90), ("ego", 1, 13)])
    outs, g_feat, leaves = run_group(ops, g, bf16=bf16)
    feat = torch.as_tensor(g["feat"]).bfloat16().float().numpy() if bf16 else g["feat"]
    r_outs, r_feat, r_leaves = oracle_group(oracle_mod, g, feat)
    for o, r in zip(outs, r_outs):
        assert rel_err(o, r) <= FP32_TOL
    assert rel_err(g_feat, r_feat) <= (BF16_TOL if bf16 else FP32_TOL)
    for (gl, gw), (rl, rw) in zip(leaves, r_leaves):
        assert rel_err(gl, rl) <= FP32_TOL and rel_err(gw, rw) <= FP32_TOL
    assert not outs[3].any() and not leaves[3][0].any() and not leaves[3][1].any()      # ego: nothing visible


@pytest.mark.parametrize("bf16", [False, True])
def test_group_with_shared_step_buffer(ops, oracle_mod, bf16):
    """Two groups (two decoder layers) accumulating into ONE fp32 step buffer via share_feature_gradient."""
    g1 = make_group(6, 2, SMALL_LV, (64, 112), SMALL_GROUP)
    g2 = make_group(7, 2, SMALL_LV, (64, 112), SMALL_GROUP)
    g2["feat"] = g1["feat"]
    feat = dev(g1["feat"], torch.bfloat16 if bf16 else None).requires_grad_(True)
    f = ops.share_feature_gradient(feat)
    sh, st = dev(g1["shapes"]).long(), dev(g1["starts"]).long()
    outs, gos = [], []
    for g in (g1, g2):
        leaves = [(dev(c["loc"]).requires_grad_(True), dev(c["weights"]).requires_grad_(True)) for c in g["calls"]]
        outs += ops.deformable_aggregation_group(f, sh, st, leaves)
        gos += [dev(c["grad_out"]) for c in g["calls"]]
    torch.autograd.backward(outs, gos)
    torch.cuda.synchronize()
    fr = torch.as_tensor(g1["feat"]).bfloat16().float().numpy() if bf16 else g1["feat"]
    ref = oracle_group(oracle_mod, g1, fr)[1] + oracle_group(oracle_mod, g2, fr)[1]
    # fp32 accumulation across all 8 calls, ONE rounding to bf16 at the end
    assert rel_err(feat.grad.float().cpu().numpy(), ref) <= (4e-3 if bf16 else FP32_TOL)


def test_shared_buffer_does_not_leak_between_backward_passes(ops, oracle_mod):
    """a pass that never reaches the sink node must not leave a stale buffer behind."""
    g = make_group(8, 1, SMALL_LV, (64, 112), SMALL_GROUP[:2])
    feat = dev(g["feat"]).requires_grad_(True)
    f = ops.share_feature_gradient(feat)
    sh, st = dev(g["shapes"]).long(), dev(g["starts"]).long()
    leaves = [(dev(c["loc"]).requires_grad_(True), dev(c["weights"]).requires_grad_(True)) for c in g["calls"]]
    outs = [ops.deformable_aggregation_function(f, sh, st, l, w) for l, w in leaves]
    gos = [dev(c["grad_out"]) for c in g["calls"]]
    # pass 1: gradients w.r.t. the locations only -- the sink never runs
    torch.autograd.grad(outs, [l for l, _ in leaves], gos, retain_graph=True)
    # pass 2 and 3: full backward twice on the retained graph
    torch.autograd.backward(outs, gos, retain_graph=True)
    first = feat.grad.clone()
    feat.grad = None
    torch.autograd.backward(outs, gos)
    torch.cuda.synchronize()
    ref = oracle_group(oracle_mod, g)[1]
    assert rel_err(first.cpu().numpy(), ref) <= FP32_TOL
    assert torch.equal(first, feat.grad)


def test_group_abi_rejects_bad_arguments(ops, cuda_lib):
    from hipad_b200 import _lib
    import ctypes
    g = make_group(9, 1, SMALL_LV, (64, 112), SMALL_GROUP[:1])
    feat, loc, w = dev(g["feat"]), dev(g["calls"][0]["loc"]), dev(g["calls"][0]["weights"])
    sh, st = dev(g["shapes"]).int(), dev(g["starts"]).int()
    out = torch.empty((1, 40, 256), device="cuda")
    tab = _lib.call_table([(loc.data_ptr(), w.data_ptr(), None, None, 40, 13)])
    tp = ctypes.cast(tab, ctypes.c_void_p)
    s = torch.cuda.current_stream().cuda_stream
    F = feat.shape[1]
    assert cuda_lib.hipad_dfa_group_forward(0, out.data_ptr(), feat.data_ptr(), sh.data_ptr(), st.data_ptr(), tp, 0,
                                            1, 6, F, 256, 4, 8, None, 0, s) == -1
    assert cuda_lib.hipad_dfa_group_forward(0, None, feat.data_ptr(), sh.data_ptr(), st.data_ptr(), tp, 1,
                                            1, 6, F, 256, 4, 8, None, 0, s) == -1
    # 5 levels: outside the grouped kernels' family -> UNSUPPORTED (callers go call by call)
    assert cuda_lib.hipad_dfa_group_forward(0, out.data_ptr(), feat.data_ptr(), sh.data_ptr(), st.data_ptr(), tp, 1,
                                            1, 6, F, 256, 5, 8, None, 0, s) == -2
    # backward without gradient buffers for the call
    assert cuda_lib.hipad_dfa_group_backward(0, 0, feat.data_ptr(), sh.data_ptr(), st.data_ptr(), tp, 1, out.data_ptr(),
                                             None, 1, 6, F, 256, 4, 8, None, 0, s) == -1


# ------------------------------------------------------------------------------------- parity breadth (VERDICT weak #1)
@pytest.mark.parametrize("kind,A,P", [("map", 100, 300), ("plan", 480, 90), ("ego", 1, 13)])
@pytest.mark.parametrize("bf16", [False, True])
def test_stage2_map_plan_ego_full_size_vs_oracle(ops, oracle_mod, kind, A, P, bf16):
    g = make_group(10, 1, H.LEVELS_352x640, (352, 640), [(kind, A, P)])
    outs, g_feat, leaves = run_group(ops, g, bf16=bf16, grouped=False)
    feat = torch.as_tensor(g["feat"]).bfloat16().float().numpy() if bf16 else g["feat"]
    r_outs, r_feat, r_leaves = oracle_group(oracle_mod, g, feat)
    assert rel_err(outs[0], r_outs[0]) <= FP32_TOL
    assert rel_err(g_feat, r_feat) <= (BF16_TOL if bf16 else FP32_TOL)
    assert rel_err(leaves[0][0], r_leaves[0][0]) <= FP32_TOL and rel_err(leaves[0][1], r_leaves[0][1]) <= FP32_TOL


@pytest.mark.parametrize("kind,A,P", [("map", 100, 300), ("plan", 480, 90), ("ego", 1, 13)])
def test_stage2_map_plan_ego_full_size_vs_reference_cuda_op(ops, kind, A, P):
    ext = _reference_ext()
    g = make_group(11, 1, H.LEVELS_352x640, (352, 640), [(kind, A, P)])
    c = g["calls"][0]
    feat, loc, w, go = dev(g["feat"]), dev(c["loc"]), dev(c["weights"]), dev(c["grad_out"])
    shapes, starts = dev(g["shapes"]), dev(g["starts"])
    ref_out = ext.deformable_aggregation_forward(feat, shapes, starts, loc, w)
    r_feat, r_loc, r_w = torch.zeros_like(feat), torch.zeros_like(loc), torch.zeros_like(w)
    ext.deformable_aggregation_backward(feat, shapes, starts, loc, w, go, r_feat, r_loc, r_w)
    outs, g_feat, leaves = run_group(ops, g, grouped=False)
    assert rel_err(outs[0], ref_out.cpu().numpy()) <= FP32_TOL
    assert rel_err(g_feat, r_feat.cpu().numpy()) <= FP32_TOL
    assert rel_err(leaves[0][0], r_loc.cpu().numpy()) <= FP32_TOL
    assert rel_err(leaves[0][1], r_w.cpu().numpy()) <= FP32_TOL


def test_hires_512x1408_layer_bs2(ops, oracle_mod):
    """BASELINE configs[4] geometry (the only HBM-resident case), stage-1 plan size, two samples, grouped."""
    g = make_group(12, 2, H.LEVELS_512x1408, (512, 1408), [("det", 900, 13), ("map", 100, 300), ("plan", 48, 90), ("ego", 1, 13)])
    outs, g_feat, leaves = run_group(ops, g)
    r_outs, r_feat, r_leaves = oracle_group(oracle_mod, g)
    for o, r in zip(outs, r_outs):
        assert rel_err(o, r) <= FP32_TOL
    assert rel_err(g_feat, r_feat) <= FP32_TOL
    for (gl, gw), (rl, rw) in zip(leaves, r_leaves):
        assert rel_err(gl, rl) <= FP32_TOL and rel_err(gw, rw) <= FP32_TOL


def test_reference_op_cells_at_near_integer_coordinates(ops):
    """the integer contract against the COMPILED reference op, not only our oracle.  The forward value
    is continuous across a cell boundary, the location gradient is not (it is the slope inside the cell the sample was
    floored into), so agreeing location gradients at coordinates within a few ulps of pixel centres / edges show that
    both implementations floor the same way (one fused multiply-add, SURVEY.md §7 "Bit-exact indices")."""
    ext = _reference_ext()
    lv = [(16, 28), (7, 13)]
    shapes, starts, F = H.level_tables(lv, 2)
    C, G, A, P = 32, 1, 2048, 2
    rng = np.random.default_rng(0)
    n = A * P * 2
    W = np.where(rng.random(n) < 0.5, 28, 13).astype(np.float32)
    Hh = np.where(W == 28, 16, 7).astype(np.float32)
    x = ((rng.integers(0, 28, n) % W + 0.5) / W).astype(np.float32)      # exactly a pixel centre at one level
    y = ((rng.integers(0, 16, n) % Hh + 0.5) / Hh).astype(np.float32)
    for _ in range(3):                                                    # and up to 3 ulps to either side
        step = rng.integers(-1, 2, n)
        x = np.where(step > 0, np.nextafter(x, np.float32(2)), np.where(step < 0, np.nextafter(x, np.float32(-2)), x))
        step = rng.integers(-1, 2, n)
        y = np.where(step > 0, np.nextafter(y, np.float32(2)), np.where(step < 0, np.nextafter(y, np.float32(-2)), y))
    loc = np.stack([x, y], -1).reshape(1, A, P, 2, 2).astype(np.float32)
    feat = rng.standard_normal((1, F, C), dtype=np.float32)
    w = rng.standard_normal((1, A, P, 2, 2, G), dtype=np.float32)
    go = rng.standard_normal((1, A, C), dtype=np.float32)
    f, l, ww, g = dev(feat), dev(loc), dev(w), dev(go)
    sh, st = dev(shapes), dev(starts)
    ref_out = ext.deformable_aggregation_forward(f, sh, st, l, ww)
    r_feat, r_loc, r_w = torch.zeros_like(f), torch.zeros_like(l), torch.zeros_like(ww)
    ext.deformable_aggregation_backward(f, sh, st, l, ww, g, r_feat, r_loc, r_w)
    fl = l.clone().requires_grad_(True)
    out = ops.deformable_aggregation_function(f, sh.long(), st.long(), fl, ww)
    out.backward(g)
    torch.cuda.synchronize()
    assert rel_err(out.detach().cpu().numpy(), ref_out.cpu().numpy()) <= FP32_TOL
    assert rel_err(fl.grad.cpu().numpy(), r_loc.cpu().numpy()) <= FP32_TOL


def test_misaligned_sub_buffers_are_rejected_not_faulted(ops, cuda_lib):
    """a g_w / location pointer that is only 4-byte aligned (a sub-buffer) must give a clean status."""
    case = H.make_case(40, 1, 6, SMALL_LV, 256, 8, 12, 5)
    feat, go = dev(case["feat"]), dev(case["grad_out"])
    sh, st = dev(case["shapes"]).int(), dev(case["starts"]).int()
    bs, F, C = feat.shape
    dims = (bs, 6, F, C, 4, 12, 5, 8)
    loc_buf = torch.zeros(case["loc"].size + 1, device="cuda"); loc_buf[1:] = dev(case["loc"]).reshape(-1)
    w_buf = torch.zeros(case["weights"].size + 1, device="cuda"); w_buf[1:] = dev(case["weights"]).reshape(-1)
    loc_ok, w_ok = dev(case["loc"]), dev(case["weights"])
    g_loc, g_w = torch.empty_like(loc_ok), torch.empty(w_ok.numel() + 1, device="cuda")
    g_feat = torch.empty_like(feat)
    nb = cuda_lib.hipad_dfa_backward_workspace_bytes(*dims)
    ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    out = torch.empty((bs, 12, C), device="cuda")
    # forward with a location pointer that is 4 bytes off an 8-byte boundary
    assert cuda_lib.hipad_dfa_forward_f32(out.data_ptr(), feat.data_ptr(), sh.data_ptr(), st.data_ptr(),
                                          loc_buf.data_ptr() + 4, w_ok.data_ptr(), *dims, s) == -1
    # backward with a weight-gradient pointer that is 4 bytes off a 16-byte boundary
    assert cuda_lib.hipad_dfa_backward_f32(feat.data_ptr(), sh.data_ptr(), st.data_ptr(), loc_ok.data_ptr(), w_ok.data_ptr(),
                                           go.data_ptr(), g_feat.data_ptr(), g_loc.data_ptr(), g_w.data_ptr() + 4, *dims,
                                           ws.data_ptr(), nb, s) == -1
    # weights that are only 4-byte aligned are legal: the scalar kernel family takes them
    out2 = torch.empty_like(out)
    assert cuda_lib.hipad_dfa_forward_f32(out2.data_ptr(), feat.data_ptr(), sh.data_ptr(), st.data_ptr(),
                                          loc_ok.data_ptr(), w_buf.data_ptr() + 4, *dims, s) == 0
    assert cuda_lib.hipad_dfa_forward_f32(out.data_ptr(), feat.data_ptr(), sh.data_ptr(), st.data_ptr(),
                                          loc_ok.data_ptr(), w_ok.data_ptr(), *dims, s) == 0
    torch.cuda.synchronize()
    assert rel_err(out2.cpu().numpy(), out.cpu().numpy()) <= FP32_TOL


def test_aggregate_layer_equals_module_by_module(ops):
    """hipad_b200.aggregate_layer (the one-edit replacement of sparse_onedecoder.py:867-887) vs calling each module."""
    import hipad_b200
    torch.manual_seed(0)
    kps_box = dict(type="SparseBox3DKeyPointsGenerator", num_learnable_pts=6,
                   fix_scale=[[0, 0, 0], [0.45, 0, 0], [-0.45, 0, 0], [0, 0.45, 0], [0, -0.45, 0], [0, 0, 0.45], [0, 0, -0.45]])
    kps_pts = dict(type="SparsePoint3DKeyPointsGenerator", embed_dims=256, num_sample=6, num_learnable_pts=3,
                   fix_height=(0, 0.5, -0.5, 1, -1), ground_height=-1.84023)
    mods = [hipad_b200.DeformableFeatureAggregation(embed_dims=256, num_groups=8, num_levels=4, num_cams=6, attn_drop=0.0,
                                                    use_deformable_func=True, use_camera_embed=True, residual_mode="cat",
                                                    kps_generator=k).cuda() for k in (kps_box, kps_pts)]
    for m in mods:
        torch.nn.init.normal_(m.weights_fc.weight, std=0.02)
    bs = 2
    levels = [torch.randn(bs, 6, 256, h, w, device="cuda") for h, w in SMALL_LV]
    fm = hipad_b200.feature_maps_format(levels)
    proj = torch.from_numpy(np.tile(H.projection_matrices((64, 112))[None], (bs, 1, 1, 1))).cuda()
    metas = dict(projection_mat=proj, image_wh=torch.tensor([[112.0, 64.0]], device="cuda").repeat(bs, 6, 1))
    rng = np.random.default_rng(0)
    box = torch.from_numpy(np.concatenate([rng.uniform(-20, 20, (bs, 30, 2)), rng.uniform(-2, 0, (bs, 30, 1)),
                                           rng.normal(0.5, 0.2, (bs, 30, 3)), np.tile([[[0.0, 1.0, 0, 0, 0]]], (bs, 30, 1))],
                                          -1).astype(np.float32)).cuda()
    line = torch.from_numpy(rng.uniform(-15, 30, (bs, 12, 12)).astype(np.float32)).cuda()
    calls = [(mods[0], torch.randn(bs, 30, 256, device="cuda"), box, torch.randn(bs, 30, 256, device="cuda")),
             (mods[1], torch.randn(bs, 12, 256, device="cuda"), line, torch.randn(bs, 12, 256, device="cuda"))]
    with torch.enable_grad():
        fm_g = [fm[0].clone().requires_grad_(True), fm[1], fm[2]]
        outs = hipad_b200.aggregate_layer(calls, fm_g, metas)
        sum(o.square().sum() for o in outs).backward()
        fm_s = [fm[0].clone().requires_grad_(True), fm[1], fm[2]]
        for m in mods:
            m.zero_grad()
        refs = [m(i, a, e, fm_s, metas) for m, i, a, e in calls]
        sum(o.square().sum() for o in refs).backward()
    for o, r in zip(outs, refs):
        assert rel_err(o.detach().cpu().numpy(), r.detach().cpu().numpy()) <= FP32_TOL
    assert rel_err(fm_g[0].grad.cpu().numpy(), fm_s[0].grad.cpu().numpy()) <= FP32_TOL


# ------------------------------------------------------------------------------------- weights producer (row f2)
def _reference_weights_chain(logits, cams, L, P, G, keep, drop_p):
    """blocks.py:196-212 + :147-158 in torch ops: softmax over cams*L*P per group, mask, permute to the op's layout."""
    bs, A = logits.shape[:2]
    w = logits.reshape(bs, A, -1, G).softmax(dim=-2).reshape(bs, A, cams, L, P, G)
    if drop_p > 0:
        w = (keep[:, :, :, None, :, None] * w) / (1 - drop_p)
    return w.permute(0, 1, 4, 2, 3, 5).contiguous()


@pytest.mark.parametrize("shape", [(2, 37, 6, 4, 13, 8), (1, 5, 6, 4, 300, 8), (2, 9, 3, 4, 90, 4), (1, 3, 2, 3, 7, 1)])
@pytest.mark.parametrize("drop_p", [0.0, 0.15])
def test_weights_producer_matches_torch_chain(ops, shape, drop_p):
    bs, A, cams, L, P, G = shape
    g = torch.Generator(device="cuda").manual_seed(1)
    logits = (torch.randn(bs, A, cams, L * P * G, device="cuda", generator=g) * 3).requires_grad_(True)
    keep = (torch.rand(bs, A, cams, P, device="cuda", generator=g) > drop_p).float()
    go = torch.randn(bs, A, P, cams, L, G, device="cuda", generator=g)
    w = ops.aggregation_weights(logits, cams, L, P, G, drop_p, keep_mask=keep if drop_p > 0 else None)
    w.backward(go)
    got_g = logits.grad.clone()
    logits.grad = None
    ref = _reference_weights_chain(logits, cams, L, P, G, keep, drop_p)
    ref.backward(go)
    assert w.shape == ref.shape
    assert rel_err(w.detach().cpu().numpy(), ref.detach().cpu().numpy()) <= 2e-6
    assert rel_err(got_g.cpu().numpy(), logits.grad.cpu().numpy()) <= FP32_TOL


def test_weights_producer_draws_its_mask_on_the_device(ops):
    """In-kernel attn-drop: whole (cam, point) columns are dropped with probability p, kept ones scaled by 1/(1-p),
    the same seed reproduces the same mask, forward and backward agree on it."""
    bs, A, cams, L, P, G, p = 2, 200, 6, 4, 13, 8, 0.15
    logits = torch.randn(bs, A, cams, L * P * G, device="cuda").requires_grad_(True)
    base = ops.aggregation_weights(logits.detach(), cams, L, P, G, 0.0)
    w1 = ops.aggregation_weights(logits, cams, L, P, G, p, seed=1234)
    w2 = ops.aggregation_weights(logits.detach(), cams, L, P, G, p, seed=1234)
    w3 = ops.aggregation_weights(logits.detach(), cams, L, P, G, p, seed=99)
    assert torch.equal(w1.detach(), w2) and not torch.equal(w2, w3)
    ratio = (w1.detach() / base)                                   # [bs, A, P, cams, L, G]: 0 or 1/(1-p), constant over (L, G)
    kept = ratio > 0.5
    assert torch.allclose(ratio[kept], torch.full_like(ratio[kept], 1 / (1 - p)), rtol=1e-5)
    assert (kept == kept[..., :1, :1]).all()
    frac = float(kept[..., 0, 0].float().mean())
    assert abs(frac - (1 - p)) < 0.02
    go = torch.randn_like(w1)
    w1.backward(go)
    keep = kept[..., 0, 0].permute(0, 1, 3, 2).float()             # [bs, A, cams, P]
    l2 = logits.detach().clone().requires_grad_(True)
    _reference_weights_chain(l2, cams, L, P, G, keep, p).backward(go)
    assert rel_err(logits.grad.cpu().numpy(), l2.grad.cpu().numpy()) <= FP32_TOL


def test_module_training_path_uses_the_fused_weights_producer(ops):
    """Module forward + backward in training mode (attn_drop off for comparability) vs the reference chain of torch ops."""
    import hipad_b200
    torch.manual_seed(0)
    kps = dict(type="SparseBox3DKeyPointsGenerator", num_learnable_pts=6,
               fix_scale=[[0, 0, 0], [0.45, 0, 0], [-0.45, 0, 0], [0, 0.45, 0], [0, -0.45, 0], [0, 0, 0.45], [0, 0, -0.45]])
    m = hipad_b200.DeformableFeatureAggregation(embed_dims=256, num_groups=8, num_levels=4, num_cams=6, attn_drop=0.0,
                                                use_deformable_func=True, use_camera_embed=True, residual_mode="cat",
                                                kps_generator=kps).cuda().train()
    torch.nn.init.normal_(m.weights_fc.weight, std=0.05)
    bs = 2
    fm = hipad_b200.feature_maps_format([torch.randn(bs, 6, 256, h, w, device="cuda") for h, w in SMALL_LV])
    proj = torch.from_numpy(np.tile(H.projection_matrices((64, 112))[None], (bs, 1, 1, 1))).cuda()
    metas = dict(projection_mat=proj, image_wh=torch.tensor([[112.0, 64.0]], device="cuda").repeat(bs, 6, 1))
    rng = np.random.default_rng(0)
    box = torch.from_numpy(np.concatenate([rng.uniform(-20, 20, (bs, 30, 2)), rng.uniform(-2, 0, (bs, 30, 1)),
                                           rng.normal(0.5, 0.2, (bs, 30, 3)), np.tile([[[0.0, 1.0, 0, 0, 0]]], (bs, 30, 1))],
                                          -1).astype(np.float32)).cuda()
    inst = torch.randn(bs, 30, 256, device="cuda", requires_grad=True)
    emb = torch.randn(bs, 30, 256, device="cuda")
    out = m(inst, box, emb, fm, metas)
    out.square().sum().backward()
    g1, gw1 = inst.grad.clone(), m.weights_fc.weight.grad.clone()
    inst.grad = None
    m.zero_grad()
    # reference chain: _get_weights (torch softmax) + permute + the same op
    kp = m.kps_generator(box, emb, inst)
    w = m._get_weights(inst, emb, metas).permute(0, 1, 4, 2, 3, 5).contiguous()
    p2d = m.project_points(kp, metas["projection_mat"], metas["image_wh"]).permute(0, 2, 3, 1, 4).reshape(bs, 30, 13, 6, 2)
    f = ops.deformable_aggregation_function(*fm, p2d, w)
    ref = torch.cat([m.output_proj(f), inst], dim=-1)
    ref.square().sum().backward()
    assert rel_err(out.detach().cpu().numpy(), ref.detach().cpu().numpy()) <= FP32_TOL
    assert rel_err(g1.cpu().numpy(), inst.grad.cpu().numpy()) <= 2e-5
    assert rel_err(gw1.cpu().numpy(), m.weights_fc.weight.grad.cpu().numpy()) <= 2e-5


# ------------------------------------------------------------------------------------- fused inference forward at full size
@pytest.mark.parametrize("kind,A,P", [("det", 900, 13), ("map", 100, 300), ("plan", 480, 90)])
@pytest.mark.parametrize("bf16", [False, True])
def test_fused_forward_stage2_shapes(ops, oracle_mod, kind, A, P, bf16):
    """the fused kernel (projection + group softmax + aggregation; for map rows a two-pass
    softmax over 57 600 logits split across a thread-block cluster) at the stage-2 shapes it is benchmarked on, against the
    reference chain in torch ops feeding the unfused op, and against the C oracle."""
    import hipad_b200
    case = H.make_geo_case(31, kind, 1, H.LEVELS_352x640, (352, 640), A=A, P=P)
    bs, cams, F, C, L, A_, P_, G = case["dims"]
    feat = dev(case["feat"], torch.bfloat16 if bf16 else None)
    fm = [feat, dev(case["shapes"]).long(), dev(case["starts"]).long()]
    logits = dev(case["logits"]).reshape(bs, A, cams, L * P * G) * 2.0
    kp, pm, wh = dev(case["key_points"]), dev(case["projection_mat"]), dev(case["image_wh"])
    out, loc = ops.fused_deformable_aggregation(fm, kp, pm, wh, logits, return_locations=True)
    p2d = hipad_b200.DeformableFeatureAggregation.project_points(kp, pm, wh).permute(0, 2, 3, 1, 4).contiguous()
    assert rel_err(loc.cpu().numpy(), p2d.cpu().numpy()) <= 1e-5
    assert float((((loc > 0) & (loc < 1)).all(-1) != ((p2d > 0) & (p2d < 1)).all(-1)).float().mean()) <= 1e-3
    w = logits.reshape(bs, A, -1, G).softmax(dim=-2).reshape(bs, A, cams, L, P, G).permute(0, 1, 4, 2, 3, 5).contiguous()
    ref = ops.deformable_aggregation_function(*fm, loc, w)           # same locations, torch softmax, unfused op
    assert rel_err(out.cpu().numpy(), ref.cpu().numpy()) <= FP32_TOL
    f_cpu = feat.float().cpu().numpy()
    cpu = oracle_mod.forward(f_cpu, case["shapes"], case["starts"], loc.cpu().numpy(), w.cpu().numpy())
    assert rel_err(out.cpu().numpy(), cpu) <= FP32_TOL
    # and the training-side producer gives the same weights as torch's softmax + permute
    w2 = ops.aggregation_weights(logits, cams, L, P, G)
    assert rel_err(w2.cpu().numpy(), w.cpu().numpy()) <= 2e-6
