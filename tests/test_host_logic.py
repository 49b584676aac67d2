"""CPU tests of the host-side mirror of the reference interface (no GPU, no CUDA calls):
feature_maps_format, the key-point generators, weight/projection glue and the module's explicit
torch branch, all against fixtures produced by the unmodified reference (tests/golden)."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def build_module(d, use_deformable_func):
    import hipad_b200
    kind = str(d["kind"])
    embed = d["instance_feature"].shape[-1]
    L = len([k for k in d.files if k.startswith("fmap")])
    if kind == "det":
        kps = dict(type="SparseBox3DKeyPointsGenerator", num_learnable_pts=6,
                   fix_scale=[[0, 0, 0], [0.45, 0, 0], [-0.45, 0, 0], [0, 0.45, 0],
                              [0, -0.45, 0], [0, 0, 0.45], [0, 0, -0.45]])
    else:
        kps = dict(type="SparsePoint3DKeyPointsGenerator", embed_dims=embed, num_sample=6,
                   num_learnable_pts=3, fix_height=(0, 0.5, -0.5, 1, -1), ground_height=-1.84023)
    m = hipad_b200.DeformableFeatureAggregation(
        embed_dims=embed, num_groups=int(d["num_groups"]), num_levels=L, num_cams=6, attn_drop=0.15,
        use_deformable_func=use_deformable_func, use_camera_embed=True, residual_mode="cat", kps_generator=kps)
    sd = {k[3:]: torch.tensor(d[k]) for k in d.files if k.startswith("sd.")}
    m.load_state_dict(sd, strict=True)       # same state-dict keys as the reference module
    return m.eval(), L


@pytest.mark.parametrize("name", ["module_det", "module_plan", "module_det_daf", "module_det_c256", "module_plan_c256"])
def test_feature_maps_format_matches_reference(name):
    import hipad_b200
    d = np.load(os.path.join(GOLD, name + ".npz"))
    L = len([k for k in d.files if k.startswith("fmap")])
    fmaps = [torch.tensor(d[f"fmap{i}"]).float() for i in range(L)]      # (the C=256 fixtures store fp16-exact values)
    col, shapes, starts = hipad_b200.feature_maps_format(fmaps)
    assert list(col.shape) == d["col_feats_shape"].tolist()
    assert shapes.dtype == torch.int64 and starts.dtype == torch.int64
    assert np.array_equal(shapes.numpy(), d["spatial_shape"])
    assert np.array_equal(starts.numpy(), d["scale_start_index"])
    assert abs(col.double().sum().item() - float(d["col_feats_checksum"])) <= 1e-6 * max(1, abs(float(d["col_feats_checksum"])))
    back = hipad_b200.feature_maps_format([col, shapes, starts], inverse=True)
    assert len(back) == 1 and all(torch.equal(a, b) for a, b in zip(back[0], fmaps))
    # inverse also works on tensors that lost the host-side cache (reference behaviour: reads the device tables)
    back2 = hipad_b200.feature_maps_format([col, shapes.clone(), starts.clone()], inverse=True)
    assert all(torch.equal(a, b) for a, b in zip(back2[0], fmaps))


def test_feature_maps_format_camera_groups():
    import hipad_b200
    g = torch.Generator().manual_seed(0)
    grp_a = [torch.randn(2, 4, 8, 6, 10, generator=g), torch.randn(2, 4, 8, 3, 5, generator=g)]
    grp_b = [torch.randn(2, 2, 8, 4, 6, generator=g), torch.randn(2, 2, 8, 2, 3, generator=g)]
    col, shapes, starts = hipad_b200.feature_maps_format([grp_a, grp_b])
    assert col.shape == (2, 4 * 75 + 2 * 30, 8) and shapes.shape == (6, 2, 2)
    # rows of later groups start after every row of earlier groups
    assert starts[4, 0].item() == 4 * 75 and starts[5, 1].item() == 4 * 75 + 30 + 24
    back = hipad_b200.feature_maps_format([col, shapes, starts], inverse=True)
    assert len(back) == 2
    assert all(torch.equal(a, b) for a, b in zip(back[0], grp_a))
    assert all(torch.equal(a, b) for a, b in zip(back[1], grp_b))


@pytest.mark.parametrize("name", ["module_det", "module_plan", "module_det_daf", "module_det_c256", "module_plan_c256"])
def test_key_points_weights_and_projection_match_reference(name):
    d = np.load(os.path.join(GOLD, name + ".npz"))
    m, L = build_module(d, use_deformable_func=False)
    inst, anchor, emb = (torch.tensor(d[k]) for k in ("instance_feature", "anchor", "anchor_embed"))
    metas = dict(projection_mat=torch.tensor(d["projection_mat"]), image_wh=torch.tensor(d["image_wh"]))
    with torch.no_grad():
        kp = m.kps_generator(anchor, emb, inst)
        w = m._get_weights(inst, emb, metas)
        p2d = m.project_points(kp, metas["projection_mat"], metas["image_wh"])
    assert rel_err(kp.numpy(), d["key_points"]) <= 1e-6
    assert rel_err(w.numpy(), d["weights"]) <= 1e-5
    assert rel_err(p2d.numpy(), d["points_2d"]) <= 1e-5


@pytest.mark.parametrize("name", ["module_det", "module_det_c256", "module_plan_c256"])
def test_module_torch_branch_matches_reference_module(name):
    """use_deformable_func=False is the reference's own pure-torch branch, kept as an explicit opt-in.  The C=256
    fixtures are the shipped layout (G=8, 4 levels) run through the unmodified reference's torch branch, det and plan."""
    d = np.load(os.path.join(GOLD, name + ".npz"))
    assert not bool(d["via_oracle_daf"])
    m, L = build_module(d, use_deformable_func=False)
    fmaps = [torch.tensor(d[f"fmap{i}"]).float() for i in range(L)]
    metas = dict(projection_mat=torch.tensor(d["projection_mat"]), image_wh=torch.tensor(d["image_wh"]))
    with torch.no_grad():
        out = m(torch.tensor(d["instance_feature"]), torch.tensor(d["anchor"]), torch.tensor(d["anchor_embed"]),
                fmaps, metas)
    assert rel_err(out.numpy(), d["out"]) <= 1e-5


def test_module_cuda_branch_has_no_cpu_fallback():
    d = np.load(os.path.join(GOLD, "module_det.npz"))
    m, L = build_module(d, use_deformable_func=True)
    import hipad_b200
    fm = hipad_b200.feature_maps_format([torch.tensor(d[f"fmap{i}"]) for i in range(L)])
    metas = dict(projection_mat=torch.tensor(d["projection_mat"]), image_wh=torch.tensor(d["image_wh"]))
    with pytest.raises(RuntimeError, match="no CPU path"), torch.no_grad():
        m(torch.tensor(d["instance_feature"]), torch.tensor(d["anchor"]), torch.tensor(d["anchor_embed"]), fm, metas)


def test_state_dict_keys_match_reference_module():
    d = np.load(os.path.join(GOLD, "module_det.npz"))
    m, _ = build_module(d, use_deformable_func=True)
    ref_keys = sorted(k[3:] for k in d.files if k.startswith("sd."))
    assert sorted(m.state_dict().keys()) == ref_keys
    assert {"kps_generator.fix_scale", "kps_generator.learnable_fc.weight", "camera_encoder.0.weight",
            "camera_encoder.5.bias", "weights_fc.weight", "output_proj.weight"} <= set(ref_keys)


def test_temporal_key_points():
    import hipad_b200
    g = hipad_b200.SparseBox3DKeyPointsGenerator(embed_dims=16, num_learnable_pts=2,
                                                 fix_scale=[[0, 0, 0], [0.45, 0, 0]])
    anchor = torch.randn(2, 5, 11)
    feat = torch.randn(2, 5, 16)
    T = torch.eye(4).repeat(2, 1, 1)
    T[:, 0, 3] = 1.5
    kp, temp = g(anchor, feat, [T], torch.tensor([1.0, 1.0]), [torch.tensor([0.5, 0.5])])
    expect = kp - (anchor[..., 8:] * 0.5)[:, :, None] + torch.tensor([1.5, 0.0, 0.0])
    assert kp.shape == (2, 5, 4, 3) and torch.allclose(temp[0], expect, atol=1e-5)


def test_share_feature_gradient_is_a_no_op_off_the_gpu():
    """The shared-gradient wiring only exists for CUDA tensors that require grad; anything else passes through."""
    import hipad_b200
    x = torch.randn(2, 5, 8, requires_grad=True)
    assert hipad_b200.share_feature_gradient(x) is x
    y = torch.randn(2, 5, 8)
    assert hipad_b200.share_feature_gradient(y) is y


def test_reference_arm_prints_one_json_line():
    """bench.py --impl reference: the reference's CPU torch path, exactly one JSON line on stdout, the contract keys."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1", "--ref-max-steps", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GB/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    # both arms print the same `config` (the driver compares them)
    sys.path.insert(0, root)
    import bench
    assert d["config"] == bench.shared_config(1, "f32")


def test_static_feature_maps_refresh_policy():
    """graph_inference reads the feature maps from one static buffer per shape: refreshed for a new tensor object (a new
    frame) and for an in-place update of the same tensor, left alone when the same triple comes back (host logic only)."""
    from hipad_b200 import blocks
    blocks._STATIC_MAPS.clear()
    col = torch.arange(24, dtype=torch.float32).reshape(1, 6, 4)
    shapes = torch.tensor([[[2, 3]]], dtype=torch.int64)
    starts = torch.tensor([[0]], dtype=torch.int64)
    s1 = blocks._static_feature_maps([col, shapes, starts])
    assert torch.equal(s1[0], col) and s1[1].dtype == torch.int32 and torch.equal(s1[1].long(), shapes)
    s1[0].fill_(-1.0)                                   # scribble: a second call with the SAME triple must not copy again
    s2 = blocks._static_feature_maps([col, shapes, starts])
    assert s2[0] is s1[0] and float(s2[0].sum()) == -24.0
    col.add_(1.0)                                       # in-place update bumps the version: refreshed
    s3 = blocks._static_feature_maps([col, shapes, starts])
    assert s3[0] is s1[0] and torch.equal(s3[0], col)
    col2 = col.clone() * 2                              # a new frame: new tensor object, same shape -> same buffer, new content
    s4 = blocks._static_feature_maps([col2, shapes, starts])
    assert s4[0] is s1[0] and torch.equal(s4[0], col2)
    other = blocks._static_feature_maps([torch.zeros(1, 8, 4), shapes, starts])      # another shape: another buffer
    assert other[0] is not s1[0]
    blocks._STATIC_MAPS.clear()
