"""Decoder harness: the UNMODIFIED reference ``SparseOneDecoder`` (vendored copy, harness/vendor.py) built through a
mmcv/mmdet stand-in (harness/mmcv_shim.py) with a chosen ``projects.mmdet3d_plugin.ops`` package plugged in.

TEST / BENCH INFRASTRUCTURE.  The decoder is the caller of the hot path, not part of it; it is
loaded only to (a) prove the drop-in claim of row a11 — the reference's own model code runs on ``hipad_b200.ops`` with
zero edits — and (b) measure BASELINE configs[1]/[2] ("HiP-AD decoder samples/s").

Injection point: ``models/blocks.py:21`` does ``from ..ops import
deformable_aggregation_function as DAF`` and ``models/{sparse_detector,ego/instance_bank,plan/instance_bank}.py``
import ``feature_maps_format`` from the same package, so registering a module under
``sys.modules['projects.mmdet3d_plugin.ops']`` before importing ``models`` swaps the op.

Variants of the op package:
  "reference"  the reference's own ops/*.py (vendored) over its own CUDA extension rebuilt for sm_100a
               (oracle/_ref/deformable_aggregation_ext.so, oracle/build_ref.py)
  "ours"       hipad_b200.ops (same decoder, same DeformableFeatureAggregation module class as the reference)
  "ours_module" hipad_b200.ops AND hipad_b200.DeformableFeatureAggregation registered in the ATTENTION registry
               under the reference's name (the fused inference path / grouped launches are then what runs)
  "ours_module_graph" the same with the module's CUDA-graph inference forward switched on (graph_inference)
  any module object: used as given (tests pass a torch emulation for CPU runs)
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

from . import mmcv_shim, vendor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CONFIG = "projects/configs/hipad_b2d_stage2.py"

def camera_matrices(hw):
    """projection_mat [6,4,4] (lidar -> augmented image pixels) and image_wh [6,2] for a final image of hw=(H,W):
    the Bench2Drive camera table with the agent's test-time resize + top crop (tests/helpers.py)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    proj = helpers.projection_matrices(hw)
    return proj.astype(np.float32), np.tile(np.array([[hw[1], hw[0]]], dtype=np.float32), (proj.shape[0], 1))


# ----------------------------------------------------------------------------------------------- loading
def _purge(prefix):
    for name in [n for n in sys.modules if n == prefix or n.startswith(prefix + ".")]:
        del sys.modules[name]


def _bare_package(name, path):
    m = types.ModuleType(name)
    m.__path__ = [path]
    sys.modules[name] = m
    return m


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def reference_ops_module(root):
    """The reference's own ``ops`` package (vendored .py files) bound to its CUDA extension from oracle/_ref."""
    from oracle import build_ref
    if not build_ref.available():
        raise RuntimeError("oracle/_ref/deformable_aggregation_ext.so is missing (oracle/build_ref.py builds it "
                           "where /root/reference exists)")
    ext = build_ref.load()
    pkg = "projects.mmdet3d_plugin.ops"
    sys.modules[pkg + ".deformable_aggregation_ext"] = ext
    spec = importlib.util.spec_from_file_location(
        pkg, os.path.join(root, "projects", "mmdet3d_plugin", "ops", "__init__.py"),
        submodule_search_locations=[os.path.join(root, "projects", "mmdet3d_plugin", "ops")])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[pkg] = mod
    spec.loader.exec_module(mod)
    return mod


class CountingOps(types.ModuleType):
    """Wraps an ops package: counts ``deformable_aggregation_function`` calls and records their (A, P) shapes."""

    def __init__(self, inner):
        super().__init__("projects.mmdet3d_plugin.ops")
        self._inner = inner
        self.calls = []
        for k in dir(inner):
            if not k.startswith("__"):
                setattr(self, k, getattr(inner, k))
        fn = inner.deformable_aggregation_function

        self.record = None          # set to a list to keep (inputs, output) of every call (in-situ parity checks)

        def counted(feature_maps, spatial_shape, scale_start_index, sampling_location, weights):
            self.calls.append(tuple(sampling_location.shape[1:3]))
            out = fn(feature_maps, spatial_shape, scale_start_index, sampling_location, weights)
            if self.record is not None:
                self.record.append((feature_maps, spatial_shape, scale_start_index, sampling_location, weights, out))
            return out
        self.deformable_aggregation_function = counted


def load_models(ops="ours"):
    """Import the vendored reference ``models`` package with `ops` plugged in.  Returns (models_module, registries,
    ops_module).  Every call re-imports the package, so decoders built from different calls hold different ops."""
    root = vendor.vendor()
    if root is None:
        raise RuntimeError("baseline/_ref/hipad is missing: run `python harness/vendor.py` where /root/reference exists")
    regs = mmcv_shim.install()
    _purge("projects")
    plug = os.path.join(root, "projects", "mmdet3d_plugin")
    _bare_package("projects", os.path.join(root, "projects"))
    _bare_package("projects.mmdet3d_plugin", plug)            # the heavy __init__ (datasets, ...) never runs
    _bare_package("projects.mmdet3d_plugin.datasets", os.path.join(plug, "datasets"))
    _bare_package("projects.mmdet3d_plugin.datasets.pipelines", os.path.join(plug, "datasets", "pipelines"))
    _stub("projects.mmdet3d_plugin.datasets.evaluation", PlanningMetric=type("PlanningMetric", (), {
        "__init__": lambda self, *a, **k: None}))
    _stub("projects.mmdet3d_plugin.datasets.pipelines.vectorize_numpy", VectorizeMapNumpy=type(
        "VectorizeMapNumpy", (), {"__init__": lambda self, *a, **k: None}))

    def _no_corners(*a, **k):
        raise NotImplementedError("box3d_to_corners is only used by the collision metrics, not by the decoder forward")
    _stub("projects.mmdet3d_plugin.datasets.utils", box3d_to_corners=_no_corners, box3d_to_corners_gpu=_no_corners)

    want_module = False
    if isinstance(ops, str):
        kind = ops
        if kind == "reference":
            ops_mod = reference_ops_module(root)
        elif kind in ("ours", "ours_module", "ours_module_graph"):
            import hipad_b200
            ops_mod = hipad_b200.ops
            want_module = kind != "ours"
        else:
            raise ValueError(kind)
    else:
        ops_mod = ops
    counted = CountingOps(ops_mod)
    sys.modules["projects.mmdet3d_plugin.ops"] = counted
    models = importlib.import_module("projects.mmdet3d_plugin.models")
    if want_module:
        import hipad_b200
        # only the aggregation module is swapped: the instance banks keep using the reference's key-point generators
        # as anchor handlers (anchor_projection etc. is bank logic, outside the path); our module builds its own
        # generators (hipad_b200.blocks), whose state-dict keys are the reference's
        regs["ATTENTION"].register_module("DeformableFeatureAggregation", module=hipad_b200.DeformableFeatureAggregation)
    return models, regs, counted


def load_config(root, hw=(352, 640)):
    """exec the vendored stage-2 config with the authors' absolute ``project_dir`` pointed at the vendored root and
    ``input_shape`` set to hw (feature_map_scale of the ego/plan banks follows it)."""
    path = os.path.join(root, CONFIG)
    src = open(path).read()
    src = src.replace('project_dir = "/opt/data/private/project/HiP-AD"', 'project_dir = %r' % root)
    src = src.replace("input_shape = (640, 352)", "input_shape = (%d, %d)" % (hw[1], hw[0]))
    ns = {}
    exec(compile(src, path, "exec"), ns)
    return ns


def build_decoder(ops="ours", hw=(352, 640), seed=0, device="cpu", weights_fc_std=0.02):
    """SparseOneDecoder(**model.head.onedecoder_head), init_weights(), weights_fc re-initialised N(0, std)
    (the reference zero-init would make all aggregation weights uniform), eval mode."""
    models, regs, counted = load_models(ops)
    cfg = load_config(vendor.DST, hw)
    head = dict(cfg["model"]["head"]["onedecoder_head"])
    head.pop("type")
    torch.manual_seed(seed)
    dec = models.SparseOneDecoder(**head)
    dec.init_weights()
    g = torch.Generator().manual_seed(seed + 1)
    for name, m in dec.named_modules():
        if name.endswith("weights_fc") and isinstance(m, torch.nn.Linear):
            with torch.no_grad():
                m.weight.copy_(torch.randn(m.weight.shape, generator=g) * weights_fc_std)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * weights_fc_std)
    dec = dec.to(device).eval()
    if ops == "ours_module_graph":      # our module replays its inference forward as one CUDA graph per call
        import hipad_b200
        for m in dec.modules():
            if isinstance(m, hipad_b200.DeformableFeatureAggregation):
                m.graph_inference = True
    dec._hipad_ops = counted
    dec._hipad_models = models
    return dec


def copy_weights(dst, src):
    """Same parameters in two decoders built from different op packages."""
    missing = dst.load_state_dict(src.state_dict(), strict=True)
    return missing


def make_frames(n_frames, bs=1, hw=(352, 640), seed=0, device="cpu", dtype=torch.float32):
    """Synthetic inputs of consecutive frames: feature maps N(0,1) (4 levels, 6 cams, 256 ch), camera matrices of a
    Bench2Drive-like rig, identity ego motion, timestamps +0.5 s per frame."""
    rng = np.random.default_rng(seed)
    proj, wh = camera_matrices(hw)
    frames = []
    for f in range(n_frames):
        levels = [torch.from_numpy(rng.standard_normal((bs, 6, 256, hw[0] // s, hw[1] // s), dtype=np.float32))
                  .to(device=device, dtype=dtype) for s in (4, 8, 16, 32)]
        metas = dict(
            projection_mat=torch.from_numpy(np.tile(proj[None], (bs, 1, 1, 1))).to(device),
            image_wh=torch.from_numpy(np.tile(wh[None], (bs, 1, 1))).to(device),
            timestamp=torch.full((bs,), 0.5 * f, dtype=torch.float64, device=device),
            img_metas=[dict(T_global=np.eye(4), T_global_inv=np.eye(4), timestamp=0.5 * f) for _ in range(bs)],
            target_point=torch.from_numpy(np.tile(np.array([[0.0, 30.0]], dtype=np.float32), (bs, 1))).to(device),
            gt_ego_fut_cmd=torch.from_numpy(np.tile(np.eye(6, dtype=np.float32)[3:4], (bs, 1))).to(device),
        )
        frames.append((levels, metas))
    return frames


def reset(dec):
    for name in ("det_instance_bank", "map_instance_bank", "ego_instance_bank", "plan_instance_bank"):
        bank = getattr(dec, name, None)
        if bank is not None and hasattr(bank, "reset"):
            bank.reset()
    if getattr(dec, "is_init_bank_list", False):
        for lst in ("det_instance_bank_list", "map_instance_bank_list", "ego_instance_bank_list",
                    "plan_instance_bank_list"):
            for bank in getattr(dec, lst, []):
                if hasattr(bank, "reset"):
                    bank.reset()


def run_frame(dec, levels, metas):
    """feature_maps_format (of the plugged-in ops package) + decoder forward; returns the decoder's output dict."""
    ops = dec._hipad_ops
    fm = ops.feature_maps_format(levels)
    img = levels[0].new_zeros((levels[0].shape[0], 6, 3, 8, 8))
    return dec(img, fm, metas)


def flatten_outputs(out, prefix=""):
    """{name: tensor} over every tensor in the (nested) decoder output."""
    flat = {}
    if torch.is_tensor(out):
        flat[prefix or "out"] = out
    elif isinstance(out, dict):
        for k, v in out.items():
            flat.update(flatten_outputs(v, "%s.%s" % (prefix, k) if prefix else str(k)))
    elif isinstance(out, (list, tuple)):
        for i, v in enumerate(out):
            flat.update(flatten_outputs(v, "%s[%d]" % (prefix, i)))
    return flat


def use_sdpa_attention(dec):
    """Replace the reference's flash-attn call (fp16, CUDA only, models/attention.py:53-100) by
    torch.nn.functional.scaled_dot_product_attention in the input dtype.  Used for CPU runs of the harness and for
    fp32 parity runs on the GPU (flash-attn's fp16 rounding would otherwise sit between the two op variants)."""
    import torch.nn.functional as F
    attention = dec._hipad_models.attention      # this decoder's own import of models/attention.py

    def forward(self, q, kv, causal=False, key_padding_mask=None):
        # q (B,T,H,D), kv (B,S,2,H,D), key_padding_mask (B,S) True = keep
        k, v = kv[:, :, 0], kv[:, :, 1]
        mask = None
        if key_padding_mask is not None:
            mask = key_padding_mask[:, None, None, :].to(torch.bool)
        out = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2),
                                             attn_mask=mask, dropout_p=0.0, is_causal=causal,
                                             scale=self.softmax_scale)
        return out.transpose(1, 2), None

    attention.FlashAttention.forward = forward
    return dec
