"""CPU oracle for the deformable-aggregation hot path.  TEST INFRASTRUCTURE ONLY.

Nothing in the product package (``hip-ad_b200/``) imports this module.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may use it, and only as the checker / CPU baseline.

Two restatements live here:

* ``dfa_oracle.c`` (wrapped below): the reference CUDA op's semantics
  (``projects/mmdet3d_plugin/ops/src/deformable_aggregation_cuda.cu:13-262``),
  plain C, fp32 per-sample arithmetic, fp64 cross-sample accumulation.
* ``torch_path.py``: the reference's pure-torch ``grid_sample`` path
  (``projects/mmdet3d_plugin/models/blocks.py:216-264``), which is also the
  reference's own CPU implementation timed by ``bench.py --impl reference``.

Parity pin: both are checked against ``tests/golden/*.npz`` — outputs of the
reference's unmodified ``blocks.py`` imported in the build container by
``tests/golden/make_golden.py`` — and, on the GPU box, against the reference
CUDA op built from its own sources into ``oracle/_ref/`` (``build_ref.py``).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
_LIB_PATH = os.path.join(_BUILD, "libdfa_oracle.so")
_SRC = os.path.join(_HERE, "dfa_oracle.c")
_lib = None


def build_oracle(force=False):
    """gcc the C restatement into oracle/_build/libdfa_oracle.so."""
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(_SRC)):
        return _LIB_PATH
    os.makedirs(_BUILD, exist_ok=True)
    cmd = ["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-shared", "-fPIC",
           "-Wall", "-o", _LIB_PATH, _SRC, "-lm"]
    subprocess.run(cmd, check=True)
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build_oracle()
        lib = ctypes.CDLL(_LIB_PATH)
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int32)
        i = ctypes.c_int
        lib.dfa_oracle_forward.argtypes = [fp, fp, ip, ip, fp, fp] + [i] * 8
        lib.dfa_oracle_forward.restype = None
        lib.dfa_oracle_backward.argtypes = [fp, ip, ip, fp, fp, fp, fp, fp, fp] + [i] * 8
        lib.dfa_oracle_backward.restype = None
        lib.dfa_oracle_indices.argtypes = [ip, ip, ip, fp] + [i] * 5
        lib.dfa_oracle_indices.restype = None
        lib.dfa_oracle_unique_rows.argtypes = [ip, ip, fp] + [i] * 6
        lib.dfa_oracle_unique_rows.restype = ctypes.c_int64
        _lib = lib
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def _dims(feat, shapes, loc, weights):
    bs, num_feat, C = feat.shape
    cams, L = shapes.shape[:2]
    A, P = loc.shape[1:3]
    G = weights.shape[5]
    assert loc.shape == (bs, A, P, cams, 2), loc.shape
    assert weights.shape == (bs, A, P, cams, L, G), weights.shape
    assert C % G == 0
    return bs, cams, num_feat, C, L, A, P, G


def forward(feat, shapes, starts, loc, weights):
    """numpy in / numpy out: [bs,A,C] float32."""
    feat, pf = _f(feat); loc, pl = _f(loc); weights, pw = _f(weights)
    shapes, ps = _i(shapes); starts, pt = _i(starts)
    d = _dims(feat, shapes, loc, weights)
    out = np.empty((d[0], d[5], d[3]), np.float32)
    _load().dfa_oracle_forward(out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                               pf, ps, pt, pl, pw, *d)
    return out


def backward(feat, shapes, starts, loc, weights, grad_out):
    """Returns (g_feat, g_loc, g_w), all dense float32."""
    feat, pf = _f(feat); loc, pl = _f(loc); weights, pw = _f(weights)
    grad_out, pg = _f(grad_out)
    shapes, ps = _i(shapes); starts, pt = _i(starts)
    d = _dims(feat, shapes, loc, weights)
    assert grad_out.shape == (d[0], d[5], d[3])
    g_feat = np.empty_like(feat); g_loc = np.empty_like(loc); g_w = np.empty_like(weights)
    c = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
    _load().dfa_oracle_backward(pf, ps, pt, pl, pw, pg, c(g_feat), c(g_loc), c(g_w), *d)
    return g_feat, g_loc, g_w


def indices(shapes, starts, loc):
    """int32 [bs,A,P,cams,L,6]: valid, h_low, w_low, level_offset, corner_mask, row0."""
    loc, pl = _f(loc)
    shapes, ps = _i(shapes); starts, pt = _i(starts)
    bs, A, P, cams, _ = loc.shape
    L = shapes.shape[1]
    idx = np.empty((bs, A, P, cams, L, 6), np.int32)
    _load().dfa_oracle_indices(idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                               ps, pt, pl, bs, cams, L, A, P)
    return idx


def unique_rows(shapes, starts, loc, num_feat):
    """U of SURVEY.md §8(d): distinct feature rows touched by the call."""
    loc, pl = _f(loc)
    shapes, ps = _i(shapes); starts, pt = _i(starts)
    bs, A, P, cams, _ = loc.shape
    L = shapes.shape[1]
    return int(_load().dfa_oracle_unique_rows(ps, pt, pl, bs, cams, num_feat, L, A, P))
