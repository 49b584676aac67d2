"""ctypes loader for lib/libhipad_dfa.so (the C ABI declared in include/hipad_dfa.h).

There is no fallback of any kind: if the library is missing or a call fails, we raise.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libhipad_dfa.so")
_lib = None

_p = ctypes.c_void_p
_i = ctypes.c_int
_DIMS8 = [_i] * 8

_SIGNATURES = {
    "hipad_dfa_version": ([], _i),
    "hipad_dfa_error_string": ([_i], ctypes.c_char_p),
    "hipad_dfa_forward_f32": ([_p] * 6 + _DIMS8 + [_p], _i),
    "hipad_dfa_forward_bf16": ([_p] * 6 + _DIMS8 + [_p], _i),
    "hipad_dfa_backward_workspace_bytes": (_DIMS8, ctypes.c_size_t),
    "hipad_dfa_backward_f32": ([_p] * 9 + _DIMS8 + [_p, ctypes.c_size_t, _p], _i),
    "hipad_dfa_backward_bf16": ([_p] * 9 + _DIMS8 + [_p, ctypes.c_size_t, _p], _i),
    "hipad_dfa_backward_accumulate_f32": ([_p] * 9 + _DIMS8 + [_p, ctypes.c_size_t, _p], _i),
    "hipad_dfa_backward_accumulate_bf16": ([_p] * 9 + _DIMS8 + [_p, ctypes.c_size_t, _p], _i),
    "hipad_dfa_backward_stages": ([_i, _i] + [_p] * 9 + _DIMS8 + [_p, ctypes.c_size_t, _p], _i),
    "hipad_dfa_sample_indices": ([_p] * 4 + [_i] * 5 + [_p], _i),
    "hipad_dfa_fused_forward_f32": ([_p] * 9 + _DIMS8 + [_p], _i),
    "hipad_dfa_fused_forward_bf16": ([_p] * 9 + _DIMS8 + [_p], _i),
    "hipad_dfa_format_features": ([_i, _i, _i, _p, _p, _p, _i, _i, _i, _i, _p], _i),
    "hipad_dfa_group_forward_workspace_bytes": ([_p, _i, _i, _i, _i], ctypes.c_size_t),
    "hipad_dfa_group_forward": ([_i, _p, _p, _p, _p, _p, _i] + [_i] * 6 + [_p, ctypes.c_size_t, _p], _i),
    "hipad_dfa_group_backward_workspace_bytes": ([_p, _i] + [_i] * 6, ctypes.c_size_t),
    "hipad_dfa_group_backward": ([_i, _i, _p, _p, _p, _p, _i, _p, _p] + [_i] * 6 + [_p, ctypes.c_size_t, _p], _i),
    "hipad_dfa_group_backward_stages": ([_i, _i, _i, _p, _p, _p, _p, _i, _p, _p] + [_i] * 6 + [_p, ctypes.c_size_t, _p], _i),
    "hipad_dfa_debug_counters_offset": (_DIMS8, ctypes.c_size_t),
    "hipad_dfa_weights_forward": ([_p, _p, _p, _p, ctypes.c_ulonglong, ctypes.c_float, ctypes.c_longlong, _i, _i, _i, _i, _p], _i),
    "hipad_dfa_weights_backward": ([_p, _p, _p, _p, _p, ctypes.c_ulonglong, ctypes.c_float, ctypes.c_longlong, _i, _i, _i, _i, _p], _i),
}


class CallT(ctypes.Structure):
    """hipad_dfa_call_t of include/hipad_dfa.h."""
    _fields_ = [("sample_location", _p), ("weights", _p), ("grad_sampling_location", _p), ("grad_weights", _p),
                ("num_anchors", ctypes.c_int32), ("num_pts", ctypes.c_int32)]


MAX_GROUP_CALLS = 8
ERR_UNSUPPORTED = -2


def call_table(entries):
    """entries: list of (loc_ptr, w_ptr, g_loc_ptr or None, g_w_ptr or None, A, P) -> ctypes array of hipad_dfa_call_t"""
    arr = (CallT * len(entries))()
    for i, (loc, w, gl, gw, A, P) in enumerate(entries):
        arr[i].sample_location = loc
        arr[i].weights = w
        arr[i].grad_sampling_location = gl
        arr[i].grad_weights = gw
        arr[i].num_anchors = A
        arr[i].num_pts = P
    return arr
EXPORTED_SYMBOLS = tuple(_SIGNATURES)


class HipadDfaError(RuntimeError):
    pass


def get():
    """Load (once) and return the ctypes handle.  Raises if the CUDA library is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HipadDfaError(
                "hipad_dfa: %s not found. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `python hip-ad_b200/build.py`). There is no CPU or PyTorch fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (argtypes, restype) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = restype
        _lib = lib
    return _lib


def check(status, what):
    if status != 0:
        msg = get().hipad_dfa_error_string(status)
        raise HipadDfaError("%s failed with status %d: %s" % (what, status, msg.decode() if msg else "?"))
