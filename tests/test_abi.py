"""CPU tests of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and exports every
symbol include/hipad_dfa.h declares (no compute calls: there is no GPU here)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "hipad_dfa.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hipad_dfa_\w+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    names = declared_functions()
    for must in ("hipad_dfa_forward_f32", "hipad_dfa_forward_bf16", "hipad_dfa_backward_f32",
                 "hipad_dfa_backward_bf16", "hipad_dfa_backward_workspace_bytes", "hipad_dfa_sample_indices",
                 "hipad_dfa_fused_forward_f32", "hipad_dfa_fused_forward_bf16", "hipad_dfa_error_string",
                 "hipad_dfa_version", "hipad_dfa_backward_stages"):
        assert must in names


def test_library_exports_every_declared_symbol(cuda_lib):
    from hipad_b200 import _lib
    for name in declared_functions():
        assert hasattr(cuda_lib, name), name
    assert set(_lib.EXPORTED_SYMBOLS) == set(declared_functions())
    assert cuda_lib.hipad_dfa_version() == 1


def test_header_is_plain_c(tmp_path):
    c = tmp_path / "t.c"
    c.write_text('#include "hipad_dfa.h"\nint main(void){return HIPAD_DFA_VERSION == 1 ? 0 : 1;}\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    "-c", str(c), "-o", str(tmp_path / "t.o")], check=True)


def test_library_is_sm100a_only_and_torch_free(cuda_lib):
    from hipad_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs
    ldd = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libtorch" not in ldd and "libc10" not in ldd


def test_host_side_status_codes(cuda_lib):
    # argument validation happens before any CUDA call, so it is testable without a GPU
    assert cuda_lib.hipad_dfa_forward_f32(None, None, None, None, None, None, 1, 1, 1, 1, 1, 1, 1, 1, None) == -1
    p = ctypes.c_void_p(16)
    assert cuda_lib.hipad_dfa_forward_f32(p, p, p, p, p, p, 1, 6, 100, 255, 4, 10, 13, 8, None) == -1   # C % G != 0
    assert cuda_lib.hipad_dfa_forward_f32(p, p, p, p, p, p, 0, 6, 100, 256, 4, 10, 13, 8, None) == -1   # bs <= 0
    assert cuda_lib.hipad_dfa_forward_f32(p, p, p, p, p, p, 1, 6, 100, 4096, 4, 10, 13, 8, None) == -2  # C too wide
    assert cuda_lib.hipad_dfa_backward_f32(p, p, p, p, p, p, p, p, p, 1, 6, 100, 256, 4, 10, 13, 8,
                                           None, 0, None) == -3                                        # no workspace
    assert b"workspace" in cuda_lib.hipad_dfa_error_string(-3)
    n = cuda_lib.hipad_dfa_backward_workspace_bytes(1, 6, 112200, 256, 4, 900, 13, 8)
    assert n > 0 and n % 256 == 0
    assert cuda_lib.hipad_dfa_backward_workspace_bytes(0, 6, 112200, 256, 4, 900, 13, 8) == 0


def test_missing_library_fails_loudly(monkeypatch):
    from hipad_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libhipad_dfa.so")
    with pytest.raises(_lib.HipadDfaError, match="no CPU or PyTorch fallback"):
        _lib.get()


def test_cpu_tensors_are_rejected():
    import torch
    import hipad_b200
    with pytest.raises(RuntimeError, match="no CPU path"):
        hipad_b200.deformable_aggregation_function(
            torch.zeros(1, 4, 8), torch.tensor([[[2, 2]]]), torch.tensor([[0]]),
            torch.zeros(1, 1, 1, 1, 2), torch.zeros(1, 1, 1, 1, 1, 2))


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "hip-ad_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, os.path.join(dirpath, f)
