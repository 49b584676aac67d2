#!/usr/bin/env python
"""Compile the reference's own CUDA op, from its sources where they lie, into oracle/_ref/.

TEST INFRASTRUCTURE ONLY.  Sources are read from /root/reference and never copied;
only the built ``deformable_aggregation_ext*.so`` lands in ``oracle/_ref/`` (git-ignored,
shipped to the GPU box by gpurun).  It is (a) the on-GPU parity oracle and (b) the
"reference kernel recompiled for sm_100a on the same B200" that bench.py times beside ours.

Recipe = the reference's ops/setup.py (torch CUDAExtension, the three -D__CUDA_NO_HALF*
flags, setup.py:27-31) with an explicit -gencode for sm_100a, run directly with nvcc/g++
instead of through setuptools (no writes into the read-only source tree).
"""
import os
import subprocess
import sys
import sysconfig

REF_SRC = "/root/reference/projects/mmdet3d_plugin/ops/src"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_ref")
EXT_NAME = "deformable_aggregation_ext"


def ext_path():
    return os.path.join(OUT_DIR, EXT_NAME + ".so")


def available():
    return os.path.exists(ext_path())


def build(force=False, verbose=True):
    """Returns the .so path, or None when the reference tree is absent (GPU box)."""
    if not os.path.isdir(REF_SRC):
        return ext_path() if available() else None
    if available() and not force:
        return ext_path()
    import torch
    from torch.utils import cpp_extension as ce

    os.makedirs(OUT_DIR, exist_ok=True)
    tmp = os.path.join(OUT_DIR, "obj")
    os.makedirs(tmp, exist_ok=True)
    inc = [f"-I{p}" for p in ce.include_paths("cuda")] + [f"-I{sysconfig.get_paths()['include']}"]
    abi = f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"
    defs = [f"-DTORCH_EXTENSION_NAME={EXT_NAME}", "-DTORCH_API_INCLUDE_EXTENSION_H", abi]
    cu_o, cpp_o = os.path.join(tmp, "cuda.o"), os.path.join(tmp, "glue.o")
    nvcc = ["nvcc", "-c", os.path.join(REF_SRC, "deformable_aggregation_cuda.cu"), "-o", cu_o,
            "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3",
            "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC",
            "-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__",
            "-D__CUDA_NO_HALF2_OPERATORS__"] + defs + inc
    gxx = ["g++", "-c", os.path.join(REF_SRC, "deformable_aggregation.cpp"), "-o", cpp_o,
           "-std=c++17", "-O2", "-fPIC"] + defs + inc
    lib_dirs = ce.library_paths("cuda")
    link = ["g++", "-shared", cu_o, cpp_o, "-o", ext_path()] + [f"-L{d}" for d in lib_dirs] + \
           [f"-Wl,-rpath,{d}" for d in lib_dirs] + \
           ["-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart"]
    procs = [subprocess.Popen(c, stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for c in (nvcc, gxx)]
    for p, c in zip(procs, (nvcc, gxx)):
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("reference build failed: %s\n%s" % (" ".join(c), out.decode()))
    subprocess.run(link, check=True)
    if verbose:
        print("built", ext_path())
    return ext_path()


def load():
    """Import the built reference extension (needs torch; runs only where a GPU is)."""
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded before the extension)
    spec = importlib.util.spec_from_file_location(EXT_NAME, ext_path())
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    p = build(force="--force" in sys.argv)
    print(p or "reference sources absent and no prebuilt oracle/_ref")
