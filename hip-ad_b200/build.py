"""Build the C-ABI CUDA library in-tree: hip-ad_b200/lib/libhipad_dfa.so (sm_100a only).

nvcc cross-compiles without a GPU; the .so is git-ignored but travels with gpurun snapshots.
"""
import concurrent.futures
import hashlib
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_DIR = os.path.join(_HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libhipad_dfa.so")
OBJ_DIR = os.path.join(LIB_DIR, "obj")
SOURCES = ["dfa_forward.cu", "dfa_backward.cu", "dfa_group.cu", "dfa_api.cu", "dfa_format.cu", "dfa_weights.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"] + os.environ.get("HIPAD_DFA_NVCC_EXTRA", "").split()


def _fingerprint():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(_HERE, "..", "include")):
        for name in sorted(os.listdir(root)):
            if name.endswith((".cu", ".cuh", ".h")):
                h.update(name.encode())
                h.update(open(os.path.join(root, name), "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile (if sources changed) and return the library path."""
    os.makedirs(OBJ_DIR, exist_ok=True)
    stamp = os.path.join(LIB_DIR, "build.stamp")
    fp = _fingerprint()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp) and open(stamp).read() == fp:
        return LIB_PATH

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        cmd = ["nvcc"] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s" % (" ".join(cmd), r.stdout + r.stderr))
        if verbose:
            print(" ".join(cmd))
        return obj

    with concurrent.futures.ThreadPoolExecutor(len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = ["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs + ["-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed: %s\n%s" % (" ".join(cmd), r.stdout + r.stderr))
    open(stamp, "w").write(fp)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
