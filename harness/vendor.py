"""Vendor the reference files the decoder harness needs into baseline/_ref/hipad/ (git-ignored, shipped by gpurun).

Runs only where /root/reference exists (the build container): the GPU box has no reference checkout, and nothing under
baseline/_ref is ever committed.  What is copied, verbatim and read-only:
  projects/mmdet3d_plugin/{models,core}   the decoder and everything it imports
  projects/mmdet3d_plugin/ops/*.py        the reference's own op package (Python side; its CUDA extension is
                                          oracle/_ref/deformable_aggregation_ext.so, built by oracle/build_ref.py)
  projects/configs/hipad_b2d_stage{1,2}.py
  data/kmeans/*.npy                       shipped anchors
The harness (harness/decoder.py) imports these through a mmcv/mmdet stand-in; no line of them is edited.
"""
import os
import shutil
import stat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("HIPAD_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref", "hipad")

_TREES = [
    ("projects/mmdet3d_plugin/models", "projects/mmdet3d_plugin/models"),
    ("projects/mmdet3d_plugin/core", "projects/mmdet3d_plugin/core"),
    ("projects/configs", "projects/configs"),
    ("data/kmeans", "data/kmeans"),
]
_FILES = [
    ("projects/mmdet3d_plugin/ops/__init__.py", "projects/mmdet3d_plugin/ops/__init__.py"),
    ("projects/mmdet3d_plugin/ops/deformable_aggregation.py", "projects/mmdet3d_plugin/ops/deformable_aggregation.py"),
    ("projects/mmdet3d_plugin/ops/deformable_aggregation_a800.py", "projects/mmdet3d_plugin/ops/deformable_aggregation_a800.py"),
]


def available():
    return os.path.isfile(os.path.join(DST, "projects", "mmdet3d_plugin", "models", "sparse_onedecoder.py"))


def vendor(force=False):
    """Copy (once) and return the vendored root, or None when neither the reference nor a previous copy exists."""
    if available() and not force:
        return DST
    if not os.path.isdir(REF):
        return DST if available() else None
    if os.path.isdir(DST):
        for dirpath, _, files in os.walk(DST):
            os.chmod(dirpath, 0o755)
            for f in files:
                os.chmod(os.path.join(dirpath, f), 0o644)
        shutil.rmtree(DST)
    for src, dst in _TREES:
        shutil.copytree(os.path.join(REF, src), os.path.join(DST, dst),
                        ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    for src, dst in _FILES:
        os.makedirs(os.path.dirname(os.path.join(DST, dst)), exist_ok=True)
        shutil.copy2(os.path.join(REF, src), os.path.join(DST, dst))
    for dirpath, _, files in os.walk(DST):          # read-only: the harness never edits a reference file
        for f in files:
            os.chmod(os.path.join(dirpath, f), stat.S_IRUSR | stat.S_IRGRP | stat.S_IROTH)
    return DST


if __name__ == "__main__":
    print(vendor(force=True))
