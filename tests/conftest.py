import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build_oracle()
    return oracle


@pytest.fixture(scope="session")
def cuda_lib():
    """Build (if stale) and load the C-ABI library; GPU tests fail loudly if it is missing."""
    sys.path.insert(0, os.path.join(ROOT, "hip-ad_b200"))
    import importlib.util
    spec = importlib.util.spec_from_file_location("hipad_build", os.path.join(ROOT, "hip-ad_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()
    from hipad_b200 import _lib
    return _lib.get()
