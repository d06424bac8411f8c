import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(ROOT, "drone-sim-python_b200")
for p in (ROOT, PKG_DIR):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # the native pieces are build artefacts (git-ignored): compile them when a fresh checkout is tested before `__graft_entry__.build()`
    # ran (nvcc cross-compiles without a GPU; make is a no-op when everything is up to date)
    import shutil
    import subprocess
    need = [os.path.join(PKG_DIR, "d2d_b200", "libd2dx.so"), os.path.join(ROOT, "tests", "native", "libd2dx_hostcheck.so"),
            os.path.join(ROOT, "oracle", "_build", "liboracle.so")]
    if not all(os.path.exists(f) for f in need) and shutil.which("nvcc") and shutil.which("make"):
        subprocess.run(["make", "-C", os.path.join(PKG_DIR, "csrc"), "-j4"], check=False, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        if os.path.exists(os.path.join(ROOT, "oracle", "Makefile")):
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=False, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    class _G:
        def __getitem__(self, name):
            return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return _G()
