import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_sessionstart(session):
    """The shared library is a build artefact (git-ignored): build it in-tree if a fresh checkout has none,
    and the reference's tracking.c checker when the reference is mounted."""
    from sydr_b200 import build as B
    if not os.path.exists(B.LIB):
        B.build()
    ref_so = os.path.join(ROOT, "oracle", "_ref", "tracking.so")
    if not os.path.exists(ref_so) and os.path.isdir("/root/reference/sydr/c_functions"):
        import subprocess
        subprocess.call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    return load
