import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    from classpp_public_b200.modules import Inputs
    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = Inputs.load(os.path.join(GOLDEN, name + ".npz"))
        return cache[name]
    return load


@pytest.fixture(scope="session")
def reference():
    """Live oracle: the unmodified reference built into oracle/_ref (None if not built)."""
    from oracle import refprobe
    if not refprobe.available():
        return None
    from classpp_public_b200.configs import CONFIGS
    cache = {}

    def get(name, level="lensing"):
        key = (name, level)
        if key not in cache:
            cache[key] = refprobe.RefCosmology(CONFIGS[name], threads=os.cpu_count()).compute(level)
        return cache[key]
    return get


def cl_table(cl, ct_size):
    return np.asarray(cl).reshape(-1, ct_size)
