import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """`pytest tests/` on a box without CUDA: gpu-marked tests are skipped instead of failing in the driver."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device (gpu-marked test; run on the B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_meta():
    with open(os.path.join(GOLDEN_DIR, "golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_small():
    return np.load(os.path.join(GOLDEN_DIR, "golden_small.npz"))


@pytest.fixture(scope="session")
def golden_cubic():
    return np.load(os.path.join(GOLDEN_DIR, "golden_cubic.npz"))


@pytest.fixture(scope="session", autouse=True)
def built_library():
    """The C-ABI library must exist for every test; build it here if nvcc is available."""
    import cp360_b200
    import shutil
    if not os.path.exists(cp360_b200.LIB_PATH):
        if shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"):
            pytest.fail("libcp360.so missing and nvcc not available")
        cp360_b200.build_library()
    return cp360_b200.LIB_PATH
