import math
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle32():
    from oracle.oracle import Oracle, build
    build()
    return Oracle("f32")


@pytest.fixture(scope="session")
def oracle64():
    from oracle.oracle import Oracle, build
    build()
    return Oracle("f64")


GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_FILES = ["hand_placed_7.npz", "random_300_72x40.npz", "init_500_64x64.npz"]


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name))
    return {k: z[k] for k in z.files}


def split14(g):
    """[N,14] -> means, opacity, scales, rots, rgb  (/root/reference/core/gs.py:45-49)."""
    return g[:, 0:3].copy(), g[:, 3].copy(), g[:, 4:7].copy(), g[:, 7:11].copy(), g[:, 11:14].copy()


def tan_half(fovy):
    return math.tan(0.5 * math.radians(fovy))
