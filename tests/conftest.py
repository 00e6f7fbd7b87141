import os
import sys
from os.path import dirname, join, realpath

import numpy as np
import pytest

ROOT = dirname(dirname(realpath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden(name):
    return np.load(join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def structures():
    return golden("structures.npz")


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
