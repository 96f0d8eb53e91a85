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


def load_golden(name):
    """Golden fixture written by tests/golden/make_golden.py (reference outputs)."""
    import torch
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))
    ck = torch.load(os.path.join(GOLDEN, name + ".pth"), map_location="cpu", weights_only=True)
    return g, ck


@pytest.fixture(scope="session", params=["std_v010", "ens_v010", "std_v014", "ens_v014", "std_wex"])
def golden(request):
    g, ck = load_golden(request.param)
    return request.param, g, ck
