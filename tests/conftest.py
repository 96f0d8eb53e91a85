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


def vcf_identity(got, ref, score_tol=2.5e-4):
    """Line-by-line comparison of two VCF texts written by write_vcf.  Returns (identical lines, lines); asserts that the
    records are the same (chrom, pos, ref, alt) in the same order and that a line that is not byte-identical differs only
    in its 4-decimal SCORE / QUAL by rounding noise: |dSCORE| <= score_tol, and FILTER only if SCORE sits on a threshold."""
    a, b = got.splitlines(), ref.splitlines()
    assert len(a) == len(b), "record count %d vs %d" % (len(a), len(b))
    same = 0
    for x, y in zip(a, b):
        if x == y:
            same += 1
            continue
        fx, fy = x.split("\t"), y.split("\t")
        assert fx[:5] == fy[:5] and fx[8:] == fy[8:], "different record: %r vs %r" % (x, y)
        sx, sy = float(fx[7].split("=")[1]), float(fy[7].split("=")[1])
        assert abs(sx - sy) <= score_tol, "SCORE %r vs %r" % (x, y)
        if fx[6] != fy[6]:
            assert min(abs(sy - 0.7), abs(sy - 0.4)) <= score_tol, "FILTER %r vs %r" % (x, y)
    return same, len(b)
