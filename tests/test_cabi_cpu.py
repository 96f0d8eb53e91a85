"""CPU: the C-ABI library builds, loads and exports every symbol include/*.h declares;
host-side logic of the call.py mirror.  No compute calls (there is no GPU here)."""
import ctypes as C
import glob
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from oracle import nsnet_oracle as oracle


@pytest.fixture(scope="module")
def lib():
    from neusomatic_b200 import build, _lib
    build.build_lib()
    return _lib.load()


def declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names |= set(re.findall(r"\b(nsc_[a-z0-9_]+)\s*\(", src))
    return names


def test_every_declared_symbol_is_exported_and_bound(lib):
    from neusomatic_b200 import _lib
    decl = declared_symbols()
    assert len(decl) >= 15
    for name in decl:
        assert hasattr(lib, name), "library does not export " + name
    assert decl == set(_lib.SIGNATURES), "ctypes table and header disagree"


def test_abi_version_and_loud_failure_without_gpu(lib):
    import torch
    assert lib.nsc_abi_version() == 1
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = lib.nsc_model_create(C.byref(h), 0, 26)
    assert rc != 0 and not h.value
    assert b"no CPU fallback" in lib.nsc_last_error()


def test_module_refuses_cpu_inference():
    import torch
    from neusomatic_b200.network import NeuSomaticNet
    net = NeuSomaticNet(26).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.zeros(1, 26, 5, 32))


def test_module_state_dict_layout_matches_reference_checkpoints(golden):
    """call.py:441-460: the bundled .pth must load with strict key matching."""
    from neusomatic_b200.network import NeuSomaticNet
    from neusomatic_b200.call import load_checkpoint
    name, g, ck = golden
    sd, meta = load_checkpoint(ck)
    net = NeuSomaticNet(119 if "anns" in g else 26)
    missing, unexpected = net.load_state_dict(sd, strict=False)
    assert [k for k in missing if not k.endswith("num_batches_tracked")] == [] and unexpected == []
    assert meta["coverage_thr"] == int(g["coverage_thr"])


def test_build_pred_dicts_equals_reference_loop(golden):
    """Vectorised post-loop == restated call.py:86-114 loop, value for value."""
    from neusomatic_b200.call import build_pred_dicts
    name, g, ck = golden
    tags = [str(t) for t in g["tags"]]
    o1, o2, o3 = g["o1"], g["o2"], g["o3"]
    f_ref, n_ref = oracle.call_variants(o1, o2, o3, tags)
    f, n = build_pred_dicts(o1, o2, o3, oracle.softmax(o1), oracle.softmax(o3), o1.argmax(1), o3.argmax(1), tags)
    assert set(f) == set(f_ref) and set(n) == set(n_ref)
    for d, d_ref in ((f, f_ref), (n, n_ref)):
        for k in d_ref:
            a, b = d[k], d_ref[k]
            assert a[0] == b[0] and a[2] == b[2] and np.array_equal(a[1], b[1])
            for i in (3, 4, 5, 6):
                assert [float(v) for v in a[i]] == [float(v) for v in b[i]]
                assert all(isinstance(v, np.float32) for v in a[i])


def test_shard_ranges_cover_everything():
    from neusomatic_b200.call import shard_ranges
    for n in (0, 1, 7, 8, 1000003):
        for ws in (1, 2, 3, 8):
            r = shard_ranges(n, ws)
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_anns_to_255_matches_reference_float64_path():
    from neusomatic_b200 import anns_to_255
    a = np.float32([[0.1234, 0.0, 1.0, 0.3333]])
    ref = (a.astype(np.float64) * 255.0).astype(np.float32)
    assert np.array_equal(anns_to_255(a), ref)
