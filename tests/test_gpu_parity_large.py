"""GPU, large N: the north-star's argmax-identity bar (>= 99.99 % of candidates) can only be resolved on >= 1e4
candidates, and BASELINE configs[2] is a 1M-candidate call.  2 x 100 000 candidates (S2: bundled v0.1.4 weights; S3:
seeded random weights with near-tied logits) through the C ABI against oracle/torch_port.py on the host cores;
tools/parity_large.py is the full report (2e5 per family, C = 26 and 119, plus VCF-line identity) committed under
profiles/."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.gpu

N = 100000
# per mode: share of 4-decimal probability entries that may differ from the reference's (measured on 2e5 candidates,
# profiles/r2_parity_large.json, plus a margin): NSC_PREC_SPLIT8 (fast mode), NSC_PREC_SPLIT16 (default), NSC_PREC_FP32
MAX_MM4 = {3: 0.075, 0: 0.015, 2: 0.006}


@pytest.mark.parametrize("family", ["S2", "S3"])
def test_large_n_parity(family):
    import parity_large as pl
    from neusomatic_b200 import synth
    from neusomatic_b200.call import load_checkpoint
    torch.set_num_threads(os.cpu_count() or 1)
    if family == "S2":
        sd, meta = load_checkpoint(os.path.join(ROOT, "tests", "golden", "std_v014.pth"))
        thr = int(meta["coverage_thr"])
    else:
        sd, thr = synth.random_state_dict(26, seed=0), 100
    rep = pl.run_family(family, sd, 26, thr, N, 51000 if family == "S2" else 61000, [3, 0, 2], torch.device("cuda", 0))
    for mode, name in ((3, "SPLIT8"), (0, "SPLIT16"), (2, "FP32")):
        r = rep["modes"][name]
        assert r["n"] == N
        assert r["flips_type"] <= 1e-4 * N and r["flips_len"] <= 1e-4 * N, (name, r)          # >= 99.99 % identical
        assert r["max_dprob"] <= 2e-3 and r["max_dpos"] <= 1e-3, (name, r)
        assert r["mismatch_4dp_entries"] <= MAX_MM4[mode], (name, r)
