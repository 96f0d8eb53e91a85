"""Small forward passes for compute-sanitizer (run: compute-sanitizer --tool memcheck python tools/sanitize_small.py)."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from neusomatic_b200 import Engine, synth
from oracle import nsnet_oracle as oracle
for C in (26, 119):
    sd = synth.random_state_dict(C, seed=3)
    cand = synth.structured_pileups(130, seed=5)
    anns = np.random.default_rng(0).random((130, 93)).astype(np.float32) if C == 119 else None
    x = oracle.assemble_channels(cand.matrices, cand.center, cand.tumor_cov, cand.normal_cov, anns, 100)
    for prec in (3, 0):
        eng = Engine(sd, C, precision=prec)
        o = eng.forward_f32(torch.from_numpy(x).cuda())
        torch.cuda.synchronize()
        print("C=%d precision=%d ok: %s" % (C, prec, [float(t.float().abs().max()) for t in o[:3]]))
        eng.close()
