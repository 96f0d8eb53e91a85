"""Parity report of the CUDA path against the golden vectors (reference outputs) -- run on the GPU box."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden
from oracle import nsnet_oracle as oracle
from neusomatic_b200 import Engine
from neusomatic_b200.call import load_checkpoint

for name in ("std_v010", "ens_v010", "std_v014", "ens_v014", "std_wex"):
    g, ck = load_golden(name)
    x = oracle.assemble_channels(g["matrices"], g["center"], g["tumor_cov"], g["normal_cov"], g.get("anns"),
                                 int(ck["coverage_thr"]))
    sd, meta = load_checkpoint(ck)
    for prec in (0, 3, 1, 2):
        eng = Engine(sd, x.shape[1], precision=prec)
        o1, o2, o3 = [t.cpu().numpy() for t in eng.forward_f32(torch.from_numpy(x).cuda())[:3]]
        p1, p3 = oracle.softmax(o1), oracle.softmax(o3)
        r1, r3 = oracle.softmax(g["o1"]), oracle.softmax(g["o3"])
        rnd = lambda a: np.round(a, 4)
        mm = np.mean(rnd(p1) != rnd(r1)) * 0.5 + np.mean(rnd(p3) != rnd(r3)) * 0.5
        print("%s prec=%d n=%d: max|dlogit|=%.2e max|dprob|=%.2e max|dpos|=%.2e argmax flips type/len=%d/%d 4dp-mismatch=%.4f" % (
            name, prec, len(o1), max(np.abs(o1 - g["o1"]).max(), np.abs(o3 - g["o3"]).max()),
            max(np.abs(p1 - r1).max(), np.abs(p3 - r3).max()), np.abs(o2 - g["o2"]).max(),
            int((o1.argmax(1) != g["o1"].argmax(1)).sum()), int((o3.argmax(1) != g["o3"].argmax(1)).sum()), mm))
        eng.close()
