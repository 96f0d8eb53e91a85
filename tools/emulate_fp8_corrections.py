"""CPU emulation of the tensor-path operand arithmetic (oracle/split8_emulation.py) against the golden reference outputs:
how far each mode is from the reference with EXACT accumulation, i.e. the error budget of the operand formats alone."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden
from oracle import nsnet_oracle as oracle
from oracle import split8_emulation as emu

torch.set_num_threads(8)
for name in ("std_v010", "ens_v010", "std_v014", "ens_v014", "std_wex"):
    g, ck = load_golden(name)
    p = {k.replace("module.", ""): v for k, v in ck["state_dict"].items()}
    x = torch.from_numpy(oracle.assemble_channels(g["matrices"], g["center"], g["tumor_cov"], g["normal_cov"], g.get("anns"), int(ck["coverage_thr"])))
    for mode in ("exact", "split16", "split8", "single16"):
        o1, o2, o3 = emu.forward(p, x, mode)
        r1, r3 = oracle.softmax(g["o1"]), oracle.softmax(g["o3"])
        print("%s %-9s max|dlogit|=%.2e max|dprob|=%.2e max|dpos|=%.2e flips=%d/%d" % (
            name, mode, max(np.abs(o1 - g["o1"]).max(), np.abs(o3 - g["o3"]).max()),
            max(np.abs(oracle.softmax(o1) - r1).max(), np.abs(oracle.softmax(o3) - r3).max()), np.abs(o2 - g["o2"]).max(),
            (o1.argmax(1) != g["o1"].argmax(1)).sum(), (o3.argmax(1) != g["o3"].argmax(1)).sum()), flush=True)
