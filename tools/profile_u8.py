"""Per-kernel times of the uint8 interface (forward_u8, device-resident compact inputs): python tools/profile_u8.py [--ensemble]"""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden
from neusomatic_b200 import Engine, synth
from neusomatic_b200.call import load_checkpoint
from neusomatic_b200.engine import anns_to_255
ens = "--ensemble" in sys.argv
g, ck = load_golden("ens_v010" if ens else "std_v010")
sd, meta = load_checkpoint(ck)
C = 119 if ens else 26
B = 65536
eng = Engine(sd, C)
rng = np.random.default_rng(0)
mat = torch.from_numpy(rng.integers(0, 256, (B, 5, 32, 23), dtype=np.uint8)).cuda()
m3 = torch.from_numpy(np.stack([rng.integers(2, 30, B), rng.integers(5, 400, B), rng.integers(5, 400, B)], 1).astype(np.int32)).cuda()
ann = torch.from_numpy(anns_to_255(np.round(rng.random((B, 93)), 4).astype(np.float32))).cuda() if ens else None
for _ in range(2):
    eng.forward_u8(mat, m3, ann, float(meta["coverage_thr"]))
torch.cuda.synchronize()
eng.set_profile(True)
K = 4
for _ in range(K):
    eng.forward_u8(mat, m3, ann, float(meta["coverage_thr"]))
torch.cuda.synchronize()
p = eng.get_profile()
tot = sum(v[0] for v in p.values())
print("C=%d B=%d: %.3f ms per pass -> %.2f M cand/s" % (C, B, tot / K, B * K / tot / 1e3))
print({k: round(v[0] / K, 3) for k, v in p.items() if v[1]})
