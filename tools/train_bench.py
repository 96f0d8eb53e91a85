"""Single-GPU train-step timing (CUDA events) -- tools/train_dp_check.py is the multi-GPU variant."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from neusomatic_b200 import Engine, synth
from neusomatic_b200.train import Trainer
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
K = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda", 0)
cand = synth.structured_pileups(B, seed=100)
_eng = Engine(synth.random_state_dict(26, seed=1), 26, device=0)      # channel assembly by the library, on the device
x = _eng.assemble(torch.from_numpy(cand.matrices).to(dev), torch.from_numpy(cand.meta).to(dev), None, 100.0).clone()
_eng.close()
vt = torch.tensor([synth.VARTYPE_CLASSES.index(t.split(".")[4]) for t in cand.tags], dtype=torch.int32, device=dev)
ln = torch.tensor([min(int(t.split(".")[6]), 3) for t in cand.tags], dtype=torch.int32, device=dev)
ce = torch.from_numpy(cand.center).float().to(dev)
w1 = torch.tensor([0.2, 0.21, 0.13, 0.16], device=dev); w3 = torch.tensor([0.12, 0.2, 0.23, 0.24], device=dev)
tr = Trainer(synth.random_state_dict(26, seed=11), 26, max_batch=B)
for _ in range(2):
    tr.step(x, vt, ce, ln, w1, w3)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(K):
    tr.step(x, vt, ce, ln, w1, w3)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print("train step B=%d: %.2f ms/step, %.0f samples/s, loss %.4f" % (B, ms, B / ms * 1000, float(tr.losses[0])))
