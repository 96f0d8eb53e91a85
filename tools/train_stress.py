"""GPU: tensor-path training step against the fp32 CUDA-core step of the same library over many batch sizes (including
1, non-multiples of the 8-candidate groups and of the 16-candidate padding) and both channel counts: losses and every
gradient tensor.  Gross disagreement (> 5e-2 of a tensor's maximum) = a size-dependent defect; small ones are the ReLU /
max-pool discontinuities described in DESIGN.md section 4b.  usage: train_stress.py [seed]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from oracle import nsnet_oracle as oracle
from neusomatic_b200 import synth
from neusomatic_b200.train import Trainer, param_names
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dev = "cuda"
bad = 0
for C in (26, 119):
    sd = synth.random_state_dict(C, seed=seed)
    for B in (2, 7, 8, 9, 15, 16, 17, 31, 33, 100, 257, 1000):    # the library requires a batch of at least 2 (train-mode BatchNorm)
        cand = synth.structured_pileups(B, seed=seed + B, ensemble=(C == 119))
        x = torch.from_numpy(oracle.assemble_channels(cand.matrices, cand.center, cand.tumor_cov, cand.normal_cov,
                                                      cand.anns if C == 119 else None, 100)).to(dev)
        labels = torch.tensor([synth.VARTYPE_CLASSES.index(t.split(".")[4]) for t in cand.tags], dtype=torch.int32, device=dev)
        lens = torch.tensor([min(int(t.split(".")[6]), 3) for t in cand.tags], dtype=torch.int32, device=dev)
        ce = torch.from_numpy(cand.center).float().to(dev)
        w1 = torch.tensor([0.2, 0.21, 0.13, 0.16], device=dev); w3 = torch.tensor([0.12, 0.2, 0.23, 0.24], device=dev)
        res = {}
        for mode in ("fp32", "tensor"):
            os.environ.pop("NSC_TRAIN_FP32", None)
            if mode == "fp32":
                os.environ["NSC_TRAIN_FP32"] = "1"
            tr = Trainer(sd, C, max_batch=max(B, 2))
            losses = tr.forward_backward(x, labels, ce, lens, w1, w3).cpu().numpy().copy()
            res[mode] = (losses, {n: tr.grad(n).cpu().numpy().copy() for n in param_names()})
            tr.close()
        dl = float(np.abs(res["tensor"][0][0] - res["fp32"][0][0]) / max(abs(float(res["fp32"][0][0])), 1e-9))
        worst, wname = 0.0, ""
        for n in param_names():
            a, b = res["tensor"][1][n], res["fp32"][1][n]
            if not np.isfinite(a).all() or not np.isfinite(b).all():
                worst, wname = float("inf"), n
                break
            m = float(np.abs(b).max())
            if m > 1e-5:
                e = float(np.abs(a - b).max() / m)
                if e > worst:
                    worst, wname = e, n
        flag = "" if (worst <= 5e-2 and dl <= 1e-4) else "   <-- CHECK"
        bad += bool(flag)
        print("C=%3d B=%4d  loss rel diff %.1e   worst gradient tensor %-28s %.1e%s" % (C, B, dl, wname, worst, flag), flush=True)
print("flagged:", bad)
