"""GPU: gradient error per tensor (against the float64 autograd oracle) of one training step for each subset of the
tensor-path convolution families (NSC_TRAIN_TENSOR), on a small synthetic batch.  usage: train_mode_probe.py [C] [B]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from oracle import nsnet_oracle as oracle, train_port
from neusomatic_b200 import synth
from neusomatic_b200.train import Trainer
C = int(sys.argv[1]) if len(sys.argv) > 1 else 26
B = int(sys.argv[2]) if len(sys.argv) > 2 else 13
cand = synth.structured_pileups(B, seed=5, ensemble=(C == 119))
x = oracle.assemble_channels(cand.matrices, cand.center, cand.tumor_cov, cand.normal_cov, cand.anns if C == 119 else None, 100)
labels = np.array([synth.VARTYPE_CLASSES.index(t.split(".")[4]) for t in cand.tags], np.int32)
lens = np.array([min(int(t.split(".")[6]), 3) for t in cand.tags], np.int32)
w1 = np.array([0.2, 0.21, 0.13, 0.16], np.float32); w3 = np.array([0.12, 0.2, 0.23, 0.24], np.float32)
sd = synth.random_state_dict(C, seed=3)
p = {k: torch.from_numpy(v).double() for k, v in sd.items()}
_, _, ref = train_port.grads(p, torch.from_numpy(x).double(), torch.from_numpy(labels).long(), torch.from_numpy(cand.center).double(),
                             torch.from_numpy(lens).long(), torch.from_numpy(w1).double(), torch.from_numpy(w3).double())
dev = "cuda"
for mode in ("", "fwd", "bwd", "wgrad", "fwdbwd", "fwdwgrad", "bwdwgrad", "fwdbwdwgrad"):
    os.environ["NSC_TRAIN_TENSOR"] = mode if mode else "none"
    tr = Trainer(sd, C, max_batch=max(B, 32))
    tr.forward_backward(torch.from_numpy(x).to(dev), torch.from_numpy(labels).to(dev), torch.from_numpy(cand.center).float().to(dev),
                        torch.from_numpy(lens).to(dev), torch.from_numpy(w1).to(dev), torch.from_numpy(w3).to(dev))
    worst = {}
    for name, r in ref.items():
        if float(r.abs().max()) > 1e-5 and name.endswith("weight") and "conv" in name:
            got = tr.grad(name).cpu().numpy().reshape(r.shape)
            worst[name.replace("res_layers.", "rl").replace(".weight", "")] = "%.0e" % float(np.abs(got - r.numpy()).max() / float(r.abs().max()))
    print("%-12s" % (mode or "fp32"), worst, flush=True)
    tr.close()
