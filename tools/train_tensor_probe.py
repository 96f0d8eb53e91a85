"""GPU: compares the internal forward tensors of one training step (nsc_trainer_debug_tensor) with the float64 autograd
oracle restated layer by layer, and the final backward scratch tensors.  usage: train_tensor_probe.py [B] [mode]"""
import os, sys
import numpy as np, torch, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from oracle import nsnet_oracle as oracle, train_port as tp
from neusomatic_b200 import synth
from neusomatic_b200.train import Trainer
B = int(sys.argv[1]) if len(sys.argv) > 1 else 13
if len(sys.argv) > 2: os.environ["NSC_TRAIN_TENSOR"] = sys.argv[2]
C = 26
cand = synth.structured_pileups(B, seed=5)
x = oracle.assemble_channels(cand.matrices, cand.center, cand.tumor_cov, cand.normal_cov, None, 100)
labels = np.array([synth.VARTYPE_CLASSES.index(t.split(".")[4]) for t in cand.tags], np.int32)
lens = np.array([min(int(t.split(".")[6]), 3) for t in cand.tags], np.int32)
w1 = np.array([0.2, 0.21, 0.13, 0.16], np.float32); w3 = np.array([0.12, 0.2, 0.23, 0.24], np.float32)
sd = synth.random_state_dict(C, seed=3)
p = {k: torch.from_numpy(v).double() for k, v in sd.items()}
for n in tp.param_names(): p[n].requires_grad_(True)
ref = {}
def bn(t, name): return F.batch_norm(t, None, None, p[name + ".weight"], p[name + ".bias"], training=True, eps=1e-5)
xd = torch.from_numpy(x).double()
ref["u0"] = F.conv2d(xd, p["conv1.weight"], p["conv1.bias"], padding=(0, 1))
ref["a0"] = F.relu(bn(ref["u0"], "bn1"))
h = F.max_pool2d(ref["a0"], (1, 3), stride=(1, 1), padding=(0, 1)); ref["x0"] = h
for i, (k1, k2, d1, d2, st) in enumerate(tp.NSBLOCKS):
    pre = "res_layers.%d." % i
    ref["u1.%d" % i] = F.conv2d(h, p[pre + "conv_r1.weight"], p[pre + "conv_r1.bias"], dilation=d1, padding=(d1 * (k1 - 1)) // 2)
    ref["a1.%d" % i] = F.relu(bn(ref["u1.%d" % i], pre + "bn_r1"))
    ref["u2.%d" % i] = F.conv2d(ref["a1.%d" % i], p[pre + "conv_r2.weight"], p[pre + "conv_r2.bias"], dilation=d2, padding=(d2 * (k2 - 1)) // 2)
    ref["z.%d" % i] = h + bn(ref["u2.%d" % i], pre + "bn_r2")
    h = F.max_pool2d(ref["z.%d" % i], (1, 3), stride=(1, st), padding=(0, 1)); ref["x%d" % (i + 1)] = h
for t in ref.values(): t.retain_grad()
x3 = F.relu(F.linear(h.reshape(h.shape[0], -1), p["fc1.weight"], p["fc1.bias"]))
o = (F.linear(x3, p["fc2.weight"], p["fc2.bias"]), F.linear(x3, p["fc3.weight"], p["fc3.bias"]), F.linear(x3, p["fc4.weight"], p["fc4.bias"]))
loss, _ = tp.loss_fn(*o, torch.from_numpy(labels).long(), torch.from_numpy(cand.center).double(), torch.from_numpy(lens).long(),
                     torch.from_numpy(w1).double(), torch.from_numpy(w3).double())
loss.backward()
dev = "cuda"
tr = Trainer(sd, C, max_batch=max(B, 32))
tr.forward_backward(torch.from_numpy(x).to(dev), torch.from_numpy(labels).to(dev), torch.from_numpy(cand.center).float().to(dev),
                    torch.from_numpy(lens).to(dev), torch.from_numpy(w1).to(dev), torch.from_numpy(w3).to(dev))
for name, r in ref.items():
    got = tr.debug_tensor(name, tuple(r.shape)).double().cpu()
    d = (got - r.detach()).abs()
    print("%-6s max|d| %.2e (max|ref| %.2e)  elements off by > 1e-4 of max: %d" % (name, d.max(), r.abs().max(), int((d > 1e-4 * r.abs().max()).sum())))
# final backward scratch: gu = du0 (gradient of conv1's output), gz = gradient of a0 after the pool backward
for name, r in (("gu", ref["u0"].grad), ("gz", ref["a0"].grad)):
    got = tr.debug_tensor(name, tuple(r.shape)).double().cpu()
    if name == "gz": r = r.clone()      # the trainer's gz is before the ReLU mask; the autograd a0.grad too
    d = (got - r).abs()
    idx = torch.nonzero(d > 1e-3 * r.abs().max())
    print("%-6s max|d| %.2e (max|ref| %.2e)  elements off by > 1e-3 of max: %d  first: %s" % (name, d.max(), r.abs().max(), len(idx), idx[:6].tolist()))
