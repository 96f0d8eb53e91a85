"""CPU: for the golden training batch, lists the BatchNorm outputs closest to the ReLU kink in each NSBlock (float64
autograd oracle) and how far flipping that one mask element would move the layer's bias gradient.  This is the
conditioning of the gradient parity test (tests/test_gpu_train.py): an implementation whose forward error exceeds |z1|
flips the element and shows exactly the listed deviation."""
import os, sys, numpy as np, torch, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from oracle import nsnet_oracle as oracle, train_port as tp
from neusomatic_b200 import synth
g = dict(np.load(os.path.join(ROOT, "tests", "golden", "train_std.npz")))
x = torch.from_numpy(oracle.assemble_channels(g["matrices"], g["center"], g["tumor_cov"], g["normal_cov"], None, 100)).double()
sd = synth.random_state_dict(26, seed=11)
p = {k: torch.from_numpy(v).double() for k, v in sd.items()}
for n in tp.param_names(): p[n].requires_grad_(True)
keep = {}
def bn(t, name):
    return F.batch_norm(t, None, None, p[name + ".weight"], p[name + ".bias"], training=True, eps=1e-5)
h = F.max_pool2d(F.relu(bn(F.conv2d(x, p["conv1.weight"], p["conv1.bias"], padding=(0, 1)), "bn1")), (1, 3), stride=(1, 1), padding=(0, 1))
for i, (k1, k2, d1, d2, st) in enumerate(tp.NSBLOCKS):
    pre = "res_layers.%d." % i
    z1 = bn(F.conv2d(h, p[pre + "conv_r1.weight"], p[pre + "conv_r1.bias"], dilation=d1, padding=(d1 * (k1 - 1)) // 2), pre + "bn_r1")
    y1 = F.relu(z1); y1.retain_grad(); keep[i] = (z1, y1)
    y2 = bn(F.conv2d(y1, p[pre + "conv_r2.weight"], p[pre + "conv_r2.bias"], dilation=d2, padding=(d2 * (k2 - 1)) // 2), pre + "bn_r2")
    h = F.max_pool2d(h + y2, (1, 3), stride=(1, st), padding=(0, 1))
x3 = F.relu(F.linear(h.reshape(h.shape[0], -1), p["fc1.weight"], p["fc1.bias"]))
o = (F.linear(x3, p["fc2.weight"], p["fc2.bias"]), F.linear(x3, p["fc3.weight"], p["fc3.bias"]), F.linear(x3, p["fc4.weight"], p["fc4.bias"]))
loss, _ = tp.loss_fn(*o, torch.from_numpy(g["labels"]).long(), torch.from_numpy(g["center"]).double(), torch.from_numpy(g["len_labels"]).long(), torch.from_numpy(g["w_type"]).double(), torch.from_numpy(g["w_len"]).double())
loss.backward()
for i in range(4):
    z1, y1 = keep[i]
    gb = p["res_layers.%d.bn_r1.bias" % i].grad
    a = z1.detach().abs().flatten()
    idx = torch.argsort(a)[:5]
    print("block", i, "max|g bn_r1.bias| %.3e" % gb.abs().max().item())
    for j in idx:
        print("   |z1| = %.2e   dL/dy1 = %.3e   (ratio to max bias grad %.2e)" % (a[j].item(), y1.grad.flatten()[j].item(), abs(y1.grad.flatten()[j].item()) / gb.abs().max().item()))


# ---- max-pool windows whose two largest values nearly tie (the other discontinuity of the backward pass) ----------------
if "--pools" in sys.argv:
    def near_ties(t, st, name):
        W = t.shape[-1]
        pad = F.pad(t.detach(), (1, 1), value=float("-inf"))
        win = torch.stack([pad[..., 0:W:st], pad[..., 1:W + 1:st], pad[..., 2:W + 2:st]], -1)
        top2 = win.topk(2, dim=-1).values
        gap = top2[..., 0] - top2[..., 1]
        idx = torch.nonzero((gap > 0) & (gap < 3e-6 * t.detach().abs().max()))
        print("%-8s pool windows with 0 < (max - second) < 3e-6 * max|z|: %d   [sample, channel, row, output column]: %s" % (
            name, len(idx), idx[:8].tolist()))
    h2 = F.max_pool2d(F.relu(bn(F.conv2d(x, p["conv1.weight"], p["conv1.bias"], padding=(0, 1)), "bn1")), (1, 3), stride=(1, 1), padding=(0, 1))
    for i, (k1, k2, d1, d2, st) in enumerate(tp.NSBLOCKS):
        pre = "res_layers.%d." % i
        y1 = F.relu(bn(F.conv2d(h2, p[pre + "conv_r1.weight"], p[pre + "conv_r1.bias"], dilation=d1, padding=(d1 * (k1 - 1)) // 2), pre + "bn_r1"))
        z = h2 + bn(F.conv2d(y1, p[pre + "conv_r2.weight"], p[pre + "conv_r2.bias"], dilation=d2, padding=(d2 * (k2 - 1)) // 2), pre + "bn_r2")
        near_ties(z, st, "block %d" % i)
        h2 = F.max_pool2d(z, (1, 3), stride=(1, st), padding=(0, 1))
