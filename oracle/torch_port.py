"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

Restatement of NeuSomaticNet.forward (reference neusomatic/python/network.py:66-78, NSBlock
network.py:31-36) with the same third-party CPU operators the reference itself executes
(torch.nn.functional conv2d / batch_norm / max_pool2d / linear on CPU fp32): the reference's
arithmetic lives in PyTorch (README.md:43 pins 1.1.0; this image has 2.11), not in the repo.
Used as the timed CPU baseline of bench.py (``cpu_baseline`` and ``--impl reference``) and
pinned against tests/golden in tests/test_oracle.py.  Only tests/, smoke() and bench.py's CPU
legs may import it.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

# (ks_1, ks_2, dl_1, dl_2, mp_st)   network.py:48-53
NSBLOCKS = [(3, 5, 1, 1, 1), (3, 5, 1, 1, 2), (3, 5, 2, 1, 2), (3, 5, 4, 2, 2)]


def prepare(state_dict):
    """Strip ``module.`` (call.py:446-455), keep fp32 CPU tensors."""
    out = {}
    for k, v in state_dict.items():
        if k.startswith("module."):
            k = k[7:]
        if k.endswith("num_batches_tracked"):
            continue
        out[k] = torch.as_tensor(v).detach().to("cpu", torch.float32).contiguous()
    return out


def _bn(x, p, prefix):
    return F.batch_norm(x, p[prefix + ".running_mean"], p[prefix + ".running_var"], p[prefix + ".weight"],
                        p[prefix + ".bias"], training=False, eps=1e-5)


@torch.no_grad()
def forward(p, x):
    """x: fp32 CPU tensor [B,C,5,32] -> (o1 [B,4], o2 [B,1], o3 [B,4])."""
    x = F.max_pool2d(F.relu(_bn(F.conv2d(x, p["conv1.weight"], p["conv1.bias"], padding=(0, 1)), p, "bn1")),
                     (1, 3), stride=(1, 1), padding=(0, 1))
    for i, (k1, k2, d1, d2, st) in enumerate(NSBLOCKS):
        pre = "res_layers.%d." % i
        y1 = F.relu(_bn(F.conv2d(x, p[pre + "conv_r1.weight"], p[pre + "conv_r1.bias"], dilation=d1,
                                 padding=(d1 * (k1 - 1)) // 2), p, pre + "bn_r1"))
        y2 = _bn(F.conv2d(y1, p[pre + "conv_r2.weight"], p[pre + "conv_r2.bias"], dilation=d2,
                          padding=(d2 * (k2 - 1)) // 2), p, pre + "bn_r2")
        x = F.max_pool2d(x + y2, (1, 3), stride=(1, st), padding=(0, 1))
    x3 = F.relu(F.linear(x.reshape(x.shape[0], -1), p["fc1.weight"], p["fc1.bias"]))
    return (F.linear(x3, p["fc2.weight"], p["fc2.bias"]), F.linear(x3, p["fc3.weight"], p["fc3.bias"]),
            F.linear(x3, p["fc4.weight"], p["fc4.bias"]))
