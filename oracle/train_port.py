"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

Restatement of one optimisation step of the reference's train.py (train.py:379-383, 416-441): forward in
training mode (BatchNorm2d with batch statistics and running-stat update, network.py:17-78), the loss
    CrossEntropyLoss(w_type)(o1, labels) + SmoothL1Loss()(o2[:,0], center) + CrossEntropyLoss(w_len)(o3, len_labels)
``loss.backward()`` and SGD(lr, momentum) -- on the PyTorch CPU operators / autograd the reference itself uses.
Parameters are a dict in the reference's state_dict naming.  Pinned against the unmodified reference
(tests/golden/make_golden_train.py -> tests/golden/train_std.npz).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

NSBLOCKS = [(3, 5, 1, 1, 1), (3, 5, 1, 1, 2), (3, 5, 2, 1, 2), (3, 5, 4, 2, 2)]   # network.py:48-53
PARAM_ORDER_TAIL = ["fc1", "fc2", "fc3", "fc4"]


def param_names():
    """Trainable tensors in nn.Module.named_parameters() order (network.py:41-64)."""
    n = ["conv1.weight", "conv1.bias", "bn1.weight", "bn1.bias"]
    for i in range(4):
        for l in ("conv_r1", "bn_r1", "conv_r2", "bn_r2"):
            n += ["res_layers.%d.%s.weight" % (i, l), "res_layers.%d.%s.bias" % (i, l)]
    for f in PARAM_ORDER_TAIL:
        n += [f + ".weight", f + ".bias"]
    return n


def bn_names():
    return ["bn1"] + ["res_layers.%d.%s" % (i, l) for i in range(4) for l in ("bn_r1", "bn_r2")]


def forward_train(p, x, bn_momentum=0.1):
    """Training-mode forward; updates p[<bn>.running_mean/var] in place.  p: dict of fp32 CPU tensors."""
    def bn(t, name):
        return F.batch_norm(t, p[name + ".running_mean"], p[name + ".running_var"], p[name + ".weight"], p[name + ".bias"],
                            training=True, momentum=bn_momentum, eps=1e-5)
    x = F.max_pool2d(F.relu(bn(F.conv2d(x, p["conv1.weight"], p["conv1.bias"], padding=(0, 1)), "bn1")),
                     (1, 3), stride=(1, 1), padding=(0, 1))
    for i, (k1, k2, d1, d2, st) in enumerate(NSBLOCKS):
        pre = "res_layers.%d." % i
        y1 = F.relu(bn(F.conv2d(x, p[pre + "conv_r1.weight"], p[pre + "conv_r1.bias"], dilation=d1, padding=(d1 * (k1 - 1)) // 2),
                       pre + "bn_r1"))
        y2 = bn(F.conv2d(y1, p[pre + "conv_r2.weight"], p[pre + "conv_r2.bias"], dilation=d2, padding=(d2 * (k2 - 1)) // 2),
                pre + "bn_r2")
        x = F.max_pool2d(x + y2, (1, 3), stride=(1, st), padding=(0, 1))
    x3 = F.relu(F.linear(x.reshape(x.shape[0], -1), p["fc1.weight"], p["fc1.bias"]))
    return (F.linear(x3, p["fc2.weight"], p["fc2.bias"]), F.linear(x3, p["fc3.weight"], p["fc3.bias"]),
            F.linear(x3, p["fc4.weight"], p["fc4.bias"]))


def loss_fn(o1, o2, o3, labels, centers, len_labels, w_type, w_len):
    """train.py:436-438."""
    l1 = F.cross_entropy(o1, labels, weight=w_type)
    l2 = F.smooth_l1_loss(o2.squeeze(1), centers)
    l3 = F.cross_entropy(o3, len_labels, weight=w_len)
    return l1 + l2 + l3, (l1, l2, l3)


def grads(p, x, labels, centers, len_labels, w_type, w_len):
    """Returns (loss, (l1,l2,l3), {name: grad}) and updates the BN running statistics in p."""
    names = param_names()
    for n in names:
        p[n] = p[n].detach().clone().requires_grad_(True)
    o1, o2, o3 = forward_train(p, x)
    loss, parts = loss_fn(o1, o2, o3, labels, centers, len_labels, w_type, w_len)
    g = torch.autograd.grad(loss, [p[n] for n in names])
    for n in names:
        p[n] = p[n].detach()
    return loss.detach(), tuple(t.detach() for t in parts), dict(zip(names, g))


def sgd_step(p, g, mom, lr, momentum):
    """optim.SGD(lr, momentum) (train.py:382-383): buf = momentum*buf + grad; p -= lr*buf."""
    for n, grad in g.items():
        buf = mom.get(n)
        buf = grad.clone() if buf is None else buf.mul(momentum).add(grad)
        mom[n] = buf
        p[n] = p[n] - lr * buf
