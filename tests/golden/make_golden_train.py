#!/usr/bin/env python
"""Golden vectors for the train step from the UNMODIFIED reference (run in the build container only).

Imports /root/reference/neusomatic/python/network.py, puts the net in train() mode, runs
loss = CE(w_type) + SmoothL1 + CE(w_len) exactly as train.py:379-383,436-441 on a seeded synthetic batch,
loss.backward(), optim.SGD(lr=0.01, momentum=0.9).step() twice, and stores inputs, losses, a few gradients,
updated parameters and BN running statistics in tests/golden/train_std.npz."""
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.optim as optim

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/neusomatic/python")
from network import NeuSomaticNet        # noqa: E402  (reference)
from neusomatic_b200 import synth         # noqa: E402
from oracle import nsnet_oracle as oracle  # noqa: E402


def main():
    torch.manual_seed(7)
    torch.set_num_threads(1)
    B = 48
    cand = synth.structured_pileups(B, seed=31)
    x = torch.from_numpy(oracle.assemble_channels(cand.matrices, cand.center, cand.tumor_cov, cand.normal_cov, None, 100))
    vt = torch.tensor([synth.VARTYPE_CLASSES.index(t.split(".")[4]) for t in cand.tags])
    length = torch.tensor([int(t.split(".")[6]) for t in cand.tags])
    centers = torch.tensor(cand.center, dtype=torch.float32)
    len_labels = torch.clamp(length, max=3)
    w_type = torch.tensor([0.2, 0.21, 0.13, 0.16])
    w_len = torch.tensor([0.12, 0.2, 0.23, 0.24])
    sd0 = {k: torch.from_numpy(v) for k, v in synth.random_state_dict(26, seed=11).items()}
    net = NeuSomaticNet(26)
    net.load_state_dict(sd0, strict=False)
    net.train()
    c1, c2, c3 = nn.CrossEntropyLoss(w_type), nn.SmoothL1Loss(), nn.CrossEntropyLoss(w_len)
    opt = optim.SGD(net.parameters(), lr=0.01, momentum=0.9)
    out = dict(matrices=cand.matrices, center=cand.center, tumor_cov=cand.tumor_cov, normal_cov=cand.normal_cov,
               labels=vt.numpy().astype(np.int32), len_labels=len_labels.numpy().astype(np.int32),
               w_type=w_type.numpy(), w_len=w_len.numpy())
    for step in range(2):
        opt.zero_grad()
        (o1, o2, o3), _ = net(x)
        l1, l2, l3 = c1(o1, vt), c2(o2.squeeze(1), centers), c3(o3, len_labels)
        loss = l1 + l2 + l3
        loss.backward()
        out["loss%d" % step] = np.float32([loss.item(), l1.item(), l2.item(), l3.item()])
        if step == 0:
            for n, prm in net.named_parameters():
                if n in ("conv1.weight", "bn1.weight", "res_layers.0.conv_r2.weight", "res_layers.3.conv_r1.weight",
                         "res_layers.3.bn_r2.bias", "res_layers.2.conv_r2.bias", "fc1.weight", "fc3.bias", "fc4.weight"):
                    out["grad." + n] = prm.grad.numpy().copy()
        opt.step()
    for n, t in net.state_dict().items():
        if n.endswith("num_batches_tracked"):
            continue
        if n.startswith(("conv1", "bn1", "res_layers.1.", "res_layers.3.bn_r2", "fc2", "fc1.bias")):
            out["final." + n] = t.numpy().copy()
    np.savez_compressed(os.path.join(HERE, "train_std.npz"), **out)
    print("losses", out["loss0"], out["loss1"])


if __name__ == "__main__":
    main()
