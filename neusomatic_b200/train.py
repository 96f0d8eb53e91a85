"""Host-side mirror of the optimisation step of the reference's train.py (train.py:379-383, 416-441) on top of
``nsc_train_forward_backward`` / ``nsc_sgd_momentum_update`` (csrc/nsc_train.cu).  Parameters, gradients and
momentum live in flat torch CUDA tensors (PyTorch = allocator + NCCL plumbing); in data-parallel runs the one
exchange step of the path is a single all-reduce of the flat gradient vector."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

BN_LAYERS = ["bn1"] + ["res_layers.%d.%s" % (i, l) for i in range(4) for l in ("bn_r1", "bn_r2")]


def param_names():
    n = ["conv1.weight", "conv1.bias", "bn1.weight", "bn1.bias"]
    for i in range(4):
        for l in ("conv_r1", "bn_r1", "conv_r2", "bn_r2"):
            n += ["res_layers.%d.%s.weight" % (i, l), "res_layers.%d.%s.bias" % (i, l)]
    for f in ("fc1", "fc2", "fc3", "fc4"):
        n += [f + ".weight", f + ".bias"]
    return n


class Trainer:
    def __init__(self, state_dict, num_channels, device=0, max_batch=1024, lr=0.01, momentum=0.9, bn_momentum=0.1):
        self.lib = _lib.load()
        self.dev = torch.device("cuda", device)
        h = C.c_void_p()
        _lib.check(self.lib.nsc_trainer_create(C.byref(h), device, num_channels, max_batch))
        self.h = h
        self.C, self.lr, self.momentum, self.bn_momentum = num_channels, lr, momentum, bn_momentum
        self.n = int(self.lib.nsc_trainer_num_params(h))
        self.slices = {}
        for name in param_names():
            off, num = C.c_int64(), C.c_int64()
            _lib.check(self.lib.nsc_trainer_param_slice(h, name.encode(), C.byref(off), C.byref(num)))
            self.slices[name] = (off.value, num.value)
        self.params = torch.zeros(self.n, dtype=torch.float32, device=self.dev)
        self.grads = torch.zeros_like(self.params)
        self.mom = torch.zeros_like(self.params)
        self.running = torch.zeros(9 * 128, dtype=torch.float32, device=self.dev)
        self.losses = torch.zeros(4, dtype=torch.float32, device=self.dev)
        self.load_state_dict(state_dict)

    def load_state_dict(self, sd):
        sd = {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}
        for name, (off, num) in self.slices.items():
            self.params[off:off + num] = torch.as_tensor(np.asarray(sd[name]), dtype=torch.float32).reshape(-1).to(self.dev)
        for i, bn in enumerate(BN_LAYERS):
            self.running[i * 128:i * 128 + 64] = torch.as_tensor(np.asarray(sd[bn + ".running_mean"])).to(self.dev)
            self.running[i * 128 + 64:i * 128 + 128] = torch.as_tensor(np.asarray(sd[bn + ".running_var"])).to(self.dev)

    def state_dict(self):
        """Reference checkpoint layout (train.py:390-395 ``state_dict``), CPU tensors."""
        shapes = {"conv1.weight": (64, self.C, 1, 3), "fc1.weight": (240, 1280), "fc2.weight": (4, 240), "fc3.weight": (1, 240),
                  "fc4.weight": (4, 240)}
        ks = {0: (3, 5), 1: (3, 5), 2: (3, 5), 3: (3, 5)}
        out = {}
        for name, (off, num) in self.slices.items():
            t = self.params[off:off + num].cpu()
            if name in shapes:
                t = t.reshape(shapes[name])
            elif name.endswith("conv_r1.weight") or name.endswith("conv_r2.weight"):
                i = int(name.split(".")[1])
                k = ks[i][0] if "conv_r1" in name else ks[i][1]
                t = t.reshape(64, 64, k, k)
            out[name] = t.clone()
        for i, bn in enumerate(BN_LAYERS):
            out[bn + ".running_mean"] = self.running[i * 128:i * 128 + 64].cpu().clone()
            out[bn + ".running_var"] = self.running[i * 128 + 64:i * 128 + 128].cpu().clone()
        return out

    def debug_tensor(self, name, shape):
        """Copy of an internal fp32 NCHW tensor of the last step (nsc_trainer_debug_tensor), as a CUDA tensor of `shape`."""
        ptr = C.c_void_p()
        _lib.check(self.lib.nsc_trainer_debug_tensor(self.h, name.encode(), C.byref(ptr)))
        class _View:      # zero-copy view of library-owned device memory through the CUDA array interface
            __cuda_array_interface__ = {"shape": tuple(int(d) for d in shape), "typestr": "<f4", "data": (int(ptr.value), False),
                                        "version": 2}
        torch.cuda.current_stream(self.dev).synchronize()
        return torch.as_tensor(_View(), device=self.dev).clone()

    def grad(self, name):
        off, num = self.slices[name]
        return self.grads[off:off + num]

    def _check(self, name, t, dtype, shape):
        """The C ABI takes raw pointers: a wrong dtype (e.g. PyTorch's default int64 labels), device, stride or shape would
        be read as garbage.  Label VALUES are range-checked on the device (a label outside [0,4) makes the step's losses NaN
        and the next call fail; the reference's CrossEntropyLoss raises)."""
        if not isinstance(t, torch.Tensor) or t.device != self.dev:
            raise ValueError("%s must be a tensor on %s" % (name, self.dev))
        if t.dtype != dtype:
            raise ValueError("%s must be %s, got %s" % (name, dtype, t.dtype))
        if tuple(t.shape) != tuple(shape):
            raise ValueError("%s must have shape %s, got %s" % (name, tuple(shape), tuple(t.shape)))
        if not t.is_contiguous():
            raise ValueError("%s must be contiguous" % name)

    def forward_backward(self, x, labels, centers, len_labels, w_type, w_len):
        """x [B,C,5,32] fp32 CUDA; labels / len_labels int32 CUDA [B]; centers fp32 CUDA [B]; class weights CUDA [4]."""
        B = x.shape[0]
        self._check("x", x, torch.float32, (B, self.C, 5, 32))
        self._check("labels", labels, torch.int32, (B,))
        self._check("centers", centers, torch.float32, (B,))
        self._check("len_labels", len_labels, torch.int32, (B,))
        self._check("w_type", w_type, torch.float32, (4,))
        self._check("w_len", w_len, torch.float32, (4,))
        stream = torch.cuda.current_stream(self.dev).cuda_stream
        _lib.check(self.lib.nsc_train_forward_backward(
            self.h, self.params.data_ptr(), self.running.data_ptr(), x.data_ptr(), labels.data_ptr(), centers.data_ptr(),
            len_labels.data_ptr(), w_type.data_ptr(), w_len.data_ptr(), B, self.bn_momentum, self.grads.data_ptr(),
            self.losses.data_ptr(), stream))
        return self.losses

    def step(self, x, labels, centers, len_labels, w_type, w_len, group=None):
        """One train.py iteration (train.py:426-441); with torch.distributed initialised the per-rank gradients of the
        GLOBAL loss are summed over ranks (NCCL) -- the only exchange step of the path.  BatchNorm uses the statistics of
        the rank's own shard and each rank keeps its own running statistics, as the replicas of the reference's
        nn.DataParallel do (rank 0's are the ones a checkpoint would hold, like device 0's there).  The returned losses
        are this rank's share of the global loss (their sum over ranks is the reference's loss)."""
        import torch.distributed as dist
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        if multi:
            # the reference's nn.DataParallel computes one loss over the global batch (train.py:436-441): normalise both
            # weighted cross entropies and SmoothL1 by the GLOBAL weight sums / batch size (3 floats all-reduced before the
            # step; the labels are known up front), then SUM the per-rank gradients
            self._norm = torch.stack([w_type[labels.long()].sum(), w_len[len_labels.long()].sum(),
                                      torch.tensor(float(x.shape[0]), device=self.dev)]).float()
            dist.all_reduce(self._norm, op=dist.ReduceOp.SUM, group=group)
            _lib.check(self.lib.nsc_trainer_set_loss_norm(self.h, self._norm.data_ptr()))
        try:
            losses = self.forward_backward(x, labels, centers, len_labels, w_type, w_len)
        finally:
            if multi:
                _lib.check(self.lib.nsc_trainer_set_loss_norm(self.h, None))
        scale = 1.0
        if multi:
            dist.all_reduce(self.grads, op=dist.ReduceOp.SUM, group=group)
        stream = torch.cuda.current_stream(self.dev).cuda_stream
        _lib.check(self.lib.nsc_sgd_momentum_update(self.params.data_ptr(), self.mom.data_ptr(), self.grads.data_ptr(), self.n,
                                                    self.lr, self.momentum, scale, stream))
        return losses

    def close(self):
        if getattr(self, "h", None):
            self.lib.nsc_trainer_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
