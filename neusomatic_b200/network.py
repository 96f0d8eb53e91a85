"""Drop-in for the reference's ``network.py`` (neusomatic/python/network.py:17-78).

Same class names, constructor, parameter / buffer names and shapes (so the bundled ``.pth``
files load unchanged, with or without a ``module.`` prefix -- call.py:441-460) and the same
return structure ``([o1, o2, o3], internal_outs)``.  In eval mode on a CUDA tensor the forward
pass is ONE call into libneusomatic_cuda (hand-written sm_100a kernels) through the PyTorch
custom op ``neusomatic_b200::forward_f32``; there is no CPU / eager fallback for inference.
In training mode on a CUDA tensor the forward is ``nsc_train_forward`` (training-mode BatchNorm, running statistics
updated in the module's buffers) behind a ``torch.autograd.Function`` whose backward is ``nsc_train_backward``: the
reference's loop -- ``outputs, _ = net(inputs); loss = ...; loss.backward(); optimizer.step()`` (train.py:426-441) -- runs
unmodified on the hand-written kernels, with stock ``torch.optim.SGD`` updating the module's parameters.
"""
from __future__ import annotations

import ctypes
import weakref

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .engine import Engine

_ENGINES = {}   # id -> Engine, addressed from inside the custom op by an integer key


@torch.library.custom_op("neusomatic_b200::forward_f32", mutates_args=())
def _forward_f32_op(x: torch.Tensor, engine_key: int) -> list[torch.Tensor]:
    eng = _ENGINES[engine_key]
    return eng.forward_f32(x, extras=False)[:3]


@_forward_f32_op.register_fake
def _(x, engine_key):
    B = x.shape[0]
    return [x.new_empty((B, 4)), x.new_empty((B, 1)), x.new_empty((B, 4))]


@torch.library.custom_op("neusomatic_b200::score_f32", mutates_args=())
def _score_f32_op(x: torch.Tensor, engine_key: int) -> list[torch.Tensor]:
    """Logits + the fused softmax / argmax of both classification heads (call.py:80-83,97-100) from the same launch."""
    return _ENGINES[engine_key].forward_f32(x, extras=True)


@_score_f32_op.register_fake
def _(x, engine_key):
    B = x.shape[0]
    i32 = dict(dtype=torch.int32)
    return [x.new_empty((B, 4)), x.new_empty((B, 1)), x.new_empty((B, 4)), x.new_empty((B, 4)), x.new_empty((B, 4)),
            x.new_empty((B,), **i32), x.new_empty((B,), **i32)]


class _TrainState:
    """One ``nsc_trainer`` handle of the library on one device, sized for batches of up to ``max_batch``."""

    def __init__(self, num_channels, device, max_batch):
        from .train import BN_LAYERS, param_names
        self.lib = _lib.load()
        self.dev = torch.device("cuda", device)
        h = ctypes.c_void_p()
        _lib.check(self.lib.nsc_trainer_create(ctypes.byref(h), device, num_channels, max_batch))
        self.h, self.max_batch = h, max_batch
        self.n = int(self.lib.nsc_trainer_num_params(h))
        self.names, self.bn_layers = param_names(), BN_LAYERS
        self.slices = []
        for name in self.names:
            off, num = ctypes.c_int64(), ctypes.c_int64()
            _lib.check(self.lib.nsc_trainer_param_slice(h, name.encode(), ctypes.byref(off), ctypes.byref(num)))
            self.slices.append((off.value, num.value))
        assert all(self.slices[i][0] + self.slices[i][1] == self.slices[i + 1][0] for i in range(len(self.slices) - 1))

    def close(self):
        if self.h:
            self.lib.nsc_trainer_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _TrainForward(torch.autograd.Function):
    """outputs = NeuSomaticNet(x) in training mode (network.py:66-78 with batch statistics); backward = the library's full
    backward pass.  Inputs: x, the flat running-statistics vector (updated in place by the forward), then the module's
    parameters in named_parameters() order; their gradients come back as slices of one flat vector."""

    @staticmethod
    def forward(ctx, st, momentum, x, running, *params):
        flat = torch.cat([p.detach().reshape(-1) for p in params])
        assert flat.numel() == st.n
        B = x.shape[0]
        logits = torch.empty((B, 9), dtype=torch.float32, device=x.device)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(st.lib.nsc_train_forward(st.h, flat.data_ptr(), running.data_ptr(), x.data_ptr(), B, float(momentum),
                                            logits.data_ptr(), stream))
        ctx.st, ctx.flat, ctx.x = st, flat, x
        ctx.shapes = [p.shape for p in params]
        return logits[:, 0:4].contiguous(), logits[:, 4:5].contiguous(), logits[:, 5:9].contiguous()

    @staticmethod
    def backward(ctx, g1, g2, g3):
        st, x = ctx.st, ctx.x
        B = x.shape[0]
        z = lambda g, k: torch.zeros((B, k), dtype=torch.float32, device=x.device) if g is None else g.float()
        dl = torch.cat([z(g1, 4), z(g2, 1), z(g3, 4)], 1).contiguous()
        grads = torch.empty(st.n, dtype=torch.float32, device=x.device)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(st.lib.nsc_train_backward(st.h, ctx.flat.data_ptr(), x.data_ptr(), dl.data_ptr(), B, grads.data_ptr(), stream))
        out = [grads[off:off + num].view(shape) for (off, num), shape in zip(st.slices, ctx.shapes)]
        return (None, None, None, None) + tuple(out)


class NSBlock(nn.Module):
    """network.py:17-36 (module structure only; the fused kernels implement the forward)."""

    def __init__(self, dim, ks_1=3, ks_2=3, dl_1=1, dl_2=1, mp_ks=3, mp_st=1):
        super().__init__()
        self.dim = dim
        self.conv_r1 = nn.Conv2d(dim, dim, kernel_size=ks_1, dilation=dl_1, padding=(dl_1 * (ks_1 - 1)) // 2)
        self.bn_r1 = nn.BatchNorm2d(dim)
        self.conv_r2 = nn.Conv2d(dim, dim, kernel_size=ks_2, dilation=dl_2, padding=(dl_2 * (ks_2 - 1)) // 2)
        self.bn_r2 = nn.BatchNorm2d(dim)
        self.pool_r2 = nn.MaxPool2d((1, mp_ks), padding=(0, (mp_ks - 1) // 2), stride=(1, mp_st))

    def forward(self, x):
        y1 = F.relu(self.bn_r1(self.conv_r1(x)))
        y2 = self.bn_r2(self.conv_r2(y1))
        return self.pool_r2(x + y2)


class NeuSomaticNet(nn.Module):
    def __init__(self, num_channels):
        super().__init__()
        dim = 64
        self.num_channels = num_channels
        self.conv1 = nn.Conv2d(num_channels, dim, kernel_size=(1, 3), padding=(0, 1), stride=1)
        self.bn1 = nn.BatchNorm2d(dim)
        self.pool1 = nn.MaxPool2d((1, 3), padding=(0, 1), stride=(1, 1))
        self.nsblocks = [[3, 5, 1, 1, 3, 1], [3, 5, 1, 1, 3, 2], [3, 5, 2, 1, 3, 2], [3, 5, 4, 2, 3, 2]]
        self.res_layers = nn.Sequential(*[NSBlock(dim, *spec) for spec in self.nsblocks])
        ds = int(np.prod([s[5] for s in self.nsblocks]))
        self.fc_dim = dim * 32 * 5 // ds
        self.fc1 = nn.Linear(self.fc_dim, 240)
        self.fc2 = nn.Linear(240, 4)
        self.fc3 = nn.Linear(240, 1)
        self.fc4 = nn.Linear(240, 4)
        # Shared (by reference) with nn.DataParallel replicas, whose __dict__ is a shallow copy
        # (call.py:417-419): lets a replica find the original parameters and reuse one packed
        # engine per device instead of re-packing on every forward.
        self._nsc = {"origin": weakref.ref(self), "engines": {}, "precision": None, "internals": False, "trainers": {}}

    # -- engine management ------------------------------------------------------------------
    def set_precision(self, precision):
        """NSC_PREC_* of include/neusomatic_cuda.h (None = library default)."""
        self._nsc["precision"] = precision
        for key, _sig in self._nsc["engines"].values():
            _ENGINES[key].set_precision(precision)

    def return_internals(self, on=True):
        """Also return [pool1_out, res_out, flat, fc1_out] like network.py:78 (debug; off by default
        because every reference caller discards them: call.py:77, train.py:96,430)."""
        self._nsc["internals"] = bool(on)

    def _signature(self, origin):
        return tuple((t.data_ptr(), t._version) for t in list(origin.parameters()) + list(origin.buffers()))

    def engine(self, device):
        origin = self._nsc["origin"]() or self
        idx = device.index if device.index is not None else torch.cuda.current_device()
        sig = self._signature(origin)
        ent = self._nsc["engines"].get(idx)
        if ent is not None and ent[1] == sig:
            return ent[0]
        if ent is not None:
            _ENGINES.pop(ent[0]).close()
        sd = {k: v for k, v in origin.state_dict().items()}
        eng = Engine(sd, self.num_channels, device=idx, precision=self._nsc["precision"])
        key = id(eng)
        _ENGINES[key] = eng
        self._nsc["engines"][idx] = (key, sig)
        return key

    # -- forward ------------------------------------------------------------------------------
    def _tensor_at(self, dotted):
        """The tensor behind ``dotted`` on THIS module object: a Parameter / buffer on the original module, the broadcast
        copy on an nn.DataParallel replica (whose parameters() is empty but whose attributes are set)."""
        obj = self
        for part in dotted.split("."):
            obj = obj[int(part)] if part.isdigit() else getattr(obj, part)
        return obj

    def _forward_train(self, x):
        """Training-mode forward on the library's kernels (nsc_train_forward / nsc_train_backward behind autograd)."""
        if not x.is_cuda:
            raise RuntimeError("neusomatic_b200.NeuSomaticNet trains on CUDA only (sm_100a kernels); no CPU fallback")
        x = x.contiguous().float()
        B = x.shape[0]
        if B < 2:
            raise ValueError("training-mode BatchNorm needs more than one candidate per batch (as nn.BatchNorm2d does)")
        idx = x.device.index if x.device.index is not None else torch.cuda.current_device()
        st = self._nsc["trainers"].get(idx)
        if st is None or st.max_batch < B:
            if st is not None:
                st.close()
            st = _TrainState(self.num_channels, idx, max(256, (B + 255) // 256 * 256))
            self._nsc["trainers"][idx] = st
        params = [self._tensor_at(n) for n in st.names]
        bns = [self._tensor_at(n) for n in st.bn_layers]
        running = torch.cat([t for bn in bns for t in (bn.running_mean.detach().float(), bn.running_var.detach().float())])
        outs = _TrainForward.apply(st, bns[0].momentum, x, running, *params)
        with torch.no_grad():           # the forward updated the flat copy: hand the statistics back to the module's buffers
            for i, bn in enumerate(bns):
                bn.running_mean.copy_(running[i * 128:i * 128 + 64])
                bn.running_var.copy_(running[i * 128 + 64:i * 128 + 128])
                if bn.num_batches_tracked is not None:
                    bn.num_batches_tracked += 1
        return list(outs), []

    def score(self, x):
        """[o1, o2, o3, softmax(o1), softmax(o3), argmax(o1), argmax(o3)] of a CUDA batch in one library call: what
        call_variants reads per batch (call.py:77-83, 97-100)."""
        if not x.is_cuda:
            raise RuntimeError("neusomatic_b200.NeuSomaticNet scores on CUDA only (sm_100a kernels); no CPU fallback")
        x = x.contiguous()
        if x.dtype != torch.float32:
            x = x.float()
        return _score_f32_op(x, self.engine(x.device))

    def forward(self, x):
        if self.training:
            return self._forward_train(x)
        if not x.is_cuda:
            raise RuntimeError("neusomatic_b200.NeuSomaticNet scores on CUDA only (sm_100a kernels); "
                               "there is no CPU fallback -- move the batch to the GPU")
        x = x.contiguous()
        if x.dtype != torch.float32:
            x = x.float()
        key = self.engine(x.device)
        if self._nsc["internals"]:
            _ENGINES[key].set_debug(x.shape[0])
        outs = _forward_f32_op(x, key)
        internals = []
        if self._nsc["internals"]:
            eng = _ENGINES[key]
            internals = [eng.internal(0, x.shape[0]), eng.internal(4, x.shape[0])]
            internals += [internals[1].reshape(-1, self.fc_dim), eng.internal(5, x.shape[0])]
        return list(outs), internals
