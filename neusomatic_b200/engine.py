"""Engine: one ``nsc_model`` handle of libneusomatic_cuda on one GPU.

Host-side owner of the C-ABI handle: feeds it the reference's state_dict
(network.py:41-64 names, ``module.`` prefix tolerated as in call.py:441-460), and exposes
the forward / scoring entry points on torch CUDA tensors (device buffers) and on numpy /
pinned torch CPU tensors (host buffers).  PyTorch is only the allocator and the stream here.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

from .predtable import VARTYPE_CLASSES          # noqa: E402  ["DEL", "INS", "NONE", "SNP"], call.py:407
BN_EPS = 1e-5                                      # nn.BatchNorm2d default (network.py:24)


def anns_to_255(anns):
    """Annotation channel values as the reference stores them before ``.float()``:
    float64(a) * 255.0 (dataloader.py:395-396), cast to fp32."""
    if anns is None:
        return None
    return (np.asarray(anns, np.float64) * 255.0).astype(np.float32)


class Engine:
    def __init__(self, state_dict, num_channels, device=0, precision=None, normalize_channels=False):
        lib = _lib.load()
        self.lib = lib
        self.C = int(num_channels)
        self.device = int(device)
        h = _lib.C.c_void_p()
        _lib.check(lib.nsc_model_create(_lib.C.byref(h), self.device, self.C))
        self.h = h
        try:
            self.load_state_dict(state_dict)
            if normalize_channels:
                _lib.check(lib.nsc_model_set_normalize_channels(self.h, 1))
            if precision is not None:
                self.set_precision(precision)
        except Exception:
            self.close()
            raise

    # -- parameters -------------------------------------------------------------------------
    def load_state_dict(self, state_dict, bn_eps=BN_EPS):
        for k, v in state_dict.items():
            if k.endswith("num_batches_tracked"):
                continue
            if hasattr(v, "detach"):
                v = v.detach().to("cpu", torch.float32).contiguous().numpy()
            a = np.ascontiguousarray(v, dtype=np.float32)
            _lib.check(self.lib.nsc_model_set_param(self.h, k.encode(), a.ctypes.data, a.size))
        _lib.check(self.lib.nsc_model_finalize(self.h, bn_eps))

    def set_precision(self, precision):
        _lib.check(self.lib.nsc_model_set_precision(self.h, int(precision)))

    @property
    def precision(self):
        return self.lib.nsc_model_get_precision(self.h)

    @property
    def launch_count(self):
        return int(self.lib.nsc_model_launch_count(self.h))

    @property
    def kernel_name(self):
        return self.lib.nsc_model_kernel_name(self.h).decode()

    # -- device-buffer entry points -----------------------------------------------------------
    def _alloc_out(self, B, dev, extras):
        o1 = torch.empty((B, 4), dtype=torch.float32, device=dev)
        o2 = torch.empty((B, 1), dtype=torch.float32, device=dev)
        o3 = torch.empty((B, 4), dtype=torch.float32, device=dev)
        if not extras:
            return [o1, o2, o3, None, None, None, None]
        return [o1, o2, o3, torch.empty_like(o1), torch.empty_like(o3),
                torch.empty((B,), dtype=torch.int32, device=dev),
                torch.empty((B,), dtype=torch.int32, device=dev)]

    def forward_f32(self, x, extras=False):
        """x: CUDA fp32 [B,C,5,32] contiguous -> [o1, o2, o3, p1, p3, argmax1, argmax3]."""
        if not x.is_cuda or x.dtype != torch.float32 or not x.is_contiguous():
            raise ValueError("forward_f32 needs a contiguous fp32 CUDA tensor")
        if x.dim() != 4 or tuple(x.shape[1:]) != (self.C, 5, 32):
            raise ValueError("expected [B,%d,5,32], got %s" % (self.C, tuple(x.shape)))
        if x.device.index != self.device:
            raise ValueError("tensor on cuda:%s but engine on cuda:%d" % (x.device.index, self.device))
        B = x.shape[0]
        outs = self._alloc_out(B, x.device, extras)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(self.lib.nsc_forward_f32(self.h, x.data_ptr(), B, *[_lib.ptr(t) for t in outs], stream))
        return outs

    def forward_u8(self, matrices, meta, anns255, coverage_thr, extras=True):
        """matrices: CUDA uint8 [B,5,32,23]; meta: CUDA int32 [B,3]; anns255: CUDA fp32 [B,C-26] or None."""
        B = matrices.shape[0]
        if matrices.dtype != torch.uint8 or tuple(matrices.shape[1:]) != (5, 32, 23) or not matrices.is_contiguous():
            raise ValueError("matrices must be contiguous uint8 [B,5,32,23]")
        if meta.dtype != torch.int32 or tuple(meta.shape) != (B, 3) or not meta.is_contiguous():
            raise ValueError("meta must be contiguous int32 [B,3]")
        if anns255 is not None and (anns255.dtype != torch.float32 or tuple(anns255.shape) != (B, self.C - 26)
                                    or not anns255.is_contiguous()):
            raise ValueError("anns255 must be contiguous fp32 [B,%d]" % (self.C - 26))
        outs = self._alloc_out(B, matrices.device, extras)
        stream = torch.cuda.current_stream(matrices.device).cuda_stream
        _lib.check(self.lib.nsc_forward_u8(self.h, matrices.data_ptr(), meta.data_ptr(), _lib.ptr(anns255),
                                           float(coverage_thr), B, *[_lib.ptr(t) for t in outs], stream))
        return outs

    # -- host-buffer entry points (call.py:68-83 batch step) ------------------------------------
    @staticmethod
    def _host_out(B, extras, pin):
        def mk(shape, dt):
            t = torch.empty(shape, dtype=dt)
            return t.pin_memory() if pin else t
        outs = [mk((B, 4), torch.float32), mk((B, 1), torch.float32), mk((B, 4), torch.float32)]
        if extras:
            outs += [mk((B, 4), torch.float32), mk((B, 4), torch.float32), mk((B,), torch.int32), mk((B,), torch.int32)]
        else:
            outs += [None] * 4
        return outs

    def score_host_f32(self, x, extras=True, out=None):
        """x: host fp32 [B,C,5,32] (numpy or torch CPU, ideally pinned). Returns torch CPU tensors."""
        B = x.shape[0]
        outs = out if out is not None else self._host_out(B, extras, pin=False)
        _lib.check(self.lib.nsc_score_host_f32(self.h, _lib.ptr(x), B, *[_lib.ptr(t) for t in outs]))
        return outs

    def score_host_u8(self, matrices, meta, anns255, coverage_thr, extras=True, out=None):
        B = matrices.shape[0]
        outs = out if out is not None else self._host_out(B, extras, pin=False)
        _lib.check(self.lib.nsc_score_host_u8(self.h, _lib.ptr(matrices), _lib.ptr(meta), _lib.ptr(anns255),
                                              float(coverage_thr), B, *[_lib.ptr(t) for t in outs]))
        return outs

    def score_host_b64(self, text, spans, meta, anns255, coverage_thr, extras=True, out=None):
        """The batch step from the TSV's own image fields (base64(zlib(matrix)) text; ``CandidateTSV.stage`` fills ``text``
        and ``spans``): inflated on the device.  Returns (outs, host_decoded)."""
        B = spans.shape[0]
        outs = out if out is not None else self._host_out(B, extras, pin=False)
        nhost = C.c_int64(0)
        _lib.check(self.lib.nsc_score_host_b64(self.h, _lib.ptr(text), int(text.numel() if hasattr(text, "numel") else text.size),
                                               _lib.ptr(spans), _lib.ptr(meta), _lib.ptr(anns255), float(coverage_thr), B,
                                               *[_lib.ptr(t) for t in outs], C.byref(nhost)))
        return outs, int(nhost.value)

    # -- per-kernel timing ------------------------------------------------------------------------
    def set_profile(self, on):
        _lib.check(self.lib.nsc_model_set_profile(self.h, 1 if on else 0))

    def get_profile(self):
        """{slot name: (milliseconds, launches)} accumulated since the last call (synchronises)."""
        ms = np.zeros(16, np.float64)
        cnt = np.zeros(16, np.int64)
        _lib.check(self.lib.nsc_model_get_profile(self.h, ms.ctypes.data, cnt.ctypes.data))
        return {self.lib.nsc_profile_slot_name(i).decode(): (float(ms[i]), int(cnt[i]))
                for i in range(16) if cnt[i]}

    # -- debug -------------------------------------------------------------------------------------
    def set_debug(self, n):
        _lib.check(self.lib.nsc_model_set_debug(self.h, int(n)))

    def internal(self, which, n):
        widths = [32, 32, 16, 8, 4]
        shape = (n, 240) if which == 5 else (n, 64, 5, widths[which])
        t = torch.empty(shape, dtype=torch.float32, device="cuda:%d" % self.device)
        stream = torch.cuda.current_stream(t.device).cuda_stream
        _lib.check(self.lib.nsc_debug_internal(self.h, which, t.data_ptr(), n, stream))
        return t

    def assemble(self, matrices, meta, anns255, coverage_thr):
        """Test hook: the fp32 [B,C,5,32] tensor nsc_forward_u8 feeds the network."""
        B = matrices.shape[0]
        x = torch.empty((B, self.C, 5, 32), dtype=torch.float32, device=matrices.device)
        stream = torch.cuda.current_stream(matrices.device).cuda_stream
        _lib.check(self.lib.nsc_debug_assemble(self.h, matrices.data_ptr(), meta.data_ptr(), _lib.ptr(anns255),
                                               float(coverage_thr), B, x.data_ptr(), stream))
        return x

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.lib.nsc_model_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
