"""ctypes binding of libneusomatic_cuda.so (the C ABI of include/neusomatic_cuda.h).

The library is the product: if it is missing or cannot be loaded this module raises --
there is no PyTorch / CPU fallback behind it.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libneusomatic_cuda.so")

PREC_SPLIT_BF16 = 0
PREC_BF16 = 1
PREC_FP32 = 2
PREC_SPLIT8 = 3

_vp, _i64, _i32, _f32 = C.c_void_p, C.c_int64, C.c_int, C.c_float
_OUT7 = [_vp] * 7   # type_logits, pos, len_logits, type_prob, len_prob, type_argmax, len_argmax

# name -> (restype, argtypes): every symbol include/neusomatic_cuda.h declares
SIGNATURES = {
    "nsc_abi_version": (_i32, []),
    "nsc_last_error": (C.c_char_p, []),
    "nsc_model_create": (_i32, [C.POINTER(_vp), _i32, _i32]),
    "nsc_model_set_param": (_i32, [_vp, C.c_char_p, _vp, _i64]),
    "nsc_model_finalize": (_i32, [_vp, _f32]),
    "nsc_model_set_precision": (_i32, [_vp, _i32]),
    "nsc_model_get_precision": (_i32, [_vp]),
    "nsc_model_set_normalize_channels": (_i32, [_vp, _i32]),
    "nsc_model_set_debug": (_i32, [_vp, _i64]),
    "nsc_forward_f32": (_i32, [_vp, _vp, _i64] + _OUT7 + [_vp]),
    "nsc_forward_u8": (_i32, [_vp, _vp, _vp, _vp, _f32, _i64] + _OUT7 + [_vp]),
    "nsc_score_host_f32": (_i32, [_vp, _vp, _i64] + _OUT7),
    "nsc_score_host_u8": (_i32, [_vp, _vp, _vp, _vp, _f32, _i64] + _OUT7),
    "nsc_score_host_b64": (_i32, [_vp, _vp, _i64, _vp, _vp, _vp, _f32, _i64] + _OUT7 + [C.POINTER(_i64)]),
    "nsc_debug_internal": (_i32, [_vp, _i32, _vp, _i64, _vp]),
    "nsc_debug_assemble": (_i32, [_vp, _vp, _vp, _vp, _f32, _i64, _vp, _vp]),
    "nsc_model_set_profile": (_i32, [_vp, _i32]),
    "nsc_model_get_profile": (_i32, [_vp, _vp, _vp]),
    "nsc_profile_slot_name": (C.c_char_p, [_i32]),
    "nsc_model_launch_count": (_i64, [_vp]),
    "nsc_model_kernel_name": (C.c_char_p, [_vp]),
    "nsc_model_destroy": (_i32, [_vp]),
    "nsc_trainer_create": (_i32, [C.POINTER(_vp), _i32, _i32, _i64]),
    "nsc_trainer_num_params": (_i64, [_vp]),
    "nsc_trainer_param_slice": (_i32, [_vp, C.c_char_p, C.POINTER(_i64), C.POINTER(_i64)]),
    "nsc_train_forward_backward": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _f32, _vp, _vp, _vp]),
    "nsc_trainer_set_loss_norm": (_i32, [_vp, _vp]),
    "nsc_train_forward": (_i32, [_vp, _vp, _vp, _vp, _i64, _f32, _vp, _vp]),
    "nsc_train_backward": (_i32, [_vp, _vp, _vp, _vp, _i64, _vp, _vp]),
    "nsc_sgd_momentum_update": (_i32, [_vp, _vp, _vp, _i64, _f32, _f32, _f32, _vp]),
    "nsc_trainer_debug_tensor": (_i32, [_vp, C.c_char_p, C.POINTER(_vp)]),
    "nsc_trainer_destroy": (_i32, [_vp]),
    "nsc_tsv_open": (_i32, [C.POINTER(_vp), C.c_char_p]),
    "nsc_tsv_count": (_i64, [_vp]),
    "nsc_tsv_decode": (_i32, [_vp, _i64, _i64, _i32, _vp, _vp, _vp, _vp, _i32]),
    "nsc_tsv_stage": (_i32, [_vp, _i64, _i64, _i32, _vp, _i64, _vp, _vp, _vp, _vp, _i32, _vp, _i64, C.POINTER(_i64), _vp, _vp]),
    "nsc_tsv_range_bytes": (_i64, [_vp, _i64, _i64]),
    "nsc_tsv_decode_rows": (_i32, [_vp, _vp, _i64, _vp, _i32]),
    "nsc_tsv_tag": (_i32, [_vp, _i64, C.POINTER(C.c_char_p), C.POINTER(C.c_int32)]),
    "nsc_tsv_tags": (_i32, [_vp, _i64, _i64, _vp, _i64, C.POINTER(_i64)]),
    "nsc_tsv_allele_lens": (_i32, [_vp, _i64, _i64, _vp]),
    "nsc_tsv_tags_lens": (_i32, [_vp, _i64, _i64, _vp, _i64, C.POINTER(_i64), _vp]),
    "nsc_debug_inflate": (_i32, [_vp, _i64, _vp, _i64, _i32]),
    "nsc_debug_inflate_b64": (_i32, [_i32, _vp, _i64, _vp, _i64, _vp, _vp]),
    "nsc_tsv_close": (_i32, [_vp]),
}

_lib = None


class NscError(RuntimeError):
    pass


def load():
    """dlopen the in-tree library and bind every declared symbol (raises if absent)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NscError(
            "libneusomatic_cuda.so is not built (%s). Run `python -c 'import __graft_entry__ as g; "
            "g.build()'` or `python neusomatic_b200/build.py`; there is no fallback path." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)     # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if lib.nsc_abi_version() != 1:
        raise NscError("libneusomatic_cuda ABI mismatch")
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise NscError("libneusomatic_cuda: %s (code %d)" % (load().nsc_last_error().decode(), rc))


def ptr(t):
    """data pointer of a torch tensor / numpy array / None."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return t.data_ptr()
    return t.ctypes.data
