"""ctypes binding of libeo_b200.so (C ABI declared in include/eo_b200.h).

The library is the product: there is no Python/PyTorch fallback for any entry point.  If
the shared object is missing this module raises at import of the symbol table (`lib()`),
telling the user to build it (`python -m eo_diffusion_b200.build`)."""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
# EO_B200_LIB: load another build of the library (a path, e.g. an A/B variant produced by build(variant=...))
LIB_PATH = os.environ.get("EO_B200_LIB") or os.path.join(HERE, "libeo_b200.so")

EO_OK = 0
EO_MODE_FP32 = 0
EO_MODE_BF16 = 1
EO_DDPM_NCOEF = 12


class EoUnetCfg(C.Structure):
    """struct eo_unet_cfg (include/eo_b200.h) -- mirrors UNetModel.__init__ kwargs
    (reference backbones/unet_openai.py:553-575)."""
    _fields_ = [
        ("in_channels", C.c_int32), ("model_channels", C.c_int32), ("out_channels", C.c_int32),
        ("num_res_blocks", C.c_int32), ("n_attention_resolutions", C.c_int32),
        ("attention_resolutions", C.c_int32 * 8), ("n_channel_mult", C.c_int32),
        ("channel_mult", C.c_int32 * 8), ("time_emb_factor", C.c_int32),
        ("num_classes", C.c_int32), ("num_heads", C.c_int32), ("num_head_channels", C.c_int32),
        ("num_heads_upsample", C.c_int32), ("use_new_attention_order", C.c_int32),
        ("dims", C.c_int32), ("conv_resample", C.c_int32), ("use_scale_shift_norm", C.c_int32),
        ("resblock_updown", C.c_int32),
    ]


_P = C.c_void_p
_I = C.c_int
_F = C.c_float
_L = C.c_int64

# name -> (restype, argtypes); every symbol include/eo_b200.h declares
SIGNATURES = {
    "eo_last_error": (C.c_char_p, []),
    "eo_version": (_I, []),
    "eo_device_check": (_I, []),
    "eo_unet_create": (_I, [C.POINTER(EoUnetCfg), C.POINTER(_P)]),
    "eo_unet_destroy": (None, [_P]),
    "eo_unet_num_weights": (_I, [_P]),
    "eo_unet_weight_name": (C.c_char_p, [_P, _I]),
    "eo_unet_weight_shape": (_I, [_P, _I, C.POINTER(_L * 4)]),
    "eo_unet_set_weight": (_I, [_P, C.c_char_p, _P, C.POINTER(_L), _I]),
    "eo_unet_finalize": (_I, [_P, _I, _I, _I, _I, _P]),
    "eo_unet_forward": (_I, [_P, _P, _I, _P, _I, _P, _P, _P, _I, _P]),
    "eo_unet_build_time_tables": (_I, [_P, _I, _P]),
    "eo_unet_clear_time_tables": (_I, [_P]),
    "eo_unet_forward_timed": (_I, [_P, _P, _I, _P, _I, _P, _P, _P, _I, _P, _P]),
    "eo_unet_num_ops": (_I, [_P]),
    "eo_unet_op_info": (_I, [_P, _I, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p),
                             C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "eo_unet_op_executed_flops": (C.c_double, [_P, _I]),
    "eo_unet_device_bytes": (_L, [_P]),
    "eo_unet_launches_per_forward": (_I, [_P]),
    "eo_unet_read_activation": (_L, [_P, C.c_char_p, _P, _L, _I, _P]),
    "eo_ddpm_sum_mix": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "eo_ddpm_step": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "eo_ddpm_step_mix": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "eo_ddim_step": (_I, [_P, _P, _P, _P, _P, _F, _F, _F, _F, _F, _F, _L, _P]),
    "eo_cfg_combine": (_I, [_P, _P, _F, _P, _L, _P]),
    "eo_sample_ddim": (_I, [_P, _P, _P, _P, _I, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "eo_sample_ddpm": (_I, [_P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "eo_post_map": (_I, [_P, _P, _L, _I, _F, _P]),
    "eo_post_dim_masked": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "eo_post_grid_u8": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _F, _I, C.POINTER(_I), _P]),
    "eo_post_stats": (_I, [_P, _L, _P, _P, _P]),
    "eo_psnr": (_I, [_P, _P, _L, _F, _P, _P, _P]),
    "eo_ssim": (_I, [_P, _P, _I, _I, _I, _I, _F, C.POINTER(_F), _P, _P, _P, _P]),
    "eo_test_conv_tc": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "eo_test_attention_tc": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "eo_debug_conv_trace": (_I, [_P, _I]),
}

_lock = threading.Lock()
_lib = None


class EoError(RuntimeError):
    """A libeo_b200 call returned a negative status."""


def lib() -> C.CDLL:
    """Load libeo_b200.so once and attach the prototypes.  Raises if it was not built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise ImportError(
                    f"{LIB_PATH} not found: the CUDA extension has not been built "
                    "(run `python -m eo_diffusion_b200.build`).  eo_diffusion_b200 has no "
                    "CPU or PyTorch fallback.")
            l = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(l, name)   # AttributeError if the .so lacks a declared symbol
                fn.restype = res
                fn.argtypes = args
            _lib = l
    return _lib


def last_error() -> str:
    msg = lib().eo_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int, what: str = "") -> int:
    """Raise on a negative status.  EO_ERR_ARG for the class-conditional mismatch keeps the
    reference's AssertionError (unet_openai.py:758-760)."""
    if rc >= 0:
        return rc
    msg = last_error()
    if "must specify y if and only if" in msg:
        raise AssertionError(msg)
    raise EoError(f"{what + ': ' if what else ''}{msg} (status {rc})")


def require_cuda_tensor(t, name: str):
    if not t.is_cuda:
        raise RuntimeError(
            f"{name} is on {t.device}: eo_diffusion_b200 runs on CUDA (sm_100a) only and has "
            "no CPU fallback; move the model and its inputs to a B200 device")


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t) -> C.c_void_p:
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(None)
