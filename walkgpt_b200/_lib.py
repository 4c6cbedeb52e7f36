"""ctypes binding of the C ABI in ``include/walkgpt_b200.h`` (``libwalkgpt_b200.so``).

There is deliberately no fallback: if the shared library is missing, or the device is not an
sm_100 GPU, every op raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwalkgpt_b200.so")

WG_OK = 0
ACT_NONE, ACT_QUICK_GELU, ACT_GELU_ERF, ACT_RELU = 0, 1, 2, 3
OUT_BF16, OUT_F32, OUT_BF16_LN = 0, 1, 2

vp = C.c_void_p
fp = C.c_void_p  # const float* (device) -- passed as raw address


class GemmArgs(C.Structure):
    _fields_ = [
        ("A", vp), ("lda", C.c_int64), ("W", vp), ("ldw", C.c_int64),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("bias", fp), ("bias_period", C.c_int32), ("act", C.c_int32), ("out_mode", C.c_int32),
        ("out", vp), ("ldo", C.c_int64), ("resid", vp), ("ln_gamma", fp), ("ln_beta", fp),
        ("ln_eps", C.c_float), ("reserved", C.c_int32),
    ]


class ClipLayer(C.Structure):
    _fields_ = [(n, vp) for n in ("ln1_g", "ln1_b", "w_qkv", "b_qkv", "w_o", "b_o", "ln2_g", "ln2_b",
                                  "w_fc1", "b_fc1", "w_fc2", "b_fc2")]


class ClipWeights(C.Structure):
    _fields_ = [
        ("hidden", C.c_int32), ("heads", C.c_int32), ("mlp", C.c_int32), ("image", C.c_int32), ("patch", C.c_int32),
        ("kpad", C.c_int32), ("n_layers", C.c_int32), ("reserved", C.c_int32),
        ("patch_w", vp), ("cls_emb", fp), ("pos_emb", fp), ("pre_ln_g", fp), ("pre_ln_b", fp),
        ("layers", C.POINTER(ClipLayer)),
    ]


_PROTOTYPES = {
    "wg_version": (C.c_int, []),
    "wg_last_error": (C.c_char_p, []),
    "wg_device_check": (C.c_int, [C.c_int]),
    "wg_gemm": (C.c_int, [C.POINTER(GemmArgs), vp]),
    "wg_layernorm": (C.c_int, [vp, C.c_int, C.c_int64, fp, fp, C.c_float, vp, C.c_int64, C.c_int64, C.c_int, vp]),
    "wg_attention_d64": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_float, vp]),
    "wg_clip_workspace_bytes": (C.c_size_t, [C.POINTER(ClipWeights), C.c_int]),
    "wg_clip_forward": (C.c_int, [C.POINTER(ClipWeights), vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int,
                                  vp, C.c_size_t, vp]),
}

_lib: Optional[C.CDLL] = None


class WalkGPTB200Error(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise WalkGPTB200Error(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C walkgpt_b200/csrc`).  walkgpt_b200 has no CPU / PyTorch fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOTYPES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def exported_symbols():
    return sorted(_PROTOTYPES)


def check(rc: int, what: str = "") -> None:
    if rc != WG_OK:
        msg = lib().wg_last_error()
        raise WalkGPTB200Error(f"{what or 'walkgpt_b200'} failed (code {rc}): {msg.decode() if msg else ''}")


def require_device(index: int = 0) -> None:
    check(lib().wg_device_check(index), "wg_device_check")
