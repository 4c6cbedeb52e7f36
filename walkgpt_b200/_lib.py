"""ctypes binding of the C ABI in ``include/walkgpt_b200.h`` (``libwalkgpt_b200.so``).

The struct layouts and function prototypes are PARSED FROM THE HEADER at import time, so the Python side cannot
drift from the ABI.  There is deliberately no fallback: if the shared library is missing, or the device is not an
sm_100 GPU, every op raises.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from typing import Dict, List, Optional, Tuple

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwalkgpt_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "walkgpt_b200.h")

WG_OK = 0

_SCALARS = {"int32_t": C.c_int32, "int": C.c_int, "int64_t": C.c_int64, "size_t": C.c_size_t, "float": C.c_float,
            "uint8_t": C.c_uint8, "long long": C.c_longlong}

STRUCTS: Dict[str, type] = {}
CONSTANTS: Dict[str, int] = {}
_PROTOTYPES: Dict[str, Tuple[object, List[object]]] = {}


def _strip_comments(src: str) -> str:
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    return re.sub(r"//[^\n]*", " ", src)


def _parse_header(path: str) -> None:
    src = _strip_comments(open(path).read())
    for m in re.finditer(r"#define\s+(WG_\w+)\s+\(?(-?\d+)\)?\s*$", src, flags=re.M):
        CONSTANTS[m.group(1)] = int(m.group(2))
    for m in re.finditer(r"typedef\s+struct\s+(\w+)\s*\{(.*?)\}\s*(\w+)\s*;", src, flags=re.S):
        name, body = m.group(3), m.group(2)
        fields: List[Tuple[str, object]] = []
        for decl in body.split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            if "*" in decl:  # pointer field(s): "const T* a" (one per declaration in this header)
                fname = decl.split("*")[-1].strip()
                fields.append((fname, C.c_void_p))
                continue
            am = re.match(r"(\w+)\s+(\w+)\[(\d+)\]$", decl)
            if am:  # nested struct array
                fields.append((am.group(2), STRUCTS[am.group(1)] * int(am.group(3))))
                continue
            parts = decl.split(" ", 1)
            ctype = STRUCTS[parts[0]] if parts[0] in STRUCTS else _SCALARS[parts[0]]  # nested struct by value, or a scalar
            for fname in parts[1].split(","):
                fields.append((fname.strip(), ctype))
        STRUCTS[name] = type(name, (C.Structure,), {"_fields_": fields})
    for m in re.finditer(r"WG_API\s+([\w\s\*]+?)\s*(wg_\w+)\s*\((.*?)\)\s*;", src, flags=re.S):
        ret_s, fname, args_s = " ".join(m.group(1).split()), m.group(2), " ".join(m.group(3).split())
        if "*" in ret_s:
            ret = C.c_char_p
        else:
            ret = _SCALARS[ret_s]
        args: List[object] = []
        if args_s and args_s != "void":
            for a in args_s.split(","):
                a = a.strip()
                if "*" in a:
                    args.append(C.c_void_p)
                else:
                    a = a.replace("const ", "")
                    args.append(_SCALARS["long long" if a.startswith("long long") else a.split(" ")[0]])
        _PROTOTYPES[fname] = (ret, args)


_parse_header(HEADER_PATH)

ACT_NONE, ACT_QUICK_GELU = CONSTANTS["WG_ACT_NONE"], CONSTANTS["WG_ACT_QUICK_GELU"]
ACT_GELU_ERF, ACT_RELU = CONSTANTS["WG_ACT_GELU_ERF"], CONSTANTS["WG_ACT_RELU"]
OUT_BF16, OUT_F32, OUT_BF16_LN = CONSTANTS["WG_OUT_BF16"], CONSTANTS["WG_OUT_F32"], CONSTANTS["WG_OUT_BF16_LN"]

GemmArgs = STRUCTS["wg_gemm_args"]
ClipLayer = STRUCTS["wg_clip_layer"]
ClipWeights = STRUCTS["wg_clip_weights"]
MsqpBlock = STRUCTS["wg_msqp_block"]
MsqpScale = STRUCTS["wg_msqp_scale"]
MsqpWeights = STRUCTS["wg_msqp_weights"]
CtpWeights = STRUCTS["wg_ctp_weights"]
ProjNeckWeights = STRUCTS["wg_proj_neck_weights"]
TwoWayLayer = STRUCTS["wg_twoway_layer"]
MaskDecoderWeights = STRUCTS["wg_mask_decoder_weights"]
SamBlock = STRUCTS["wg_sam_block"]
SamEncoderWeights = STRUCTS["wg_sam_encoder_weights"]

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
            fn = getattr(handle, name)  # AttributeError here == header declares a symbol the library lacks
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def exported_symbols() -> List[str]:
    """Every function the header declares."""
    return sorted(_PROTOTYPES)


def check(rc: int, what: str = "") -> None:
    if rc != WG_OK:
        msg = lib().wg_last_error()
        raise WalkGPTB200Error(f"{what or 'walkgpt_b200'} failed (code {rc}): {msg.decode() if msg else ''}")


def require_device(index: int = 0) -> None:
    check(lib().wg_device_check(index), "wg_device_check")
