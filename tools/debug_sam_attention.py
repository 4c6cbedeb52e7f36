"""Bring-up aid for wg_sam_attention: one mode per process (run under `timeout`), max error against a plain torch restatement of
softmax(q k^T / sqrt(80) + rel_h + rel_w) v computed on the GPU.  WG_SAM_DBG=1 drops the 16-column (32-byte swizzle) MMAs, =3 also
their TMA loads; the comparison then uses only the first 64 channels of q / k and the first 64 output channels."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from walkgpt_b200 import modules as M, ops  # noqa: E402

mode = int(sys.argv[1])
dbg = int(os.environ.get("WG_SAM_DBG", "0"))
heads, side = 2, (14 if mode == 0 else 64)
units, L = (25 if mode == 0 else 1), side * side
torch.manual_seed(0)
q, k, v = (torch.randn(units, heads, L, 80, device="cuda").bfloat16().float() for _ in range(3))
rh, rw = (torch.randn(2 * side - 1, 80, device="cuda") * 0.11).bfloat16().float(), (torch.randn(2 * side - 1, 80, device="cuda") * 0.11).bfloat16().float()


def cols(t, lo, hi):
    return t[..., lo:hi].permute(0, 2, 1, 3).reshape(units * L, heads * (hi - lo))


qkv = torch.cat([cols(q, 0, 64), cols(k, 0, 64), cols(v, 0, 64), cols(q, 64, 80), cols(k, 64, 80), cols(v, 64, 80)], 1).bfloat16().contiguous()
rel = M.sam_rel_table(rh, rw, side).bfloat16()
marks = None
if dbg & 8:
    import collections
    import ctypes
    import time
    from walkgpt_b200 import _lib
    n_cta = (2 * heads * 25) if mode == 0 else (32 * heads)
    marks = torch.zeros(n_cta * 192, dtype=torch.int32).pin_memory()
    _lib.lib().wg_debug_sam_marks.argtypes = [ctypes.c_void_p]
    print("marks rc", _lib.lib().wg_debug_sam_marks(marks.data_ptr()), flush=True)
print('launching', flush=True)
out = ops.sam_attention(qkv, rel, 1, mode, heads)
print('launched', flush=True)
if marks is not None:
    time.sleep(5)
    m = marks.view(n_cta, 192)
    hist = collections.Counter()
    for c in range(n_cta):
        row = m[c].tolist()
        key = (row[0], row[32], tuple(sorted(set(row[64:]))))
        hist[key] += 1
    for k_, v_ in hist.most_common(12):
        print("producer", k_[0], "mma", k_[1], "softmax", k_[2], "x", v_, flush=True)
    for c in range(min(n_cta, 6)):
        row = m[c].tolist()
        print("cta", c, "producer", row[0], "mma", row[32], "softmax", sorted(set(row[64:])), flush=True)
    os._exit(0)
torch.cuda.synchronize()
print("kernel returned", flush=True)
d = 64 if dbg & 1 else 80
qq, kk = q[..., :d], k[..., :d]
s = (qq * 80 ** -0.5) @ kk.transpose(-1, -2)
idx = torch.arange(side, device="cuda")[:, None] - torch.arange(side, device="cuda")[None, :] + side - 1
bh = torch.einsum("uahwc,hkc->uahwk", qq.view(units, heads, side, side, d), rh[idx][..., :d])
bw = torch.einsum("uahwc,wkc->uahwk", qq.view(units, heads, side, side, d), rw[idx][..., :d])
s = (s.view(units, heads, side, side, side, side) + bh[..., :, None] + bw[..., None, :]).view(units, heads, L, L)
ref = s.softmax(-1) @ v                                   # [units, heads, L, 80]
if mode == 0:
    full = ref.view(5, 5, heads, 14, 14, 80).permute(0, 3, 1, 4, 2, 5).reshape(70, 70, heads, 80)[:64, :64]
else:
    full = ref[0].permute(1, 0, 2).reshape(64, 64, heads, 80)
got = out.float().view(64, 64, heads, 80)
e_main = (got[..., :64] - full[..., :64]).abs().max().item()
e_rem = (got[..., 64:] - full[..., 64:]).abs().max().item()
print(f"mode {mode} dbg {dbg}: max err main {e_main:.4f} rem {e_rem:.4f} (ref abs-max {full.abs().max().item():.3f})", flush=True)
