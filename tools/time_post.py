"""Kernel-only timing of wg_postprocess_masks (bilinear 64^2 -> 448^2 + threshold + score) with the library's event profiler."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from walkgpt_b200 import _lib
from walkgpt_b200.modules import postprocess_masks_fused
lib = _lib.lib()
for n in (192, 768):
    low = torch.randn(n, 64, 64, device="cuda")
    for _ in range(3): postprocess_masks_fused(low, (448, 448), (448, 448))
    torch.cuda.synchronize()
    lib.wg_profile_enable(1)
    for _ in range(10): postprocess_masks_fused(low, (448, 448), (448, 448))
    torch.cuda.synchronize()
    buf = ctypes.create_string_buffer(1 << 16)
    lib.wg_profile_collect(buf, len(buf))
    lib.wg_profile_enable(0)
    for ln in buf.value.decode().splitlines():
        name, cnt, tms, fl, by = ln.split()
        if name.startswith("postprocess"):
            ms = float(tms) / int(cnt)
            print(f"{n} masks: {name} {ms*1e3:.1f} us per launch, {float(by)/int(cnt)/ms/1e6:.0f} GB/s (algorithmic bytes)")
