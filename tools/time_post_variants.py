"""Kernel-only timing of wg_postprocess_masks variants: logits only / + u8 mask / + score, on random low-res logits (about half of
the pixels positive: the worst case for the per-warp sigmoid skip) and on disc-shaped masks (10 % positive, like real masks)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from walkgpt_b200 import _lib
from walkgpt_b200.modules import postprocess_masks_fused
lib = _lib.lib()
yy, xx = torch.meshgrid(torch.arange(64.0), torch.arange(64.0), indexing="ij")
disc = (((yy - 30) ** 2 + (xx - 34) ** 2) < 11.5 ** 2).float().cuda() * 10 - 5   # 10 % of the map positive
for n, kind in ((192, "random"), (768, "random"), (192, "disc"), (768, "disc")):
  for ws, wm in ((True, True), (False, True), (False, False)):
    low = torch.randn(n, 64, 64, device="cuda") if kind == "random" else (disc[None] + 0.3 * torch.randn(n, 64, 64, device="cuda")).contiguous()
    for _ in range(3): postprocess_masks_fused(low, (448, 448), (448, 448), want_mask=wm, want_score=ws)
    torch.cuda.synchronize()
    lib.wg_profile_enable(1)
    for _ in range(10): postprocess_masks_fused(low, (448, 448), (448, 448), want_mask=wm, want_score=ws)
    torch.cuda.synchronize()
    buf = ctypes.create_string_buffer(1 << 16)
    lib.wg_profile_collect(buf, len(buf)); lib.wg_profile_enable(0)
    for ln in buf.value.decode().splitlines():
        name, cnt, tms, fl, by = ln.split()
        if name.startswith("postprocess"):
            ms = float(tms) / int(cnt)
            print(f"{n} masks ({kind}) score={ws} mask={wm}: {name} {ms*1e3:.1f} us  {float(by)/int(cnt)/ms/1e6:.0f} GB/s", flush=True)
