import os, sys, ctypes
sys.path.insert(0, "/root/repo")
import torch
from walkgpt_b200 import _lib
from walkgpt_b200.modules import postprocess_masks_fused
lib = _lib.lib()
for n in (192, 768):
  for ws, wm in ((True, True), (False, True), (False, False)):
    low = torch.randn(n, 64, 64, device="cuda")
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
            print(f"{n} masks score={ws} mask={wm}: {name} {ms*1e3:.1f} us", flush=True)
