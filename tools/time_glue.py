"""Kernel-only timing (library event profiler) of the SURVEY 8(f) glue kernels at the benchmark sizes."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from walkgpt_b200 import _lib, ops
lib = _lib.lib()
dev = "cuda"
vis = torch.randn(64, 36, 4096, device=dev).bfloat16()
out = torch.randint(0, 2, (192, 448, 448), device=dev, dtype=torch.uint8); tgt = torch.randint(0, 2, (192, 448, 448), device=dev, dtype=torch.uint8)
rows, Lin, H = 64, 512, 4096
ids = torch.randint(0, 32000, (rows, Lin), device=dev); ids[:, [100, 200, 300]] = 32003
hid = torch.randn(rows, Lin + 255, H, device=dev).bfloat16()
def run():
    ops.resample_tokens(vis, 16)
    ops.intersection_and_union(out, tgt, 2, 255)
    ops.seg_gather(hid, ids, 32003, list(range(rows + 1)), 255, max_out=rows * 3)
for _ in range(3): run()
torch.cuda.synchronize()
lib.wg_profile_enable(1)
for _ in range(10): run()
torch.cuda.synchronize()
buf = ctypes.create_string_buffer(1 << 16)
lib.wg_profile_collect(buf, len(buf)); lib.wg_profile_enable(0)
for ln in buf.value.decode().splitlines():
    name, cnt, tms, fl, by = ln.split()
    ms = float(tms) / 10
    print(f"{name:24s} {ms*1e3:8.1f} us per call  {float(by)/10/ms/1e6:8.0f} GB/s (algorithmic bytes)")
