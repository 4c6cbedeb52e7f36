"""The decoder's image-side projection GEMM (M = 192 prompts x 1024 positions, N = 384, K = 768, fp32 out, per-position bias table):
time with / without the bias table, against N = 256 / 512 (pair kernel) and bf16 output."""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from walkgpt_b200 import ops
dev = "cuda"
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n)
    return best
M, K, hw = 192 * 1024, 768, 1024
a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
for N in (384, 256, 512):
    w = (torch.randn(N, K, device=dev) / math.sqrt(K)).bfloat16()
    tab = torch.randn(hw * N, device=dev); b = torch.randn(N, device=dev)
    y = torch.empty(M, N, device=dev); y16 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for name, fn in (("f32, bias table", lambda: ops.gemm(a, w, tab, out_mode=ops.OUT_F32, out=y, bias_period=hw)),
                     ("f32, bias vector", lambda: ops.gemm(a, w, b, out_mode=ops.OUT_F32, out=y)),
                     ("f32, no bias", lambda: ops.gemm(a, w, None, out_mode=ops.OUT_F32, out=y)),
                     ("bf16, bias table", lambda: ops.gemm(a, w, tab, out=y16, bias_period=hw))):
        ms = timeit(fn)
        print(f"N={N} {name:18s}: {ms:.4f} ms {2*M*N*K/ms/1e9:7.1f} TF/s  out {M*N*(2 if 'bf16' in name else 4)/ms/1e6:6.0f} GB/s", flush=True)
