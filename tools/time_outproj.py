"""Where the out-proj GEMM (M=65600, N=1024, K=1024, fp32 out + fp32 residual in place) spends its time: with / without the residual
read, bf16 output for comparison, and the same bytes as plain copies."""
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
M, N = 65600, 1024
for K in (1024, 4096):
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16(); w = (torch.randn(N, K, device=dev) / math.sqrt(K)).bfloat16(); b = torch.randn(N, device=dev)
    x = torch.randn(M, N, device=dev); y = torch.empty(M, N, device=dev); o16 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    fl = 2 * M * N * K
    for name, fn, by in (("f32 + resid in place", lambda: ops.gemm(a, w, b, out_mode=ops.OUT_F32, out=x, resid=x), M * K * 2 + M * N * 8),
                         ("f32 + resid, separate out", lambda: ops.gemm(a, w, b, out_mode=ops.OUT_F32, out=y, resid=x), M * K * 2 + M * N * 8),
                         ("f32, no resid", lambda: ops.gemm(a, w, b, out_mode=ops.OUT_F32, out=y), M * K * 2 + M * N * 4),
                         ("bf16 out", lambda: ops.gemm(a, w, b, out=o16), M * K * 2 + M * N * 2)):
        ms = timeit(fn)
        print(f"K={K} {name:28s}: {ms:.4f} ms {fl/ms/1e9:7.1f} TF/s  {by/ms/1e6:6.0f} GB/s", flush=True)
ms = timeit(lambda: y.copy_(x)); print(f"copy fp32 [M,N]               : {ms:.4f} ms {M*N*8/ms/1e6:6.0f} GB/s")
ms = timeit(lambda: x.add_(y)); print(f"x += y  fp32 [M,N]            : {ms:.4f} ms {M*N*12/ms/1e6:6.0f} GB/s")
