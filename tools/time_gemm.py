"""Time wg_gemm at the CLIP-tower / projector shapes for each epilogue (bf16 out with activations, fp32 out with residual)."""
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
names = {0: "none", 1: "quick_gelu", 2: "gelu_erf", 3: "relu"}
for (M, N, K, acts) in [(65600, 3072, 1024, [0]), (65600, 4096, 1024, [0, 1, 3]), (65536, 8192, 1024, [0, 2]), (65600, 1024, 4096, [0]), (65536, 4096, 8192, [0])]:
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16(); w = (torch.randn(N, K, device=dev) / math.sqrt(K)).bfloat16(); b = torch.randn(N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for act in acts:
        ms = timeit(lambda: ops.gemm(a, w, b, act=act, out=out))
        print(f"bf16 out  M{M} N{N} K{K} {names[act]:>10}: {ms:.3f} ms {2*M*N*K/ms/1e9:7.1f} TF/s", flush=True)
    if N <= 4096:
        x = torch.randn(M, N, device=dev)
        ms = timeit(lambda: ops.gemm(a, w, b, out_mode=ops.OUT_F32, out=x, resid=x))
        print(f"f32+resid M{M} N{N} K{K}           : {ms:.3f} ms {2*M*N*K/ms/1e9:7.1f} TF/s", flush=True)
    ms = timeit(lambda: torch.matmul(a, w.T))
    print(f"cublas    M{M} N{N} K{K}           : {ms:.3f} ms {2*M*N*K/ms/1e9:7.1f} TF/s", flush=True)
