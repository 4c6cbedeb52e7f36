import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from walkgpt_b200 import ops
# the dominant launch of the step: ViT fc1 (LN2 output x W_fc1^T + bias, quick-GELU), B=64 -> M=65600
M, N, K = 65600, 4096, 1024
torch.manual_seed(0)
a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
b = torch.randn(N, device="cuda")
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.gemm(a, w, b, act=ops.ACT_QUICK_GELU, out=out)
torch.cuda.synchronize()
print("ok")
