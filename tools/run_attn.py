import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from walkgpt_b200 import ops
B, T, H = 64, 1025, 16
torch.manual_seed(0)
qkv = torch.randn(B, T, 3 * H * 64, device="cuda").bfloat16()
out = torch.empty(B, T, H * 64, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.attention_d64(qkv, H, 0.125, out=out)
torch.cuda.synchronize()
print("ok")
