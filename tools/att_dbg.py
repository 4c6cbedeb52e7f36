import os, sys
sys.path.insert(0, "/root/repo")
import torch
from walkgpt_b200 import ops
B, T, H = int(os.environ.get("B", 2)), int(os.environ.get("T", 1025)), int(os.environ.get("H", 16))
torch.manual_seed(0)
qkv = torch.randn(B, T, 3 * H * 64, device="cuda").bfloat16()
out = ops.attention_d64(qkv, H, 0.125)
torch.cuda.synchronize()
q, k, v = qkv.float().view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
ref = (torch.softmax((q @ k.transpose(-1, -2)) * 0.125, -1) @ v).transpose(1, 2).reshape(B, T, H * 64)
print("B", B, "T", T, "H", H, "max err", (out.float() - ref).abs().max().item(), flush=True)
