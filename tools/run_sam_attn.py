"""One launch of wg_sam_attention per block kind at the ViT-H bench shape (16 images, 16 heads), for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from walkgpt_b200 import ops
B, H = 16, 16
torch.manual_seed(0)
for mode in (0, 1):
    rows = B * (25 * 196 if mode == 0 else 4096)
    qkv = torch.randn(rows, 3 * H * 80, device="cuda").bfloat16()
    rel = (torch.randn(64 if mode == 0 else 256, 80, device="cuda") * 0.1).bfloat16()
    for _ in range(2):
        out = ops.sam_attention(qkv, rel, B, mode, H)
    torch.cuda.synchronize()
print("ok")
