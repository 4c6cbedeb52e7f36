"""A few launches of wg_postprocess_masks at the bench shape (192 masks of 64^2 -> 448^2, random logits), for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from walkgpt_b200.modules import postprocess_masks_fused
n = int(sys.argv[1]) if len(sys.argv) > 1 else 192
low = torch.randn(n, 64, 64, device="cuda")
for _ in range(4):
    postprocess_masks_fused(low, (448, 448), (448, 448))
torch.cuda.synchronize()
print("ok")
