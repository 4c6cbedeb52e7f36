"""Time wg_attention_d64 at the CLIP-tower shape (B=64, T=1025, 16 heads) and check it against an fp32 torch reference."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from walkgpt_b200 import ops
torch.manual_seed(0)
B, T, H = 2, int(os.environ.get("T", "1025")), 16
qkv = torch.randn(B, T, 3 * H * 64, device="cuda").bfloat16()
out = ops.attention_d64(qkv, H, 0.125)
q, k, v = qkv.float().view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
ref = (torch.softmax((q @ k.transpose(-1, -2)) * 0.125, -1) @ v).transpose(1, 2).reshape(B, T, H * 64)
err = (out.float() - ref).abs().max().item()
B = 64
qkv = torch.randn(B, T, 3 * H * 64, device="cuda").bfloat16()
out = torch.empty(B, T, H * 64, device="cuda", dtype=torch.bfloat16)
for _ in range(3): ops.attention_d64(qkv, H, 0.125, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for rep in range(3):
    e0.record()
    for _ in range(10): ops.attention_d64(qkv, H, 0.125, out=out)
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 10)
print(f"WG_ATTN_POLY={os.environ.get('WG_ATTN_POLY','-')}: {best:.4f} ms  {4*B*H*T*T*64/best/1e9:.1f} TFLOP/s  max_abs_err={err:.2e}", flush=True)
