"""Per-kernel timing of the MSQP stage (wg_msqp_forward) at the benchmark shapes: Path A (B=64, 32 x 32 tokens of width 1024) and
Path B (B=16, 64 x 64 tokens of width 256).  WG_MSQP_TC=0 runs the CUDA-core cross attention instead of the tcgen05 one."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from walkgpt_b200 import _lib
from walkgpt_b200 import modules as M
lib = _lib.lib()
for B, L, sam_dim in ((64, 1024, 1024), (16, 4096, 256)):
    torch.manual_seed(0)
    m = M.MultiScaleQFormerProjector(sam_dim, 4096, pad_to_square=True, target_square_side=6).cuda()
    x = torch.randn(B, L, sam_dim, device="cuda").bfloat16()
    for _ in range(3): y = m(x)
    torch.cuda.synchronize()
    lib.wg_profile_enable(1)
    for _ in range(10): m(x)
    torch.cuda.synchronize()
    buf = ctypes.create_string_buffer(1 << 16)
    lib.wg_profile_collect(buf, len(buf)); lib.wg_profile_enable(0)
    tot = 0.0
    for ln in buf.value.decode().splitlines():
        name, cnt, tms, fl, by = ln.split()
        tot += float(tms) / 10
        print(f"B={B} L={L}: {name:24s} {float(tms)/10:.4f} ms/forward ({int(cnt)//10} launches)", flush=True)
    print(f"B={B} L={L}: total {tot:.4f} ms/forward  WG_MSQP_TC={os.environ.get('WG_MSQP_TC','1')}  checksum {y.float().abs().mean().item():.6f}", flush=True)
