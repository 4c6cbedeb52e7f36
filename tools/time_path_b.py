"""Path B from the SAM image embeddings on (MSQP over 4096 tokens, CTP, SAM mask decoder at 64 x 64, postprocess to 1024^2):
time per 64-image batch with 3 [SEG] per image and the per-kernel breakdown of the library's event profiler."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from walkgpt_b200 import _lib
from walkgpt_b200 import modules as M
B, S, H = 64, 3, 4096
m = M.GroundingPathB(hidden_size=H, seed=0).to("cuda")
emb = torch.randn(B, 256, 64, 64, device="cuda")
seg = torch.randn(B * S, H, device="cuda")
offs = list(range(0, B * S + 1, S))
for _ in range(2): out = m(emb, seg, offs)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): out = m(emb, seg, offs)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"Path B after the encoder, B={B} S={S}: {ms:.2f} ms per batch = {B / ms * 1e3:.0f} images/s; logits {tuple(out['logits'].shape)}")
lib = _lib.lib()
lib.wg_profile_enable(1)
out = m(emb, seg, offs); torch.cuda.synchronize()
buf = ctypes.create_string_buffer(1 << 16)
lib.wg_profile_collect(buf, len(buf)); lib.wg_profile_enable(0)
rows = [ln.split() for ln in buf.value.decode().splitlines()]
for name, cnt, tms, fl, by in sorted(rows, key=lambda r: -float(r[2]))[:10]:
    print(f"  {name:28s} launches {cnt:>3s}  {float(tms):7.3f} ms")
