"""Experiment: one 64-image step as ONE call vs as two 32-image half-batches on two CUDA streams (the small MSQP / decoder / tail
kernels of one half can fill the gaps of the other half's ViT)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from walkgpt_b200.modules import GroundingPath
dev = torch.device("cuda:0")
B, S, H = 64, int(os.environ.get("S", "3")), 4096
model = GroundingPath(hidden_size=H, clip_layers=24, seed=0).to(dev)
g = torch.Generator().manual_seed(1)
px = [torch.randn(B, 3, 448, 448, generator=g).to(torch.bfloat16).to(dev) for _ in range(3)]
seg = [torch.randn(B * S, H, generator=g).to(torch.bfloat16).to(dev) for _ in range(3)]
offs = list(range(0, B * S + 1, S))
def one(i):
    return model(px[i % 3], seg[i % 3], offs)
def split(i, n, streams):
    cur = torch.cuda.current_stream()
    ev = torch.cuda.Event(); ev.record(cur)
    outs = []
    hb = B // n
    for k, st in enumerate(streams):
        st.wait_event(ev)
        with torch.cuda.stream(st):
            outs.append(model(px[i % 3][k * hb:(k + 1) * hb], seg[i % 3][k * hb * S:(k + 1) * hb * S], offs[:hb + 1]))
    for st in streams:
        cur.wait_stream(st)
    return outs
def timeit(fn, n=8):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n): fn(i)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
print(f"one call of 64: {timeit(one):.2f} ms/step", flush=True)
for n in (2, 4):
    streams = [torch.cuda.Stream() for _ in range(n)]
    print(f"{n} streams x {B // n}: {timeit(lambda i: split(i, n, streams)):.2f} ms/step", flush=True)
print(f"one call of 64: {timeit(one):.2f} ms/step", flush=True)
a = one(0); b2 = split(0, 2, [torch.cuda.Stream() for _ in range(2)]); torch.cuda.synchronize()
print("logits equal:", torch.equal(a["logits"], torch.cat([o["logits"] for o in b2])))
