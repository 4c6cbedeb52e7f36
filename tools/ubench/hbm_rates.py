"""HBM ceilings on this B200 for the roofline of the write-only tail kernel: copy (read + write), pure write, pure read."""
import torch
n = 1 << 30
a = torch.empty(n, dtype=torch.uint8, device="cuda"); b = torch.empty_like(a)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: b.copy_(a)); print(f"copy  1 GiB: {ms:.3f} ms  {2*n/ms/1e6:.0f} GB/s (read + write)")
ms = t(lambda: a.fill_(1)); print(f"fill  1 GiB: {ms:.3f} ms  {n/ms/1e6:.0f} GB/s (write only)")
af = a.view(torch.float32)
ms = t(lambda: af.sum()); print(f"sum   1 GiB: {ms:.3f} ms  {n/ms/1e6:.0f} GB/s (read only)")
c = torch.empty(192 * 448 * 448 * 5, dtype=torch.uint8, device="cuda")
ms = t(lambda: c.fill_(1)); print(f"fill  {c.numel()/1e6:.0f} MB (the tail kernel's output size at 192 masks): {ms*1e3:.1f} us  {c.numel()/ms/1e6:.0f} GB/s")
