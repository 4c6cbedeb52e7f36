"""Debug: pipeline timeline of attention_d64_kernel (library built with -DATT_TRACE).  Prints clock64 deltas per key block."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from walkgpt_b200 import ops, _lib
B, T, H = 64, 1025, 16
torch.manual_seed(0)
qkv = torch.randn(B, T, 3 * H * 64, device="cuda").bfloat16()
out = torch.empty(B, T, H * 64, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.attention_d64(qkv, H, 0.125, out=out)
torch.cuda.synchronize()
lib = ctypes.CDLL(os.path.join(os.path.dirname(_lib.__file__), "libwalkgpt_b200.so"))
buf = (ctypes.c_longlong * (3 * 3 * 16 * 8))()
rc = lib.wg_debug_attn_trace(buf)
print("rc", rc)
import numpy as np
a = np.array(buf[:]).reshape(3, 3, 16, 8)
for c in range(3):
    t0 = a[c, 0, 0, 0]
    print(f"--- CTA {c}: softmax thread (wait_s_full, got_s_full, ld_done, pre_pfree, got_pfree, p_ready_arrived) | MMA (pre_sfree, got_sfree, got_kfull, pre_pready, got_pready, got_vfull, pv_issued)")
    for j in range(9):
        sm = [int(x - t0) if x else -1 for x in a[c, 0, j, :6]]
        mm = [int(x - t0) if x else -1 for x in a[c, 1, j, :7]]
        print(j, sm, "|", mm)
