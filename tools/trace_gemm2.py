"""Debug: pipeline timeline of gemm2_bf16_kernel (library built with -DGEMM2_TRACE)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, math
from walkgpt_b200 import ops, _lib
M, N, K = 65600, 3072, 1024
a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16(); w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16(); b = torch.randn(N, device="cuda")
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for _ in range(3): ops.gemm(a, w, b, out=out)
torch.cuda.synchronize()
lib = ctypes.CDLL(os.path.join(os.path.dirname(_lib.__file__), "libwalkgpt_b200.so"))
buf = (ctypes.c_longlong * (3 * 64 * 4))()
print("rc", lib.wg_debug_gemm2_trace(buf))
import numpy as np
t = np.array(buf[:]).reshape(3, 64, 4)
t0 = t[0, 0, 0]
print("kb | MMA: pre_full got_full issued | prod0: pre_empty got_empty issued | prod1: pre_empty got_empty issued")
for i in range(32):
    print(i, [int(x - t0) for x in t[0, i, :3]], [int(x - t0) for x in t[1, i, [0,1,3,2]]], [int(x - t0) for x in t[2, i, [0,1,3,2]]])
