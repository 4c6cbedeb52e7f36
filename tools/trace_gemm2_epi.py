"""Debug: MMA / epilogue hand-off timeline of gemm2_bf16_kernel at the out-proj shape (library built with -DGEMM2_TRACE)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, math
import numpy as np
from walkgpt_b200 import ops, _lib
M, N, K = 65600, 1024, int(os.environ.get("K", "1024"))
a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16(); w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16(); b = torch.randn(N, device="cuda")
x = torch.randn(M, N, device="cuda")
mode = os.environ.get("MODE", "resid")
for _ in range(3):
    if mode == "resid": ops.gemm(a, w, b, out_mode=ops.OUT_F32, out=x, resid=x)
    elif mode == "f32": ops.gemm(a, w, b, out_mode=ops.OUT_F32, out=x)
    else: ops.gemm(a, w, b)
torch.cuda.synchronize()
lib = ctypes.CDLL(os.path.join(os.path.dirname(_lib.__file__), "libwalkgpt_b200.so"))
buf = (ctypes.c_longlong * (16 * 8))()
print("rc", lib.wg_debug_gemm2_epi(buf), "mode", mode, "K", K)
t = np.array(buf[:]).reshape(16, 8); t0 = t[0, 0]
print("tile | MMA: wait_empty got_empty committed | EPI: wait_full got_full drained   (cycles from the first stamp)")
for i in range(13):
    print(i, [int(v - t0) for v in t[i, :3]], [int(v - t0) for v in t[i, 3:6]], " epi busy", int(t[i, 5] - t[i, 4]), " mma tile", int(t[i, 2] - t[i, 1]))
