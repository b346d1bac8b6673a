"""Timeline of one leader CTA of fa_fwd_quad2_kernel from the FA_TRACE debug build (clock64 stamps).
    python flash-attention-v2-rdna3-minimal_b200/build.py --trace ; on the GPU box: python tools/trace_quad2.py [N] [D]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "flash-attention-v2-rdna3-minimal_b200")
os.environ.setdefault("FA_FWD_SM100_LIB", os.path.join(PKG, "lib", "libfa_fwd_sm100_trace.so"))
sys.path.insert(0, PKG)
import numpy as np
import torch
from rocwmma_fattn import _capi
from rocwmma_fattn.FlashAttn import FlashAttentionFunction

_capi.set_kernel(_capi.FA_KERNEL_QUAD2)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
D = int(sys.argv[2]) if len(sys.argv) > 2 else 128
q, k, v = (torch.rand(1, 16, N, D, dtype=torch.float16, device="cuda") for _ in range(3))
buf = torch.zeros(5 * 128 * 8, dtype=torch.int64, device="cuda")
for _ in range(2):
    FlashAttentionFunction.apply(q, k, v, None, False)
torch.cuda.synchronize()
_capi.lib.fa_trace_set.argtypes = [ctypes.c_void_p]
_capi.lib.fa_trace_set.restype = None
_capi.lib.fa_trace_set(buf.data_ptr())
FlashAttentionFunction.apply(q, k, v, None, False)
torch.cuda.synchronize()
_capi.lib.fa_trace_set(None)
t = buf.cpu().view(5, 128, 8).numpy().astype(np.int64)
nj = min(128, N // 128)
lo, hi = nj // 4, 3 * nj // 4
sm0, sm1, mma = t[0], t[1], t[2]
print("softmax warp 0 [step_start, s_ready, ld_done, step_end]; mma [pv_start, o_prev_done, p_early_seen, p_late_seen, pv_issued, s_issued]")
for j in list(range(0, 4)) + list(range(lo, lo + 4)):
    base = sm0[lo, 0]
    print(j, "sm0", [int(x - base) for x in sm0[j, :4]], "sm1", [int(x - base) for x in sm1[j, :4]], "mma", [int(x - base) for x in mma[j, :6]])
d = lambda a, b: float(np.mean(a[lo:hi] - b[lo:hi]))
print("period (softmax step start to next)", float(np.mean(np.diff(sm0[lo:hi, 0]))))
print("sm0: wait_s %.0f  ld %.0f  step %.0f" % (d(sm0[:, 1], sm0[:, 0]), d(sm0[:, 2], sm0[:, 1]), d(sm0[:, 3], sm0[:, 2])))
print("sm1: wait_s %.0f  ld %.0f  step %.0f" % (d(sm1[:, 1], sm1[:, 0]), d(sm1[:, 2], sm1[:, 1]), d(sm1[:, 3], sm1[:, 2])))
print("mma: wait_o_prev %.0f  wait_p_early %.0f  issue4+wait_p_late %.0f  issue4+commits %.0f  issue_s %.0f"
      % (d(mma[:, 1], mma[:, 0]), d(mma[:, 2], mma[:, 1]), d(mma[:, 3], mma[:, 2]), d(mma[:, 4], mma[:, 3]), d(mma[:, 5], mma[:, 4])))
print("softmax step end -> mma sees p_late %.0f" % d(mma[:, 3], sm0[:, 3]))
print("mma s_issued(j) [= S(j+2)] -> softmax s_ready(j+2) %.0f" % float(np.mean(sm0[lo + 2:hi + 2, 1] - mma[lo:hi, 5])))
