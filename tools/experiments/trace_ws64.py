"""Timeline of one CTA of fa_fwd_ws64_kernel (or fa_fwd_ws_kernel with argv[2] = ws) from the FA_TRACE build.
    python tools/trace_ws64.py [N] [ws64|ws]"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "flash-attention-v2-rdna3-minimal_b200")
os.environ.setdefault("FA_FWD_SM100_LIB", os.path.join(PKG, "lib", "libfa_fwd_sm100_trace.so"))
sys.path.insert(0, PKG)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from rocwmma_fattn import _capi  # noqa: E402
from rocwmma_fattn.FlashAttn import FlashAttentionFunction  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
kern = sys.argv[2] if len(sys.argv) > 2 else "ws64"
_capi.set_kernel(_capi.FA_KERNEL_WS64 if kern == "ws64" else _capi.FA_KERNEL_WS)
torch.manual_seed(0)
q, k, v = (torch.rand(1, 16, N, 128, dtype=torch.float16, device="cuda") for _ in range(3))
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
lo, hi = 40, 100  # steady state (the trace buffer keeps the last 128 iterations)


def d(a, b):
    return float(np.mean(a[lo:hi] - b[lo:hi]))


s0, s1, m = t[0], t[1], t[2]
stats = {
    "kernel": kern, "N": N,
    "period_mma_iter": float(np.mean(np.diff(m[lo:hi, 0]))),
    "sm0_period": float(np.mean(np.diff(s0[lo:hi, 0]))),
    "sm0_wait_s": d(s0[:, 1], s0[:, 0]), "sm0_ld": d(s0[:, 2], s0[:, 1]), "sm0_step": d(s0[:, 6], s0[:, 2]),
    "sm1_wait_s": d(s1[:, 1], s1[:, 0]), "sm1_ld": d(s1[:, 2], s1[:, 1]), "sm1_step": d(s1[:, 6], s1[:, 2]),
    "mma_wait_v": d(m[:, 1], m[:, 0]), "mma_wait_p0_early": d(m[:, 2], m[:, 1]), "mma_pv0_rest": d(m[:, 3], m[:, 2]),
    "mma_s0": d(m[:, 4], m[:, 3]), "mma_wait_p1_early": d(m[:, 5], m[:, 4]), "mma_pv1_rest": d(m[:, 6], m[:, 5]),
    "mma_s1": d(m[:, 7], m[:, 6]),
}
print(json.dumps(stats, indent=1))
for j in range(60, 64):
    print("j", j, "sm0", [int(x - m[60, 0]) for x in s0[j, [0, 1, 2, 6]]], "sm1", [int(x - m[60, 0]) for x in s1[j, [0, 1, 2, 6]]],
          "mma", [int(x - m[60, 0]) for x in m[j]])
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", f"trace_{kern}_n{N}.json"), "w") as fh:
    json.dump(stats, fh, indent=1)
