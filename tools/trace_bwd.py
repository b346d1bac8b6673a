"""Timeline of one CTA of fa_bwd_ws_kernel from the FA_TRACE debug build (clock64 stamps; tools like trace_ws.py).

    python flash-attention-v2-rdna3-minimal_b200/build.py --trace ; python tools/trace_bwd.py [N]      # on the GPU box
"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "flash-attention-v2-rdna3-minimal_b200")
os.environ["FA_FWD_SM100_LIB"] = os.path.join(PKG, "lib", "libfa_fwd_sm100_trace.so")
sys.path.insert(0, PKG)
import numpy as np
import torch
from rocwmma_fattn import _capi
from rocwmma_fattn.FlashAttn import flash_attn_wmma

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
H, D = 16, 128
torch.manual_seed(0)
q, k, v, d_o = (torch.rand(1, H, N, D, dtype=torch.float16, device="cuda") for _ in range(4))
o, qp, kp, vp, o_pad, L = flash_attn_wmma.forward(q, k, v, 64, 128, False, D ** -0.5, False)
buf = torch.zeros(4 * 128 * 8, dtype=torch.int64, device="cuda")
for _ in range(2):
    flash_attn_wmma.backward(qp, kp, vp, o_pad, d_o, L, N, N, D, 128, 128, False, D ** -0.5, False)
torch.cuda.synchronize()
_capi.lib.fa_trace_set.argtypes = [ctypes.c_void_p]
_capi.lib.fa_trace_set.restype = None
_capi.lib.fa_trace_set(buf.data_ptr())
flash_attn_wmma.backward(qp, kp, vp, o_pad, d_o, L, N, N, D, 128, 128, False, D ** -0.5, False)
torch.cuda.synchronize()
_capi.lib.fa_trace_set(None)
t = buf.cpu().view(4, 128, 8).numpy().astype(np.int64)
nj = min(128, N // 128)
lo, hi = nj // 4, 3 * nj // 4
m, c, d = t[0], t[1], t[2]


def rel(a):
    return round(float(np.mean(a[lo:hi] - m[lo:hi, 0])), 1)


out = {"N": N, "period": float(np.mean(np.diff(m[lo:hi, 0]))),
       "mma (rel. to iteration start)": {n: rel(m[:, i]) for i, n in enumerate(
           ["start", "p_ready seen", "dV issued", "S(i+1) issued", "ds_ready seen", "dK,dQ issued", "do_full+drained seen", "do_full seen"])},
       "P/dS warp 0": {n: rel(c[:, i]) for i, n in enumerate(["A: wait begin", "S ready", "P handed over", "dP ready", "dS handed over", "B: chunk 0 loaded", "B: chunk 0 stored", "B: chunk 1 loaded"])},
       "drain warp": {n: rel(d[:, i]) for i, n in enumerate(["wait begin", "dQ ready", "dQ in registers (warp 8)", "(warp 9)", "(warp 10)", "(warp 11)"])}}
out["cycles between do_full seen and drained seen (mean)"] = float(np.mean(m[lo:hi, 6] - m[lo:hi, 7]))
print(json.dumps(out, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"trace_bwd_n{N}.json"), "w"), indent=1)
