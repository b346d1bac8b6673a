"""Timeline of the softmax step's inner phases (FA_TRACE build, clock64 stamps of one CTA of fa_fwd_ws_kernel).

    python flash-attention-v2-rdna3-minimal_b200/build.py --trace
    python tools/trace_ws_softmax.py [N] [lib-name]        # on the GPU box

Rows 0/1: softmax warp 0 of tile 0/1 [wait_s, s_ready, ld_done, -, -, -, step_end]; rows 5/6: the same warps inside
ws_softmax_step [first-half exps issued, pair barrier passed, early arrive, mid arrive, late arrive]; row 2: MMA thread.
"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "flash-attention-v2-rdna3-minimal_b200")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
libname = sys.argv[2] if len(sys.argv) > 2 else "trace"
os.environ["FA_FWD_SM100_LIB"] = os.path.join(PKG, "lib", f"libfa_fwd_sm100_{libname}.so")
sys.path.insert(0, PKG)
import numpy as np
import torch
from rocwmma_fattn import _capi
from rocwmma_fattn.FlashAttn import FlashAttentionFunction

_capi.set_kernel(_capi.FA_KERNEL_WS)
torch.manual_seed(0)
q, k, v = (torch.rand(1, 16, N, 128, dtype=torch.float16, device="cuda") for _ in range(3))
buf = torch.zeros(8 * 128 * 8, dtype=torch.int64, device="cuda")
for _ in range(2):
    FlashAttentionFunction.apply(q, k, v, None, False)
torch.cuda.synchronize()
_capi.lib.fa_trace_set.argtypes = [ctypes.c_void_p]
_capi.lib.fa_trace_set.restype = None
_capi.lib.fa_trace_set(buf.data_ptr())
FlashAttentionFunction.apply(q, k, v, None, False)
torch.cuda.synchronize()
_capi.lib.fa_trace_set(None)
t = buf.cpu().view(8, 128, 8).numpy().astype(np.int64)
nj = min(128, N // 128)
lo, hi = nj // 4, 3 * nj // 4
m = t[2]


def rel(a):  # mean over the steady state of (stamp - start of the same MMA iteration)
    return float(np.mean(a[lo:hi] - m[lo:hi, 0]))


out = {"N": N, "lib": libname, "period": float(np.mean(np.diff(m[lo:hi, 0])))}
out["mma"] = {n: round(rel(m[:, i]), 1) for i, n in enumerate(
    ["iter_start", "v_ready", "p0_early_seen", "pv0_issued", "s0_issued", "p1_early_seen", "pv1_issued", "iter_end"])}
for tile in (0, 1):
    s, f = t[tile], t[5 + tile]
    # softmax step j of tile t works on S_t(j), issued in MMA iteration j-1: show relative to iteration j's start
    d = {"wait_begin": rel(s[:, 0]), "s_ready": rel(s[:, 1]), "ld_done": rel(s[:, 2]), "first_half_issued": rel(f[:, 0]),
         "pair_bar_passed": rel(f[:, 1]), "early_arrive": rel(f[:, 2]), "mid_arrive": rel(f[:, 3]),
         "late_arrive": rel(f[:, 4]), "step_end": rel(s[:, 6])}
    out[f"softmax{tile}"] = {k_: round(v_, 1) for k_, v_ in d.items()}
    out[f"softmax{tile}_durations"] = {
        "wait_s": round(d["s_ready"] - d["wait_begin"], 1), "ld": round(d["ld_done"] - d["s_ready"], 1),
        "first_half": round(d["first_half_issued"] - d["ld_done"], 1),
        "pair_bar": round(d["pair_bar_passed"] - d["first_half_issued"], 1),
        "pack_store_early": round(d["early_arrive"] - d["pair_bar_passed"], 1),
        "to_mid": round(d["mid_arrive"] - d["early_arrive"], 1), "to_late": round(d["late_arrive"] - d["mid_arrive"], 1),
        "sum_tail": round(d["step_end"] - d["late_arrive"], 1)}
print(json.dumps(out, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", f"trace_softmax_{libname}_n{N}.json"), "w") as fh:
    json.dump(out, fh, indent=1)
