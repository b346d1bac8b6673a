"""Timeline of one CTA of fa_fwd_ws_kernel from the FA_TRACE debug build (clock64 stamps).

    python flash-attention-v2-rdna3-minimal_b200/build.py --trace      # here (no GPU needed)
    FA_FWD_SM100_LIB=.../lib/libfa_fwd_sm100_trace.so python tools/trace_ws.py 16384   # on the GPU box
"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "flash-attention-v2-rdna3-minimal_b200")
os.environ.setdefault("FA_FWD_SM100_LIB", os.path.join(PKG, "lib", "libfa_fwd_sm100_trace.so"))
sys.path.insert(0, PKG)
import torch
from rocwmma_fattn import _capi
from rocwmma_fattn.FlashAttn import FlashAttentionFunction

_capi.set_kernel(_capi.FA_KERNEL_WS)  # the timeline stamps live in the one-shot kernel
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
causal = len(sys.argv) > 2 and sys.argv[2] == "causal"
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 2  # launches before the traced one (clock state)
D = int(sys.argv[4]) if len(sys.argv) > 4 else 128
torch.manual_seed(0)
q, k, v = (torch.rand(1, 16, N, D, dtype=torch.float16, device="cuda") for _ in range(3))
buf = torch.zeros(5 * 128 * 8, dtype=torch.int64, device="cuda")
for _ in range(warm):
    FlashAttentionFunction.apply(q, k, v, None, causal)
torch.cuda.synchronize()
_capi.lib.fa_trace_set.argtypes = [ctypes.c_void_p]
_capi.lib.fa_trace_set.restype = None
_capi.lib.fa_trace_set(buf.data_ptr())
FlashAttentionFunction.apply(q, k, v, None, causal)
torch.cuda.synchronize()
_capi.lib.fa_trace_set(None)
t = buf.cpu().view(5, 128, 8).numpy()
nj = min(128, N // 128)
base = t[2, 0, 0]
out = {"N": N, "causal": causal, "roles": {}}
names = {0: "softmax0 [wait_s, s_ready, ld_done, max_done, p_early, p_late, sum_done]",
         1: "softmax1 [same]",
         2: "mma [iter_start, v_ready, po0_ready, pv0_issued, s0_issued, po1_ready, pv1_issued, iter_end]",
         3: "softmax0 partner warp 4 [same as softmax0]", 4: "tma [k_slot_free, v_slot_free]"}
for role in range(5):
    print("#", names[role])
    for j in list(range(0, min(nj, 6))) + list(range(max(6, nj // 2), min(nj, nj // 2 + 6))):
        row = [int(x - base) if x else None for x in t[role, j]]
        print(role, j, row)
# steady-state statistics over the middle of the loop
import numpy as np
lo, hi = nj // 4, 3 * nj // 4
def d(a, b):
    return float(np.mean(a[lo:hi].astype(np.int64) - b[lo:hi].astype(np.int64)))
s0, s1, m, s0b = t[0], t[1], t[2], t[3]
stats = {
    "period_mma": float(np.mean(np.diff(m[lo:hi, 0].astype(np.int64)))),
    "sm0_wait_s": d(s0[:, 1], s0[:, 0]), "sm0_ld": d(s0[:, 2], s0[:, 1]), "sm0_max": d(s0[:, 3], s0[:, 2]),
    "sm0_exp_early": d(s0[:, 4], s0[:, 3]), "sm0_exp_late": d(s0[:, 5], s0[:, 4]), "sm0_sum": d(s0[:, 6], s0[:, 5]),
    "sm1_wait_s": d(s1[:, 1], s1[:, 0]), "sm1_ld": d(s1[:, 2], s1[:, 1]), "sm1_max": d(s1[:, 3], s1[:, 2]),
    "sm1_exp_early": d(s1[:, 4], s1[:, 3]), "sm1_exp_late": d(s1[:, 5], s1[:, 4]), "sm1_sum": d(s1[:, 6], s1[:, 5]),
    "sm0b_wait_s": d(s0b[:, 1], s0b[:, 0]), "sm0b_ld": d(s0b[:, 2], s0b[:, 1]), "sm0b_max": d(s0b[:, 3], s0b[:, 2]),
    "sm0b_exp_early": d(s0b[:, 4], s0b[:, 3]), "sm0b_exp_late": d(s0b[:, 5], s0b[:, 4]), "sm0b_sum": d(s0b[:, 6], s0b[:, 5]),
    "sm0_vs_sm0b_s_ready": d(s0b[:, 1], s0[:, 1]), "sm0_vs_sm0b_max_done": d(s0b[:, 3], s0[:, 3]),
    "mma_commit_s0_to_sm0_ready": float(np.mean(s0[lo + 1:hi + 1, 1].astype(np.int64) - m[lo:hi, 4].astype(np.int64))),
    "mma_wait_v": d(m[:, 1], m[:, 0]), "mma_wait_po0": d(m[:, 2], m[:, 1]), "mma_issue_pv0": d(m[:, 3], m[:, 2]),
    "mma_issue_s0": d(m[:, 4], m[:, 3]), "mma_wait_po1": d(m[:, 5], m[:, 4]), "mma_issue_pv1": d(m[:, 6], m[:, 5]),
    "mma_issue_s1": d(m[:, 7], m[:, 6]),
    "p0_early_to_mma_seen": d(m[:, 2], s0[:, 4]), "p1_early_to_mma_seen": d(m[:, 5], s1[:, 4]),
    # S_t(j+1) issued at m[j,4] / m[j,7]; softmax sees it at s[j+1,1]
    "s0_issue_to_ready": float(np.mean(s0[lo + 1:hi + 1, 1].astype(np.int64) - m[lo:hi, 4].astype(np.int64))),
    "s1_issue_to_ready": float(np.mean(s1[lo + 1:hi + 1, 1].astype(np.int64) - m[lo:hi, 7].astype(np.int64))),
}
c0, g0, c1, g1 = (int(x) for x in t[4, 0, 4:8])
stats["cta_cycles"] = c1 - c0
stats["cta_ns"] = g1 - g0
stats["sm_mhz_in_kernel"] = round((c1 - c0) / max(1, g1 - g0) * 1e3, 1)
print(json.dumps(stats, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", f"trace_ws_n{N}_d{D}{'_causal' if causal else ''}.json"), "w") as fh:
    json.dump({"stats": stats, "raw": t.tolist()}, fh)
