"""Per-CTA timeline of fa_fwd_sk_kernel from the FA_TRACE debug build (globaltimer stamps of tile 0's leader).

    python flash-attention-v2-rdna3-minimal_b200/build.py --trace      # here (no GPU needed)
    python tools/trace_sk.py 2048 [heads]                              # on the GPU box
"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "flash-attention-v2-rdna3-minimal_b200")
os.environ.setdefault("FA_FWD_SM100_LIB", os.path.join(PKG, "lib", "libfa_fwd_sm100_trace.so"))
sys.path.insert(0, PKG)
import torch  # noqa: E402
from rocwmma_fattn import _capi  # noqa: E402
from rocwmma_fattn.FlashAttn import FlashAttentionFunction  # noqa: E402

_capi.set_kernel(_capi.FA_KERNEL_SK)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
H = int(sys.argv[2]) if len(sys.argv) > 2 else 16
torch.manual_seed(0)
q, k, v = (torch.rand(1, H, N, 128, dtype=torch.float16, device="cuda") for _ in range(3))
buf = torch.zeros(5 * 128 * 8, dtype=torch.int64, device="cuda")
for _ in range(3):
    FlashAttentionFunction.apply(q, k, v, None, False)
torch.cuda.synchronize()
_capi.lib.fa_trace_set.argtypes = [ctypes.c_void_p]
_capi.lib.fa_trace_set.restype = None
_capi.lib.fa_trace_set(buf.data_ptr())
FlashAttentionFunction.apply(q, k, v, None, False)
torch.cuda.synchronize()
_capi.lib.fa_trace_set(None)
t = buf.cpu().numpy()[:148 * 32].reshape(148, 32)
t0 = int(t[:, 0][t[:, 0] > 0].min())
rows = []
for cta in range(148):
    if t[cta, 0] == 0:
        continue
    rows.append([cta] + [round((int(x) - t0) / 1e3, 2) if x else None for x in t[cta, :13]])
print("cta, start, then per segment (loop_done, partials_awaited, epilogue_done) in us")
for r in rows[:12] + rows[70:76] + rows[-6:]:
    print(r)
ends = [max(int(x) for x in t[c] if x) - t0 for c in range(148) if t[c, 0]]
print("last stamp per CTA (us): min %.2f max %.2f mean %.2f" % (min(ends) / 1e3, max(ends) / 1e3, sum(ends) / len(ends) / 1e3))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", f"trace_sk_n{N}_h{H}.json"), "w") as fh:
    json.dump(rows, fh)
