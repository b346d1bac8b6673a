"""Tiny driver for ncu: a few launches of the forward at one shape (no timing claims).
    python tools/prof_kernel.py N [causal|x] [f16|bf16] [iters] [ws|sk|auto] [head_dim]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flash-attention-v2-rdna3-minimal_b200"))
import torch
from rocwmma_fattn import _capi
from rocwmma_fattn.FlashAttn import FlashAttentionFunction

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
causal = len(sys.argv) > 2 and sys.argv[2] == "causal"
dt = torch.bfloat16 if (len(sys.argv) > 3 and sys.argv[3] == "bf16") else torch.float16
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
D = int(sys.argv[6]) if len(sys.argv) > 6 else 128
if len(sys.argv) > 5 and sys.argv[5] != "auto":
    _capi.set_kernel({"ws": _capi.FA_KERNEL_WS, "sk": _capi.FA_KERNEL_SK}[sys.argv[5]])
torch.manual_seed(0)
q, k, v = (torch.rand(1, 16, N, D, dtype=dt, device="cuda") for _ in range(3))
for _ in range(iters):
    o = FlashAttentionFunction.apply(q, k, v, None, causal)
torch.cuda.synchronize()
print("done", o.float().mean().item())
