"""Tiny driver for ncu: a few launches of one Stable-Diffusion attention shape through the ComfyUI-shaped hook (no timing
claims).  python tools/prof_sd_shape.py B N Nkv heads dim_head [iters]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flash-attention-v2-rdna3-minimal_b200"))
import torch
from rocwmma_fattn import hooks

B, N, Nkv, heads, d = (int(x) for x in sys.argv[1:6])
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 4
torch.manual_seed(0)
q = torch.randn(B, N, heads * d, dtype=torch.float16, device="cuda")
k, v = (torch.randn(B, Nkv, heads * d, dtype=torch.float16, device="cuda") for _ in range(2))
for _ in range(iters):
    o = hooks.comfy_attention(q, k, v, heads)
torch.cuda.synchronize()
print("done", o.float().mean().item())
