"""Tiny driver for ncu: a few backward launches at one head dim above 128 (no timing claims)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flash-attention-v2-rdna3-minimal_b200"))
import torch
from rocwmma_fattn.FlashAttn import flash_attn_wmma

D = int(sys.argv[1]) if len(sys.argv) > 1 else 160
N, H = 4096, 16
torch.manual_seed(0)
q, k, v, d_o = (torch.rand(1, H, N, D, dtype=torch.float16, device="cuda") for _ in range(4))
o, qp, kp, vp, o_pad, L = flash_attn_wmma.forward(q, k, v, 64, 128, False, D ** -0.5, False)
for _ in range(2):
    out = flash_attn_wmma.backward(qp, kp, vp, o_pad, d_o, L, N, N, D, 128, 128, False, D ** -0.5, False)
torch.cuda.synchronize()
print("done", out[0].float().abs().mean().item())
