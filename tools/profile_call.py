"""cProfile of the Python side of one forward call (launch-bound shape).  python tools/profile_call.py"""
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flash-attention-v2-rdna3-minimal_b200"))
from rocwmma_fattn.FlashAttn import FlashAttentionFunction as F, flash_attn_forward  # noqa: E402

q, k, v = (torch.rand((1, 16, 512, 128), dtype=torch.float16, device="cuda") for _ in range(3))
for _ in range(20):
    F.apply(q, k, v, None, False)
torch.cuda.synchronize()


def loop(fn, n=3000):
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


print("apply            %.1f us/call" % loop(lambda: F.apply(q, k, v, None, False)))
print("functional       %.1f us/call" % loop(lambda: flash_attn_forward(q, k, v)))
print("torch SDPA       %.1f us/call" % loop(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v)))
print("torch.empty_like %.1f us/call" % loop(lambda: torch.empty_like(q)))
pr = cProfile.Profile()
pr.enable()
for _ in range(3000):
    F.apply(q, k, v, None, False)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
