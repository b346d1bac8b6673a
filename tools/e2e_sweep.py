"""bench.py's e2e loop on its own: the c2 sweep (B=1 H=16 D=128 fp16, N = 512..16384) through
flash_attn_forward_host, pinned host buffers, wall clock per sweep.  FA_HOST_CHUNKS overrides the
chunk planner (see fa_capi.cu) so planners can be compared on one box."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flash-attention-v2-rdna3-minimal_b200"))
from rocwmma_fattn import FlashAttn as FA  # noqa: E402

NS = (512, 1024, 2048, 4096, 8192, 16384)
host = {}
for n in NS:
    host[n] = tuple(torch.rand((1, 16, n, 128), dtype=torch.float16).pin_memory() for _ in range(3)) + (
        torch.empty((1, 16, n, 128), dtype=torch.float16).pin_memory(),)
h2d = sum(3 * 16 * n * 128 * 2 for n in NS)


def sweep(steps=10):
    for _ in range(2):
        for n in NS:
            q, k, v, o = host[n]
            FA.flash_attn_forward_host(q, k, v, out=o)
    t0 = time.perf_counter()
    for _ in range(steps):
        for n in NS:
            q, k, v, o = host[n]
            FA.flash_attn_forward_host(q, k, v, out=o)
    return (time.perf_counter() - t0) / steps * 1e3


for rep in range(2):
    for setting in sys.argv[1:] or ["auto"]:
        if setting == "auto":
            os.environ.pop("FA_HOST_CHUNKS", None)
        else:
            os.environ["FA_HOST_CHUNKS"] = setting
        ms = sweep()
        print("chunks=%-5s  %.3f ms per sweep   H2D %.1f GB/s" % (setting, ms, h2d / ms / 1e6), flush=True)
