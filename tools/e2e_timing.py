"""Per-call, per-chunk timeline of the host-buffer path in bench.py's e2e order (FA_HOST_TIMING=1
makes the C-ABI print it to stderr)."""
import os
import sys

os.environ["FA_HOST_TIMING"] = "1"
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flash-attention-v2-rdna3-minimal_b200"))
from rocwmma_fattn import FlashAttn as FA  # noqa: E402

NS = (512, 1024, 2048, 4096, 8192, 16384)
host = {}
for n in NS:
    host[n] = tuple(torch.rand((1, 16, n, 128), dtype=torch.float16).pin_memory() for _ in range(3)) + (
        torch.empty((1, 16, n, 128), dtype=torch.float16).pin_memory(),)
for step in range(4):
    print("---- sweep %d" % step, file=sys.stderr, flush=True)
    for n in NS:
        q, k, v, o = host[n]
        FA.flash_attn_forward_host(q, k, v, out=o)
