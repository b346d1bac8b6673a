"""Forward time at the attention shapes Stable Diffusion makes (the reference's use case, README.md:104-154):
self- and cross-attention (Nkv = 77) of SDXL (head dim 64) and SD 1.5 (head dims 40 / 80 / 160), batch 2 (cfg),
CUDA-graph timed, auto kernel choice next to the two-tile kernel forced.  python tools/bench_sd_shapes.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flash-attention-v2-rdna3-minimal_b200"))
from rocwmma_fattn import _capi  # noqa: E402
from rocwmma_fattn.FlashAttn import FlashAttentionFunction as F  # noqa: E402

SHAPES = [  # B, H, Nq, Nkv, D
    (2, 10, 4096, 4096, 64), (2, 10, 4096, 77, 64), (2, 20, 1024, 1024, 64), (2, 20, 1024, 77, 64),
    (2, 8, 4096, 4096, 40), (2, 8, 4096, 77, 40), (2, 8, 1024, 1024, 80), (2, 8, 256, 256, 160),
]


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        keep = [fn() for _ in range(20)]
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    del keep
    return a.elapsed_time(b) / 100


for (B, H, Nq, Nkv, D) in SHAPES:
    q = torch.rand((B, H, Nq, D), dtype=torch.float16, device="cuda")
    k, v = (torch.rand((B, H, Nkv, D), dtype=torch.float16, device="cuda") for _ in range(2))
    fl = 4.0 * B * H * Nq * Nkv * D
    row = []
    for name in ("auto", "ws"):
        prev = _capi.set_kernel({"auto": _capi.FA_KERNEL_AUTO, "ws": _capi.FA_KERNEL_WS}[name])
        try:
            ms = timed(lambda: F.apply(q, k, v, None, False))
            row.append("%s %.4f ms %6.1f TFLOPS" % (name, ms, fl / ms / 1e9))
        except RuntimeError as e:  # forced kernel cannot serve the head dim
            row.append("%s n/a" % name)
        finally:
            _capi.set_kernel(prev)
    ms_t = timed(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v))
    DP = D + (-D) % 8
    sel = _capi.KERNEL_NAMES[_capi.select_kernel(B, H, Nq, Nkv, DP, (H * Nq * DP, Nq * DP, DP, 1), (H * Nkv * DP, Nkv * DP, DP, 1),
                                                  (H * Nkv * DP, Nkv * DP, DP, 1), (H * Nq * DP, Nq * DP, DP, 1), 0, False, D ** -0.5)]
    print("%-26s -> %-5s | %s | %s | torch SDPA %.4f ms" % ((B, H, Nq, Nkv, D), sel, row[0], row[1], ms_t), flush=True)
