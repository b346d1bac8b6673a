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

KERNELS = os.environ.get("SD_KERNELS", "auto,ws").split(",")
if os.environ.get("SD_SHAPES"):  # "B,H,Nq,Nkv,D;B,H,..." replaces the built-in list
    SHAPES_OVERRIDE = [tuple(int(x) for x in item.split(",")) for item in os.environ["SD_SHAPES"].split(";")]
else:
    SHAPES_OVERRIDE = None
SHAPES = [  # B, H, Nq, Nkv, D
    (2, 10, 4096, 4096, 64), (2, 10, 4096, 77, 64), (2, 20, 1024, 1024, 64), (2, 20, 1024, 77, 64),
    (2, 8, 4096, 4096, 40), (2, 8, 4096, 77, 40), (2, 8, 1024, 1024, 80), (2, 8, 256, 256, 160),
    (2, 10, 4096, 77, 128), (1, 10, 4096, 4096, 64), (2, 8, 16384, 16384, 40), (2, 8, 16384, 77, 40),
]


def timed(fn):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):  # warm up on the capture stream: per-stream workspaces are allocated outside capture
        for _ in range(3):
            fn()
    torch.cuda.synchronize()
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


for (B, H, Nq, Nkv, D) in (SHAPES_OVERRIDE or SHAPES):
    q = torch.rand((B, H, Nq, D), dtype=torch.float16, device="cuda")
    k, v = (torch.rand((B, H, Nkv, D), dtype=torch.float16, device="cuda") for _ in range(2))
    fl = 4.0 * B * H * Nq * Nkv * D
    row = []
    for name in KERNELS:
        prev = _capi.set_kernel({"auto": _capi.FA_KERNEL_AUTO, "ws": _capi.FA_KERNEL_WS, "ws3": _capi.FA_KERNEL_WS3,
                                 "sk": _capi.FA_KERNEL_SK, "wide": _capi.FA_KERNEL_WIDE}[name])
        try:
            ms = timed(lambda: F.apply(q, k, v, None, False))
            row.append("%s %.4f ms %5.1f TF" % (name, ms, fl / ms / 1e9))
        except RuntimeError as e:  # forced kernel cannot serve the head dim
            row.append("%s n/a" % name)
        finally:
            _capi.set_kernel(prev)
    ms_t = timed(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v))
    DP = D + (-D) % 8
    sel = _capi.KERNEL_NAMES[_capi.select_kernel(B, H, Nq, Nkv, DP, (H * Nq * DP, Nq * DP, DP, 1), (H * Nkv * DP, Nkv * DP, DP, 1),
                                                  (H * Nkv * DP, Nkv * DP, DP, 1), (H * Nq * DP, Nq * DP, DP, 1), 0, False, D ** -0.5)]
    print("%-26s -> %-5s | %s | torch SDPA %.4f ms" % ((B, H, Nq, Nkv, D), sel, " | ".join(row), ms_t), flush=True)
