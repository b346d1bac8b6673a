"""Forward throughput at head dims > 128 (fa_fwd_wide.cuh) next to torch SDPA on the same tensors.
Not a bench.py metric: a development table for DESIGN.md.  python tools/bench_wide.py
(BENCH_D=40,64,80 BENCH_N=1024,4096 selects other head dims / lengths)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flash-attention-v2-rdna3-minimal_b200"))
from rocwmma_fattn import _capi  # noqa: E402
from rocwmma_fattn.FlashAttn import FlashAttentionFunction as F  # noqa: E402


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if os.environ.get("BENCH_GRAPH"):  # launch-bound sizes: time a CUDA graph of 20 calls instead
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            keep = [fn() for _ in range(20)]
        g.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(max(1, iters // 4)):
            g.replay()
        b.record()
        torch.cuda.synchronize()
        del keep
        return a.elapsed_time(b) / (20 * max(1, iters // 4))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


DS = tuple(int(x) for x in os.environ.get("BENCH_D", "160,192,256").split(","))
NS = tuple(int(x) for x in os.environ.get("BENCH_N", "2048,4096,8192,16384").split(","))


def main():
    H = 16
    forced = os.environ.get("BENCH_KERNEL")
    if forced:
        _capi.set_kernel({v: k for k, v in _capi.KERNEL_NAMES.items()}[forced])
    if os.environ.get("BENCH_PAIRS"):
        _capi.lib.fa_set_wide_pairs(int(os.environ["BENCH_PAIRS"]))
    for dtype in (torch.float16, torch.bfloat16):
        for causal in (False, True):
            for D in DS:
                for N in NS:
                    q, k, v = (torch.rand((1, H, N, D), dtype=dtype, device="cuda") for _ in range(3))
                    fl = 4.0 * H * N * N * D * (0.5 if causal else 1.0)
                    iters = max(5, min(50, int(3e12 / fl)))
                    ms = timed(lambda: F.apply(q, k, v, None, causal), iters)
                    try:
                        ms_t = timed(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=causal), iters)
                    except Exception:  # noqa: BLE001
                        ms_t = float("nan")
                    print("%s causal=%d D=%3d N=%5d  ours %8.3f ms %7.1f TFLOPS   torch SDPA %8.3f ms %7.1f TFLOPS"
                          % (str(dtype)[6:], causal, D, N, ms, fl / ms / 1e9, ms_t, fl / ms_t / 1e9), flush=True)
    print("launches", _capi.launch_count())


if __name__ == "__main__":
    main()
