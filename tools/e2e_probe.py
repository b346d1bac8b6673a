"""Where does the end-to-end (host buffer) sweep spend its time?  Prints raw pinned-memory PCIe rates
(H2D, D2H, both at once) and the per-sequence-length wall time of flash_attn_forward_host next to the
pure transfer time those bytes would need.  Usage: python tools/e2e_probe.py"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flash-attention-v2-rdna3-minimal_b200"))
from rocwmma_fattn import FlashAttn as FA  # noqa: E402


def rate(fn, nbytes, reps=10):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return nbytes * reps / (time.perf_counter() - t0) / 1e9


def main():
    dev = torch.device("cuda:0")
    n = 256 << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(n, dtype=torch.uint8, device=dev)
    d_b = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    print("H2D  %.1f GB/s" % rate(lambda: d_a.copy_(h_in, non_blocking=True), n))
    print("D2H  %.1f GB/s" % rate(lambda: h_out.copy_(d_b, non_blocking=True), n))

    def both():
        with torch.cuda.stream(s1):
            d_a.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_b, non_blocking=True)

    print("H2D+D2H together  %.1f GB/s total (each direction moves 256 MiB)" % rate(both, 2 * n))
    for sz in (1 << 20, 4 << 20, 16 << 20):
        print("H2D %3d MiB copies  %.1f GB/s" % (sz >> 20, rate(lambda: d_a[:sz].copy_(h_in[:sz], non_blocking=True), sz, 40)))

    H, D = 16, 128
    h2d_peak = rate(lambda: d_a.copy_(h_in, non_blocking=True), n)
    for N in (512, 1024, 2048, 4096, 8192, 16384):
        q, k, v = (torch.rand((1, H, N, D), dtype=torch.float16).pin_memory() for _ in range(3))
        o = torch.empty_like(q).pin_memory()
        def timed(reps=20):
            for _ in range(3):
                FA.flash_attn_forward_host(q, k, v, out=o)
            t0 = time.perf_counter()
            for _ in range(reps):
                FA.flash_attn_forward_host(q, k, v, out=o)
            return (time.perf_counter() - t0) / reps * 1e3

        os.environ.pop("FA_HOST_CHUNKS", None)
        ms = timed()
        nb = 3 * q.numel() * 2
        sweep = []
        for nc in (1, 2, 3, 4, 6, 8, 16):
            os.environ["FA_HOST_CHUNKS"] = str(nc)
            sweep.append("%d:%.3f" % (nc, timed(10)))
        os.environ.pop("FA_HOST_CHUNKS", None)
        print("N=%5d  host call %.3f ms   H2D bytes %.1f MB -> %.3f ms at the measured H2D rate (%.0f%%)   chunks:ms %s"
              % (N, ms, nb / 1e6, nb / h2d_peak / 1e6, 100 * nb / h2d_peak / 1e6 / ms, " ".join(sweep)))


if __name__ == "__main__":
    main()
