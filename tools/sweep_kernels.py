"""Per-sequence-length A/B of the forward kernels (forced through fa_set_kernel) with PDL on and off, next to
the library kernel torch dispatches to (cuDNN / flash SDPA) on the same box: the data the cost model in
fa_capi.cu (estimate_costs) is fitted to.

    python tools/sweep_kernels.py [--dtype f16|bf16] [--causal] [--kernels auto,ws,sk,wide,ws2] [--ns 512,1024,...]
                                  [--heads 16] [--dim 128] [--batch 1] [--out gpurun_out/sweep.json]

Timing: `reps` launches on rotating inputs (> 2 x L2 in total) captured into one CUDA graph, median of 5 replays,
CUDA events on the replay stream.  All kernels of one N are timed back to back so that they see the same clocks.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flash-attention-v2-rdna3-minimal_b200"))
import torch  # noqa: E402

from rocwmma_fattn import _capi  # noqa: E402
from rocwmma_fattn.FlashAttn import FlashAttentionFunction  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--dtype", default="f16")
ap.add_argument("--causal", action="store_true")
ap.add_argument("--kernels", default="auto,ws,sk,wide,ws2")
ap.add_argument("--ns", default="512,1024,2048,4096,8192,16384")
ap.add_argument("--heads", type=int, default=16)
ap.add_argument("--dim", type=int, default=128)
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--pdl", default="1,0")
ap.add_argument("--lib", action="store_true", help="also time torch SDPA (cuDNN / flash backend)")
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep_kernels.json"))
a = ap.parse_args()

dt = torch.bfloat16 if a.dtype == "bf16" else torch.float16
names = {v: k for k, v in _capi.KERNEL_NAMES.items()}
fa = FlashAttentionFunction.apply
side = torch.cuda.Stream()
torch.manual_seed(0)


def timed_graph(fn, reps):
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        keep = fn()
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / reps)
    del keep, g
    return sorted(ts)[2]


res = {"dtype": a.dtype, "causal": a.causal, "B": a.batch, "H": a.heads, "D": a.dim, "per_n": {}}
for n in [int(x) for x in a.ns.split(",")]:
    shape = (a.batch, a.heads, n, a.dim)
    per_set = 4 * a.batch * a.heads * n * a.dim * 2
    sets = max(2, min(64, (260 << 20) // per_set + 1))
    pool = [tuple(torch.rand(shape, dtype=dt, device="cuda") for _ in range(3)) for _ in range(sets)]
    fl = 4.0 * a.batch * a.heads * n * n * a.dim * (0.5 if a.causal else 1.0)
    reps = max(sets, min(64, int(3e12 / fl) + 1))
    row = {}
    for kname in a.kernels.split(","):
        for pdl in [int(x) for x in a.pdl.split(",")]:
            _capi.set_kernel(names[kname])
            _capi.set_pdl(bool(pdl))
            try:
                ms = timed_graph(lambda: [fa(*pool[i % sets], None, a.causal) for i in range(reps)], reps)
                row[f"{kname}{'' if pdl else '_nopdl'}"] = {"us": round(ms * 1e3, 2), "tflops": round(fl / ms / 1e9, 1)}
            except Exception as exc:  # noqa: BLE001 - a kernel that cannot serve the shape
                row[f"{kname}{'' if pdl else '_nopdl'}"] = {"error": str(exc)[:120]}
    _capi.set_kernel(_capi.FA_KERNEL_AUTO)
    _capi.set_pdl(True)
    st = (a.heads * n * a.dim, n * a.dim, a.dim, 1)
    row["auto_is"] = _capi.KERNEL_NAMES[_capi.select_kernel(a.batch, a.heads, n, n, a.dim, st, st, st, st,
                                                            1 if a.dtype == "bf16" else 0, a.causal, a.dim ** -0.5)]
    if a.lib:
        try:
            ms = timed_graph(lambda: [torch.nn.functional.scaled_dot_product_attention(*pool[i % sets], is_causal=a.causal)
                                      for i in range(reps)], reps)
            row["torch_sdpa"] = {"us": round(ms * 1e3, 2), "tflops": round(fl / ms / 1e9, 1)}
        except Exception as exc:  # noqa: BLE001
            row["torch_sdpa"] = {"error": str(exc)[:120]}
    res["per_n"][str(n)] = row
    print(n, json.dumps(row), flush=True)
    del pool
    torch.cuda.empty_cache()

os.makedirs(os.path.dirname(a.out), exist_ok=True)
with open(a.out, "a") as fh:
    fh.write(json.dumps(res) + "\n")
