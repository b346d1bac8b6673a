"""A/B timing of library variants (lib/libfa_fwd_sm100_<name>.so): one subprocess per variant so each
loads its own build through FA_FWD_SM100_LIB.  Prints ms and TFLOPS per sequence length.

    python tools/ab_bench.py emu0 emu2 emu4 [-- N N ...]
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "flash-attention-v2-rdna3-minimal_b200")

CHILD = r'''
import os, sys, json, time
sys.path.insert(0, os.environ["FA_PKG"])
import torch
from rocwmma_fattn import _capi
from rocwmma_fattn.FlashAttn import FlashAttentionFunction as F
if os.environ.get("FA_KERNEL"):
    _capi.set_kernel({"ws": _capi.FA_KERNEL_WS, "sk": _capi.FA_KERNEL_SK, "ws2": _capi.FA_KERNEL_WS2, "ws3": _capi.FA_KERNEL_WS3,
                      "wide": _capi.FA_KERNEL_WIDE}[os.environ["FA_KERNEL"]])
ns = [int(x) for x in os.environ["FA_NS"].split(",")]
causal = os.environ.get("FA_CAUSAL", "0") == "1"
dt = torch.bfloat16 if os.environ.get("FA_BF16", "0") == "1" else torch.float16
torch.manual_seed(0)
res = {}
side = torch.cuda.Stream()
for n in ns:
    sets = max(2, min(16, (260 << 20) // (4 * 16 * n * 128 * 2) + 1))
    pool = [tuple(torch.rand(1, 16, n, 128, dtype=dt, device="cuda") for _ in range(3)) for _ in range(sets)]
    reps = max(sets, min(64, int(2e12 / (4.0 * 16 * n * n * 128)) + 1))
    def fn():
        return [F.apply(*pool[i % sets], None, causal) for i in range(reps)]
    with torch.cuda.stream(side):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        keep = fn()
    fl = 4.0 * 16 * n * n * 128 * (0.5 if causal else 1.0) * reps
    def timed():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1)
    time.sleep(0.5)            # let the clocks recover
    timed()
    burst = min(timed() for _ in range(3))
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < 1.0:   # heat up
        g.replay()
    torch.cuda.synchronize()
    ts = [timed() for _ in range(max(3, int(300 / max(burst, 1e-3))))][-50:]
    sus = sum(ts) / len(ts)
    res[n] = {"burst_tf": round(fl / burst / 1e9, 1), "sustained_tf": round(fl / sus / 1e9, 1),
              "burst_ms": round(burst / reps, 5)}
    del keep, g, pool
print(json.dumps(res))
'''

args = sys.argv[1:]
ns = [4096, 16384]
if "--" in args:
    k = args.index("--")
    ns = [int(x) for x in args[k + 1:]]
    args = args[:k]
out = {}
for name in args:
    lib = os.path.join(PKG, "lib", "libfa_fwd_sm100.so" if name == "default" else f"libfa_fwd_sm100_{name}.so")
    env = dict(os.environ, FA_FWD_SM100_LIB=lib, FA_PKG=PKG, FA_NS=",".join(map(str, ns)))
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-400:]
    print(name, line, flush=True)
    out[name] = line
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "ab_bench.json"), "a") as fh:
    fh.write(json.dumps(out) + "\n")
