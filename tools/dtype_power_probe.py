"""Why is the fp16 forward 15 % slower than the bf16 forward with the same instruction stream?  

Hypothesis: the part is power-capped (sw_power_cap is active in every bench run) and what the tensor cores and
the operand paths burn per MMA depends on how many bits toggle: U[0,1) values carry 11 random significand bits in
fp16 and 8 in bf16, and so do the probabilities P.  Test: run the SAME kernel (N=16384, B=1, H=16, D=128) for
~1.5 s on
  f16_rand      U[0,1) fp16 (the bench input)
  f16_as_bf16   the same values rounded to bf16 precision, stored as fp16 (fp16 kernel, bf16-like bit patterns)
  f16_const     all 0.5 (nothing toggles)
  bf16_rand     U[0,1) bf16
  lib:*         the same inputs through torch SDPA (cuDNN fused attention), for its clock and cycle efficiency
and sample nvidia-smi (SM clock, power) during each.  If the clock, not the cycle count, explains the gap, the
TFLOPS ratio follows the clock ratio and f16_as_bf16 lands between.

    python tools/dtype_power_probe.py            # on the GPU box; writes gpurun_out/dtype_power_probe.json
"""
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flash-attention-v2-rdna3-minimal_b200"))
import torch  # noqa: E402

from rocwmma_fattn.FlashAttn import FlashAttentionFunction  # noqa: E402

fa = FlashAttentionFunction.apply
N, H, D = 16384, 16, 128
torch.manual_seed(0)


class Smi:
    def __init__(self):
        self.lines = []
        self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits",
                                      "-lms", "50", "-i", "0"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        threading.Thread(target=self._read, daemon=True).start()

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def window(self, t0, t1):
        clk, pw = [], []
        for ts, ln in self.lines:
            if t0 + 0.3 <= ts <= t1:
                a, b = (x.strip() for x in ln.split(","))
                clk.append(float(a))
                pw.append(float(b))
        return (statistics.median(clk) if clk else None, statistics.median(pw) if pw else None, len(clk))


def make(kind):
    base = torch.rand((1, H, N, D), device="cuda")
    if kind == "f16_rand":
        return base.to(torch.float16)
    if kind == "f16_as_bf16":
        return base.to(torch.bfloat16).to(torch.float16)
    if kind == "f16_const":
        return torch.full((1, H, N, D), 0.5, device="cuda", dtype=torch.float16)
    return base.to(torch.bfloat16)


smi = Smi()
time.sleep(0.5)
out = {}
def lib(q, k, v, _m, causal):
    return torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=causal)


for kind in ("f16_rand", "f16_as_bf16", "f16_const", "bf16_rand", "f16_rand", "lib:f16_rand", "lib:bf16_rand", "lib:f16_const"):
    fa = lib if kind.startswith("lib:") else FlashAttentionFunction.apply
    sets = [tuple(make(kind.split(":")[-1]) for _ in range(3)) for _ in range(2)]
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        [fa(*sets[i % 2], None, False) for i in range(4)]
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=side):
        keep = [fa(*sets[i % 2], None, False) for i in range(4)]
    g.replay()
    torch.cuda.synchronize()
    time.sleep(1.0)  # let the part cool to the same starting point
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    reps = 240
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    ms = e0.elapsed_time(e1) / (4 * reps)
    clk, pw, ns = smi.window(t0, t1)
    key = kind if kind not in out else kind + "_again"
    out[key] = {"ms": round(ms, 4), "tflops": round(4.0 * H * N * N * D / ms / 1e9, 1), "sm_mhz_median": clk,
                "power_w_median": pw, "samples": ns,
                "tflops_per_ghz": round(4.0 * H * N * N * D / ms / 1e9 / (clk / 1e3), 1) if clk else None}
    print(key, out[key], flush=True)
    del keep, g, sets
smi.proc.terminate()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "dtype_power_probe.json"), "w") as fh:
    json.dump(out, fh, indent=1)
