"""Library reference point for the backward (context only, NOT this repo's code): torch SDPA backward (cuDNN / flash
backend) on the same shapes as tools/bench_bwd.py, timed with CUDA events over repeated .backward(retain_graph=True).

    python tools/bench_bwd_lib.py [N ...]
"""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ns = [int(x) for x in sys.argv[1:]] or [1024, 4096, 16384]
H, D = 16, 128
res = {}
for dt_name, dt in (("f16", torch.float16), ("bf16", torch.bfloat16)):
    for causal in (False, True):
        for n in ns:
            q, k, v = (torch.rand(1, H, n, D, dtype=dt, device="cuda", requires_grad=True) for _ in range(3))
            d_o = torch.rand(1, H, n, D, dtype=dt, device="cuda")
            o = F.scaled_dot_product_attention(q, k, v, is_causal=causal)
            reps = max(2, min(16, int(1e12 / (10.0 * H * n * n * D)) + 1))
            for _ in range(2):
                torch.autograd.grad(o, (q, k, v), d_o, retain_graph=True)
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    torch.autograd.grad(o, (q, k, v), d_o, retain_graph=True)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / reps)
            fl = 2.5 * 4.0 * H * n * n * D * (0.5 if causal else 1.0)
            res[f"{dt_name}_{'causal' if causal else 'full'}_n{n}"] = {"ms": round(best, 4), "tflops": round(fl / best / 1e9, 1)}
print("library", json.dumps({k: v["tflops"] for k, v in res.items()}))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "bench_bwd_library.json"), "w"), indent=1)
