"""Backward-kernel timing (secondary metric): dQ/dK/dV through flash_attn_wmma.backward on resident
tensors, CUDA-graph replay, TFLOPS = 2.5 * 4 B H N^2 D / t (bench_with_sdpa.py:39-40 credits the
backward with 2.5x the forward's FLOPs; causal x0.5).

    python tools/bench_bwd.py [N ...]        # default 1024 4096 16384
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flash-attention-v2-rdna3-minimal_b200"))
import torch
from rocwmma_fattn import _capi
from rocwmma_fattn.FlashAttn import flash_attn_wmma


ns = [int(x) for x in sys.argv[1:]] or [1024, 4096, 16384]
# FA_BWD=tc (serial kernel, round 1) | ws (pipelined kernel) | unset (library default)
_capi.set_bwd_kernel({"tc": _capi.FA_BWD_KERNEL_TC, "ws": _capi.FA_BWD_KERNEL_WS}.get(os.environ.get("FA_BWD", ""), 0))
H, D = 16, 128
torch.manual_seed(0)
res = {}
side = torch.cuda.Stream()
for dt_name, dt in (("f16", torch.float16),) if os.environ.get("FA_BWD_QUICK") else (("f16", torch.float16), ("bf16", torch.bfloat16)):
    for causal in (False, True):
        for n in ns:
            q, k, v, d_o = (torch.rand(1, H, n, D, dtype=dt, device="cuda") for _ in range(4))
            o, qp, kp, vp, o_pad, L = flash_attn_wmma.forward(q, k, v, 64, 128, causal, D ** -0.5, False)
            reps = max(2, min(16, int(1e12 / (10.0 * H * n * n * D)) + 1))

            def fn():
                return [flash_attn_wmma.backward(qp, kp, vp, o_pad, d_o, L, n, n, D, 128, 128, causal,
                                                 D ** -0.5, False) for _ in range(reps)]

            with torch.cuda.stream(side):
                fn()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                keep = fn()
            g.replay()
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                g.replay()
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / reps)
            fl = 2.5 * 4.0 * H * n * n * D * (0.5 if causal else 1.0)
            res[f"{dt_name}_{'causal' if causal else 'full'}_n{n}"] = {
                "ms": round(best, 4), "tflops": round(fl / best / 1e9, 1)}
            del keep, g
print(os.environ.get("FA_BWD", "default"), json.dumps({k: v["tflops"] for k, v in res.items()}))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"bench_bwd_{os.environ.get('FA_BWD', 'default')}.json"), "w"), indent=1)
