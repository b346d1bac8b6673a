"""Backward timing at one arbitrary shape (graph replay, best of 3): python tools/bench_bwd_shape.py B H N D [causal]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flash-attention-v2-rdna3-minimal_b200"))
import torch
from rocwmma_fattn.FlashAttn import flash_attn_wmma

B, H, N, D = (int(x) for x in sys.argv[1:5])
causal = len(sys.argv) > 5 and sys.argv[5] == "causal"
torch.manual_seed(0)
q, k, v, d_o = (torch.rand(B, H, N, D, dtype=torch.float16, device="cuda") for _ in range(4))
o, qp, kp, vp, o_pad, L = flash_attn_wmma.forward(q, k, v, 64, 128, causal, D ** -0.5, False)
reps = 4
side = torch.cuda.Stream()


def fn():
    return [flash_attn_wmma.backward(qp, kp, vp, o_pad, d_o, L, N, N, D, 128, 128, causal, D ** -0.5, False) for _ in range(reps)]


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
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / reps)
fl = 2.5 * 4.0 * B * H * N * N * D * (0.5 if causal else 1.0)
print(os.environ.get("FA_FWD_SM100_LIB", "default").split("_")[-1], B, H, N, D, causal, "ms", round(best, 4), "TFLOPS", round(fl / best / 1e9, 1))
