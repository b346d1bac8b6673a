import os, sys, json
sys.path.insert(0, os.path.join(os.environ.get("GRAFT_REPO_ROOT", "/root/repo"), "flash-attention-v2-rdna3-minimal_b200"))
import torch
from rocwmma_fattn.FlashAttn import FlashAttentionFunction as F
torch.manual_seed(0)
res = {}
for (H, N, D) in [(16, 4096, 64), (16, 16384, 64), (16, 4096, 160), (16, 4096, 256), (8, 4096, 40)]:
    q, k, v = (torch.rand(1, H, N, D, dtype=torch.float16, device="cuda", requires_grad=True) for _ in range(3))
    d_o = torch.rand(1, H, N, D, dtype=torch.float16, device="cuda")
    o = F.apply(q, k, v, None, False)
    for _ in range(2):
        torch.autograd.grad(o, (q, k, v), d_o, retain_graph=True)
    torch.cuda.synchronize()
    reps = 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        torch.autograd.grad(o, (q, k, v), d_o, retain_graph=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    res[f"h{H}_n{N}_d{D}"] = {"ms": round(ms, 3), "tflops": round(2.5 * 4 * H * N * N * D / ms / 1e9, 1)}
print(json.dumps(res))
