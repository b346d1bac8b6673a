import os, sys
sys.path.insert(0, os.path.join(os.environ["GRAFT_REPO_ROOT"], "flash-attention-v2-rdna3-minimal_b200"))
import torch
from rocwmma_fattn import _capi
from rocwmma_fattn.FlashAttn import FlashAttentionFunction as F
torch.manual_seed(0)
for name in ("ws3", "ws2", "sk"):
    _capi.set_kernel({v: k for k, v in _capi.KERNEL_NAMES.items()}[name])
    for (B, H, N, Nkv, D, dt, causal) in [(1, 2, 768, 640, 128, torch.float16, False), (1, 1, 300, 900, 64, torch.bfloat16, False),
                                          (1, 1, 640, 640, 128, torch.bfloat16, True)]:
        q, k, v = (torch.randn(B, H, n, D, dtype=dt, device="cuda") for n in (N, Nkv, Nkv))
        o = F.apply(q, k, v, None, causal)
        torch.cuda.synchronize()
        print("ok", name, B, H, N, Nkv, D, dt, causal, float(o.float().mean()))
