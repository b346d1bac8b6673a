#!/bin/bash
# compute-sanitizer over the persistent (stream-K) forward kernel and the early-S kernel at head dim 64: the kernels the cost
# model hands the Stable-Diffusion shapes to (round 2, third session)
mkdir -p gpurun_out
cat > /tmp/san_sk.py <<'PY'
import os, sys
sys.path.insert(0, os.path.join(os.environ["GRAFT_REPO_ROOT"], "flash-attention-v2-rdna3-minimal_b200"))
import torch
from rocwmma_fattn import _capi
from rocwmma_fattn.FlashAttn import FlashAttentionFunction as F
torch.manual_seed(0)
def ref(q, k, v):
    s = (q.float() @ k.float().transpose(-1, -2)) * q.shape[-1] ** -0.5
    return s.softmax(-1) @ v.float()
cases = [("sk", 1, 2, 512, 512, 128, torch.float16),      # 4 units x 4 tiles on 16 CTAs: every unit split, 3 partials each
         ("sk", 1, 3, 768, 300, 64, torch.bfloat16),      # 9 units x 3 tiles (ragged last tile) on 27 CTAs
         ("sk", 1, 40, 1024, 77, 64, torch.float16),      # 160 one-tile units on 148 CTAs: whole units, unit boundaries
         ("sk", 2, 19, 1024, 256, 128, torch.float16),    # 152 units x 2 tiles: boundaries AND splits
         ("ws3", 1, 4, 1024, 1024, 64, torch.float16), ("ws3", 1, 2, 768, 1000, 40, torch.bfloat16)]
for (kern, B, H, N, Nkv, D, dt) in cases:
    q, k, v = (torch.randn(B, H, n, D, dtype=dt, device="cuda") for n in (N, Nkv, Nkv))
    prev = _capi.set_kernel({"sk": _capi.FA_KERNEL_SK, "ws3": _capi.FA_KERNEL_WS3}[kern])
    o = F.apply(q, k, v, None, False)
    o2 = F.apply(q, k, v, None, False)   # second launch: flags lowered again
    _capi.set_kernel(prev)
    torch.cuda.synchronize()
    err = (o.float() - ref(q, k, v)).abs().max().item()
    print("ok", kern, B, H, N, Nkv, D, dt, "max err %.2e" % err, "same bits", bool(torch.equal(o, o2)))
PY
for tool in memcheck synccheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --kernel-regex kns=fa_ python /tmp/san_sk.py > gpurun_out/sanitizer_sk_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard" gpurun_out/sanitizer_sk_$tool.log | head -8; grep -c "^ok" gpurun_out/sanitizer_sk_$tool.log
done
grep "^ok" gpurun_out/sanitizer_sk_memcheck.log
