#!/bin/bash
# compute-sanitizer over small forward + backward problems (memcheck, racecheck, synccheck)
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import os, sys
sys.path.insert(0, os.path.join(os.environ["GRAFT_REPO_ROOT"], "flash-attention-v2-rdna3-minimal_b200"))
import torch
from rocwmma_fattn.FlashAttn import FlashAttentionFunction as F
torch.manual_seed(0)
for (B, H, N, Nkv, D, dt, causal) in [(1, 2, 512, 512, 128, torch.float16, False), (1, 1, 300, 200, 128, torch.bfloat16, True),
                                      (1, 2, 128, 128, 64, torch.float16, False), (1, 1, 384, 384, 64, torch.bfloat16, True),
                                      # head dims 129..256: wide forward + the three-launch tcgen05 backward (fa_bwd_wide.cuh)
                                      (1, 1, 320, 200, 160, torch.float16, True), (1, 2, 256, 384, 256, torch.bfloat16, False)]:
    q = torch.rand(B, H, N, D, dtype=dt, device="cuda", requires_grad=True)
    k = torch.rand(B, H, Nkv, D, dtype=dt, device="cuda", requires_grad=True)
    v = torch.rand(B, H, Nkv, D, dtype=dt, device="cuda", requires_grad=True)
    o = F.apply(q, k, v, None, causal)
    o.backward(torch.rand_like(o))
    torch.cuda.synchronize()
    print("ok", B, H, N, Nkv, D, dt, causal, float(o.float().mean()), float(q.grad.float().abs().mean()))
PY
cat > /tmp/san_wide.py <<'PY'
import os, sys
sys.path.insert(0, os.path.join(os.environ["GRAFT_REPO_ROOT"], "flash-attention-v2-rdna3-minimal_b200"))
import torch
from rocwmma_fattn.FlashAttn import FlashAttentionFunction as F
torch.manual_seed(0)
# head dims 129..256: forward only (fa_fwd_wide_kernel), several KV tiles so the ring and both S buffers wrap
# (head dims 193..256 run on CTA pairs: cluster of two, cta_group::2; 640 rows = 5 tiles = an odd, padded pair)
for (B, H, N, Nkv, D, dt, causal) in [(1, 2, 640, 640, 256, torch.float16, False), (1, 1, 300, 700, 160, torch.bfloat16, True),
                                      (1, 1, 512, 512, 192, torch.float16, True), (1, 1, 640, 640, 232, torch.bfloat16, True)]:
    q, k, v = (torch.randn(B, H, n, D, dtype=dt, device="cuda") for n in (N, Nkv, Nkv))
    o = F.apply(q, k, v, None, causal)
    torch.cuda.synchronize()
    print("ok", B, H, N, Nkv, D, dt, causal, float(o.float().mean()))
PY
cat >> /tmp/san_wide.py <<'PY'
# the two-tile kernel on CTA pairs (forced: auto picks it for long non-causal problems only)
from rocwmma_fattn import _capi
_capi.set_kernel(_capi.FA_KERNEL_WS2)
for (B, H, N, Nkv, D, dt) in [(1, 2, 768, 640, 128, torch.float16), (1, 1, 300, 900, 64, torch.bfloat16)]:
    q, k, v = (torch.randn(B, H, n, D, dtype=dt, device="cuda") for n in (N, Nkv, Nkv))
    o = F.apply(q, k, v, None, False)
    torch.cuda.synchronize()
    print("ok ws2", B, H, N, Nkv, D, dt, float(o.float().mean()))
_capi.set_kernel(_capi.FA_KERNEL_AUTO)
PY
if [ -n "$SAN_ONLY_WIDE" ]; then cp /tmp/san_wide.py /tmp/san.py; else cat /tmp/san_wide.py >> /tmp/san.py; fi
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --kernel-regex kns=fa_ python /tmp/san.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard" gpurun_out/sanitizer_$tool.log | head -8; grep -c "^ok" gpurun_out/sanitizer_$tool.log
done
