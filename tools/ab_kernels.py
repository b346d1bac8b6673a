"""Same-process A/B of forward kernel selectors (ws = one-shot, sk = persistent stream-K) on the sweep
shapes: CUDA-graph replay, alternating order, best of 5.   python tools/ab_kernels.py [N ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flash-attention-v2-rdna3-minimal_b200"))
import torch
from rocwmma_fattn import _capi
from rocwmma_fattn.FlashAttn import FlashAttentionFunction as F

ns = [int(x) for x in sys.argv[1:]] or [4096, 8192, 16384]
torch.manual_seed(0)
side = torch.cuda.Stream()
for n in ns:
    sets = max(2, min(12, (260 << 20) // (4 * 16 * n * 128 * 2) + 1))
    pool = [tuple(torch.rand(1, 16, n, 128, dtype=torch.float16, device="cuda") for _ in range(3)) for _ in range(sets)]
    graphs = {}
    for name, sel in (("ws", _capi.FA_KERNEL_WS), ("sk", _capi.FA_KERNEL_SK)):
        prev = _capi.set_kernel(sel)

        def fn():
            return [F.apply(*pool[i % sets], None, False) for i in range(sets)]

        with torch.cuda.stream(side):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            keep = fn()
        graphs[name] = (g, keep)
        _capi.set_kernel(prev)
    best = {"ws": 1e9, "sk": 1e9}
    for rep in range(5):
        for name in ("ws", "sk"):
            g = graphs[name][0]
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            best[name] = min(best[name], e0.elapsed_time(e1) / sets)
    fl = 4.0 * 16 * n * n * 128
    same = torch.equal(graphs["ws"][1][0], graphs["sk"][1][0])
    print(n, {k: (round(v, 5), round(fl / v / 1e9, 1)) for k, v in best.items()}, "bit-identical" if same else
          "max diff %.3e" % (graphs["ws"][1][0].float() - graphs["sk"][1][0].float()).abs().max().item(), flush=True)
    del graphs, pool
