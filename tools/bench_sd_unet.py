"""Attention stack of one Stable-Diffusion denoising step through the ComfyUI-shaped hook (rocwmma_fattn.hooks), next to
torch SDPA on the same box.  The reference reports end-to-end it/s inside ComfyUI (README.md:104-154); neither ComfyUI nor the
model weights exist here, so this times ONLY the attention calls one UNet evaluation makes (every self- and cross-attention
layer, batch 2 = cond + uncond, synthetic activations) and says so: `attn_ms_per_step` and its reciprocal, an upper bound on
it/s set by attention alone.  python tools/bench_sd_unet.py [--json out.json]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "flash-attention-v2-rdna3-minimal_b200"))
from rocwmma_fattn import hooks  # noqa: E402

# (tokens, heads, dim_head, transformer layers at this resolution); every layer = 1 self- + 1 cross-attention (77 text tokens)
MODELS = {
    # SD 1.5, 512x512 (latent 64x64): 2 down + 3 up blocks per level, 1 mid block
    "sd15_512": [(4096, 8, 40, 5), (1024, 8, 80, 5), (256, 8, 160, 5), (64, 8, 160, 1)],
    # SD 1.5, 1024x1024 (latent 128x128)
    "sd15_1024": [(16384, 8, 40, 5), (4096, 8, 80, 5), (1024, 8, 160, 5), (256, 8, 160, 1)],
    # SDXL, 1024x1024 (latent 128x128): 2 layers per block at 64x64 (5 blocks), 10 per block at 32x32 (5 blocks + mid)
    "sdxl_1024": [(4096, 10, 64, 10), (1024, 20, 64, 60)],
}
CTX = 77
BATCH = 2


def sdpa_hook(q, k, v, heads):
    """What ComfyUI's attention_pytorch does: [B, N, H*D] -> [B, H, N, D] views -> torch SDPA -> back."""
    b, n, inner = q.shape
    d = inner // heads
    q, k, v = (t.view(b, -1, heads, d).transpose(1, 2) for t in (q, k, v))
    o = torch.nn.functional.scaled_dot_product_attention(q, k, v)
    return o.transpose(1, 2).reshape(b, n, inner)


def build_stack(model, dtype):
    torch.manual_seed(0)
    layers = []
    for tokens, heads, d, n_layers in MODELS[model]:
        inner = heads * d
        x = torch.randn(BATCH, tokens, inner, dtype=dtype, device="cuda")
        xk, xv = (torch.randn(BATCH, tokens, inner, dtype=dtype, device="cuda") for _ in range(2))
        ck, cv = (torch.randn(BATCH, CTX, inner, dtype=dtype, device="cuda") for _ in range(2))
        layers.append((x, xk, xv, ck, cv, heads, n_layers))
    return layers


def run_stack(layers, attn):
    out = None
    for x, xk, xv, ck, cv, heads, n_layers in layers:
        for _ in range(n_layers):
            out = attn(x, xk, xv, heads)
            out = attn(x, ck, cv, heads)
    return out


def timed(fn, reps=20):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):  # warm up on the capture stream: per-stream workspaces are allocated outside capture
        for _ in range(3):
            fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        fn()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json")
    a = ap.parse_args()
    res = {"note": "attention calls of one UNet evaluation only (batch 2, 77 text tokens, synthetic fp16 activations), CUDA graph "
                   "replays, device events; not a full denoising step"}
    for model in MODELS:
        layers = build_stack(model, torch.float16)
        flops = sum(4.0 * BATCH * h * t * (t + CTX) * d * n for (t, h, d, n) in MODELS[model])
        o = run_stack(layers, hooks.comfy_attention)
        o_ref = run_stack(layers, sdpa_hook)
        err = (o.float() - o_ref.float()).abs().max().item()
        ms = timed(lambda: run_stack(layers, hooks.comfy_attention))
        ms_t = timed(lambda: run_stack(layers, sdpa_hook))
        res[model] = {"attn_calls": 2 * sum(n for *_, n in MODELS[model]), "attn_ms_per_step": round(ms, 4),
                      "attn_only_its": round(1e3 / ms, 1), "tflops": round(flops / ms / 1e9, 1),
                      "torch_sdpa_ms_per_step": round(ms_t, 4), "torch_sdpa_attn_only_its": round(1e3 / ms_t, 1),
                      "speedup_vs_torch_sdpa": round(ms_t / ms, 3), "max_abs_diff_vs_torch_sdpa_last_layer": err}
        print(model, json.dumps(res[model]), flush=True)
    if a.json:
        with open(a.json, "w") as fh:
            json.dump(res, fh, indent=1)


if __name__ == "__main__":
    main()
