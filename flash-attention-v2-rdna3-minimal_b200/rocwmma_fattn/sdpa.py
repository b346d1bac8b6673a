"""``torch.nn.functional.scaled_dot_product_attention``-shaped front end.

The reference is consumed in practice through two external plugins (README.md:35-37: ComfyUI and
sd-webui) that replace torch's SDPA with ``FlashAttentionFunction.apply`` for the calls the kernel can
serve.  This module is that adapter for the B200 build:

    from rocwmma_fattn.sdpa import scaled_dot_product_attention      # same signature as torch's
    o = scaled_dot_product_attention(q, k, v, is_causal=True)          # q, k, v: [B, H, N, D] (or [H, N, D])

    from rocwmma_fattn import sdpa
    sdpa.install()        # monkeypatch torch.nn.functional.scaled_dot_product_attention
    sdpa.uninstall()

Calls the kernels cannot serve (an attention mask, dropout, CPU tensors, dtypes other than fp16 / bf16,
grouped-query broadcasting) raise ``NotImplementedError`` - there is no silent fallback.  A caller who
wants the plugins' behaviour passes the stock implementation explicitly: ``fallback=`` per call, or
``install(fallback_to_torch=True)``; only then are unsupported calls forwarded, unchanged, to it.
"""
from __future__ import annotations

import torch

from .FlashAttn import FlashAttentionFunction

__all__ = ["scaled_dot_product_attention", "supported", "install", "uninstall"]

_torch_sdpa = torch.nn.functional.scaled_dot_product_attention
_installed = False


def supported(query, key, value, attn_mask=None, dropout_p=0.0, enable_gqa=False) -> str | None:
    """None if the call can run on the B200 kernels, else the reason it cannot."""
    if attn_mask is not None:
        return "attn_mask is not supported (the reference ignores its mask argument, FlashAttn.py:49,74)"
    if dropout_p:
        return "dropout is not supported"
    if not (query.is_cuda and key.is_cuda and value.is_cuda):
        return "CUDA tensors only (no CPU fallback)"
    if query.dtype not in (torch.float16, torch.bfloat16) or key.dtype != query.dtype or value.dtype != query.dtype:
        return "fp16 / bf16 only (q, k, v of one dtype)"
    if query.dim() not in (3, 4) or key.dim() != query.dim() or value.dim() != query.dim():
        return "q, k, v must be [B, H, N, D] or [H, N, D]"
    if query.shape[:-2] != key.shape[:-2] or key.shape != value.shape or query.shape[-1] != key.shape[-1]:
        return "q, k, v must share batch, heads and head dim (no grouped-query broadcasting)"
    if enable_gqa:
        return "enable_gqa is not supported"
    return None


def scaled_dot_product_attention(query, key, value, attn_mask=None, dropout_p=0.0, is_causal=False,
                                 scale=None, enable_gqa=False, *, fallback=None):
    """Same positional signature as ``torch.nn.functional.scaled_dot_product_attention``.  Differentiable
    through ``FlashAttentionFunction.backward`` (tcgen05 kernel up to head dim 128, generic CUDA kernel above)."""
    why = supported(query, key, value, attn_mask, dropout_p, enable_gqa)
    if why is not None:
        if fallback is None:
            raise NotImplementedError(f"rocwmma_fattn.sdpa: {why}")
        kw = {"enable_gqa": True} if enable_gqa else {}
        return fallback(query, key, value, attn_mask=attn_mask, dropout_p=dropout_p, is_causal=is_causal,
                        scale=scale, **kw)
    if query.dim() == 3:
        return FlashAttentionFunction.apply(query.unsqueeze(0), key.unsqueeze(0), value.unsqueeze(0), None,
                                            is_causal, scale, False).squeeze(0)
    return FlashAttentionFunction.apply(query, key, value, None, is_causal, scale, False)


def install(fallback_to_torch: bool = False) -> None:
    """Replace ``torch.nn.functional.scaled_dot_product_attention`` (what the ComfyUI / sd-webui plugins
    of the reference do).  With ``fallback_to_torch`` the calls the kernels cannot serve go to the stock
    implementation, otherwise they raise."""
    global _installed
    fb = _torch_sdpa if fallback_to_torch else None

    def patched(query, key, value, attn_mask=None, dropout_p=0.0, is_causal=False, scale=None,
                enable_gqa=False):
        return scaled_dot_product_attention(query, key, value, attn_mask, dropout_p, is_causal, scale,
                                            enable_gqa, fallback=fb)

    torch.nn.functional.scaled_dot_product_attention = patched
    _installed = True


def uninstall() -> None:
    global _installed
    torch.nn.functional.scaled_dot_product_attention = _torch_sdpa
    _installed = False
