"""Attention hooks in the two shapes the reference's plugins use.

The reference reaches Stable Diffusion through two external plugins (/root/reference/README.md:35-37):

* **ComfyUI** replaces ``comfy.ldm.modules.attention.optimized_attention``, a function of
  ``(q, k, v, heads, mask=None, attn_precision=None, skip_reshape=False, skip_output_reshape=False)`` on
  ``[B, N, heads * dim_head]`` tensors.  Viewed as ``[B, N, heads, dim_head]`` that is exactly the layout the
  reference's ``BNHD_fmt=True`` flag exists for (FlashAttn.py:49,58-60; bench_with_sdpa_BNHD.py:106), so the hook
  makes no transposed copies: the kernels read the packed activations through strides and write the output
  in the same layout.
* **sd-webui** replaces ``CrossAttention.forward`` of ``ldm`` / ``sgm``: ``to_q / to_k / to_v`` projections,
  attention over ``heads``, ``to_out``.

Both are provided here as plain functions plus ``install_*`` helpers that patch the host application's module when
it is importable (or a module object handed in, which is how the tests drive them).  Calls the kernels cannot
serve (an attention mask, fp32 activations the caller did not ask to down-cast, CPU tensors) go to the
``fallback`` the caller names - the host application's own attention function, as the reference's plugins do -
or raise ``NotImplementedError`` when none was given.  There is no silent fallback inside this package.
"""
from __future__ import annotations

import importlib
import types

import torch

from .FlashAttn import FlashAttentionFunction

__all__ = ["comfy_attention", "webui_cross_attention_forward", "install_comfyui", "uninstall_comfyui",
           "install_webui", "uninstall_webui", "unsupported_reason"]


def _flash(q, k, v, causal, scale, bnhd):
    """The one call into the kernels (a seam the CPU tests replace with a stand-in)."""
    return FlashAttentionFunction.apply(q, k, v, None, causal, scale, bnhd)


def unsupported_reason(q, k, v, mask=None) -> str | None:
    """None if the B200 kernels can serve the call, else why not."""
    if mask is not None:
        return "an attention mask / bias is not supported (the reference ignores its mask argument, FlashAttn.py:49,74)"
    if not (q.is_cuda and k.is_cuda and v.is_cuda):
        return "CUDA tensors only (no CPU path)"
    if q.dtype not in (torch.float16, torch.bfloat16) or k.dtype != q.dtype or v.dtype != q.dtype:
        return "fp16 / bf16 activations only (q, k, v of one dtype)"
    return None


def comfy_attention(q, k, v, heads, mask=None, attn_precision=None, skip_reshape=False,
                    skip_output_reshape=False, *, fallback=None, **kwargs):
    """ComfyUI's ``optimized_attention`` signature.

    ``q``: ``[B, Nq, heads * dim_head]``, ``k`` / ``v``: ``[B, Nkv, heads * dim_head]`` (or all three already
    ``[B, heads, N, dim_head]`` with ``skip_reshape``); returns ``[B, Nq, heads * dim_head]`` (or
    ``[B, heads, Nq, dim_head]`` with ``skip_output_reshape``).  ``attn_precision`` (fp32 up-cast requests) is
    accepted and ignored: the kernels accumulate in fp32 whatever the input type."""
    why = unsupported_reason(q, k, v, mask)
    if why is not None:
        if fallback is None:
            raise NotImplementedError(f"rocwmma_fattn.hooks.comfy_attention: {why}")
        return fallback(q, k, v, heads, mask=mask, attn_precision=attn_precision, skip_reshape=skip_reshape,
                        skip_output_reshape=skip_output_reshape, **kwargs)
    if skip_reshape:
        if q.dim() != 4 or q.shape[1] != heads:
            raise ValueError("skip_reshape expects [B, heads, N, dim_head]")
        b, _, nq, dim_head = q.shape
        o = _flash(q, k, v, False, None, False)                       # [B, heads, Nq, dim_head]
        return o if skip_output_reshape else o.transpose(1, 2).reshape(b, nq, heads * dim_head)
    if q.dim() != 3 or q.shape[-1] % heads != 0:
        raise ValueError("expected [B, N, heads * dim_head]")
    b, nq, inner = q.shape
    dim_head = inner // heads
    # [B, N, heads, dim_head] views of the packed activations: the BNHD layout, no copies
    o = _flash(q.unflatten(-1, (heads, dim_head)), k.unflatten(-1, (heads, dim_head)),
               v.unflatten(-1, (heads, dim_head)), False, None, True)       # [B, Nq, heads, dim_head]
    return o.transpose(1, 2) if skip_output_reshape else o.reshape(b, nq, inner)


def webui_cross_attention_forward(self, x, context=None, mask=None, **kwargs):
    """``forward`` for the ``CrossAttention`` module of ``ldm`` / ``sgm`` (sd-webui's optimisation hook point):
    ``self.to_q / to_k / to_v`` projections, attention over ``self.heads`` heads, ``self.to_out``.  The module's
    own scale (``dim_head ** -0.5``) is the kernels' default."""
    h = self.heads
    context = x if context is None else context
    q = self.to_q(x)
    k = self.to_k(context)
    v = self.to_v(context)
    why = unsupported_reason(q, k, v, mask)
    if why is not None:
        orig = getattr(type(self), "_rocwmma_fattn_orig_forward", None)
        if orig is None:
            raise NotImplementedError(f"rocwmma_fattn.hooks.webui_cross_attention_forward: {why}")
        return orig(self, x, context=context, mask=mask, **kwargs)
    out = comfy_attention(q, k, v, h)
    return self.to_out(out)


# ------------------------------------------------------------------------------------------------
# installers
# ------------------------------------------------------------------------------------------------
_COMFY_MODULE = "comfy.ldm.modules.attention"
_WEBUI_MODULES = ("ldm.modules.attention", "sgm.modules.attention")
_comfy_saved: dict = {}
_webui_saved: list = []


def _resolve(module, names):
    if isinstance(module, types.ModuleType) or (module is not None and not isinstance(module, str)):
        return [module]
    found = []
    for name in ([module] if module else names):
        try:
            found.append(importlib.import_module(name))
        except ImportError:
            pass
    if not found:
        raise ImportError("none of %s is importable; pass the host application's attention module" % (names,))
    return found


def install_comfyui(module=None, fallback_to_original: bool = True):
    """Point ``optimized_attention`` of ComfyUI's attention module at the B200 kernels (what the reference's ComfyUI
    plugin does for the rocWMMA kernels).  Masked calls keep going to ``optimized_attention_masked``, which is
    left alone; with ``fallback_to_original`` whatever else the kernels cannot serve goes to the function that was
    installed before.  Returns the patched module."""
    mod = _resolve(module, (_COMFY_MODULE,))[0]
    orig = getattr(mod, "optimized_attention")
    if getattr(orig, "_rocwmma_fattn_hook", False):
        return mod
    fb = orig if fallback_to_original else None

    def optimized_attention(q, k, v, heads, mask=None, attn_precision=None, skip_reshape=False,
                            skip_output_reshape=False, **kwargs):
        return comfy_attention(q, k, v, heads, mask, attn_precision, skip_reshape, skip_output_reshape,
                               fallback=fb, **kwargs)

    optimized_attention._rocwmma_fattn_hook = True
    _comfy_saved[id(mod)] = (mod, orig)
    mod.optimized_attention = optimized_attention
    return mod


def uninstall_comfyui():
    for mod, orig in _comfy_saved.values():
        mod.optimized_attention = orig
    _comfy_saved.clear()


def install_webui(module=None):
    """Replace ``CrossAttention.forward`` in ``ldm.modules.attention`` / ``sgm.modules.attention`` (sd-webui's
    hook point); the original forward stays reachable for the calls the kernels cannot serve.  Returns the list
    of patched classes."""
    patched = []
    for mod in _resolve(module, _WEBUI_MODULES):
        cls = getattr(mod, "CrossAttention")
        if getattr(cls, "_rocwmma_fattn_orig_forward", None) is None:
            cls._rocwmma_fattn_orig_forward = cls.forward
            cls.forward = webui_cross_attention_forward
            _webui_saved.append(cls)
        patched.append(cls)
    return patched


def uninstall_webui():
    for cls in _webui_saved:
        cls.forward = cls._rocwmma_fattn_orig_forward
        cls._rocwmma_fattn_orig_forward = None
    _webui_saved.clear()
