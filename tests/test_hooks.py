"""ComfyUI / sd-webui shaped attention hooks (/root/reference/README.md:35-37).

CPU tests cover the host logic (reshapes, strictness, installers) with the one kernel call replaced by a stand-in;
``-m gpu`` tests run the real kernels at Stable-Diffusion shapes against fp32 attention."""
import types

import pytest
import torch

from rocwmma_fattn import hooks


def _ref_attention(q, k, v):  # [B, H, N, D] fp32 attention
    q, k, v = q.float(), k.float(), v.float()
    s = torch.einsum("bhnd,bhmd->bhnm", q, k) * q.shape[-1] ** -0.5
    return torch.einsum("bhnm,bhmd->bhnd", s.softmax(-1), v)


def _split(t, heads):  # [B, N, H*D] -> [B, H, N, D]
    b, n, inner = t.shape
    return t.view(b, n, heads, inner // heads).transpose(1, 2)


@pytest.fixture
def stand_in(monkeypatch):
    """Replace the kernel call (and the CUDA-only check) so the reshapes can be checked without a GPU."""
    calls = []

    def fake(q, k, v, causal, scale, bnhd):
        calls.append((tuple(q.shape), tuple(q.stride()), bnhd))
        if bnhd:
            q, k, v = (t.transpose(1, 2) for t in (q, k, v))
        o = _ref_attention(q, k, v).to(q.dtype)
        return o.transpose(1, 2).contiguous() if bnhd else o

    monkeypatch.setattr(hooks, "_flash", fake)
    monkeypatch.setattr(hooks, "unsupported_reason", lambda q, k, v, mask=None: None if mask is None else "mask")
    return calls


def test_comfy_signature_matches_comfyui():
    import inspect

    names = list(inspect.signature(hooks.comfy_attention).parameters)
    assert names[:8] == ["q", "k", "v", "heads", "mask", "attn_precision", "skip_reshape", "skip_output_reshape"]


def test_comfy_reshapes_use_bnhd_views_without_copies(stand_in):
    torch.manual_seed(0)
    b, nq, nkv, heads, d = 2, 24, 7, 4, 8
    q = torch.rand(b, nq, heads * d)
    k, v = torch.rand(b, nkv, heads * d), torch.rand(b, nkv, heads * d)
    ref = _ref_attention(_split(q, heads), _split(k, heads), _split(v, heads))
    o = hooks.comfy_attention(q, k, v, heads)
    assert o.shape == (b, nq, heads * d)
    torch.testing.assert_close(o, ref.transpose(1, 2).reshape(b, nq, heads * d), atol=1e-5, rtol=1e-5)
    shape, stride, bnhd = stand_in[-1]
    assert bnhd and shape == (b, nq, heads, d) and stride == (nq * heads * d, heads * d, d, 1)  # a view of q itself
    o2 = hooks.comfy_attention(q, k, v, heads, skip_output_reshape=True)
    torch.testing.assert_close(o2, ref, atol=1e-5, rtol=1e-5)
    o3 = hooks.comfy_attention(_split(q, heads), _split(k, heads), _split(v, heads), heads, skip_reshape=True)
    torch.testing.assert_close(o3, ref.transpose(1, 2).reshape(b, nq, heads * d), atol=1e-5, rtol=1e-5)
    assert not stand_in[-1][2]
    o4 = hooks.comfy_attention(_split(q, heads), _split(k, heads), _split(v, heads), heads, skip_reshape=True,
                               skip_output_reshape=True)
    torch.testing.assert_close(o4, ref, atol=1e-5, rtol=1e-5)
    with pytest.raises(ValueError):
        hooks.comfy_attention(q, k, v, 5)


def test_comfy_is_strict_without_a_fallback_and_forwards_with_one():
    q = torch.rand(1, 8, 16)
    with pytest.raises(NotImplementedError, match="CUDA tensors only"):
        hooks.comfy_attention(q, q, q, 2)
    assert "mask" in hooks.unsupported_reason(q, q, q, mask=torch.zeros(8, 8))
    seen = {}

    def fb(q, k, v, heads, mask=None, attn_precision=None, skip_reshape=False, skip_output_reshape=False):
        seen.update(heads=heads, mask=mask, skip_reshape=skip_reshape)
        return "from-fallback"

    assert hooks.comfy_attention(q, q, q, 2, fallback=fb) == "from-fallback"
    assert seen == {"heads": 2, "mask": None, "skip_reshape": False}


def test_install_comfyui_patches_and_restores_a_module(stand_in):
    mod = types.ModuleType("fake_comfy_attention")
    mod.optimized_attention = lambda *a, **kw: "original"
    mod.optimized_attention_masked = mod.optimized_attention
    original = mod.optimized_attention
    assert hooks.install_comfyui(mod) is mod
    assert mod.optimized_attention is not original and mod.optimized_attention_masked is original
    hooks.install_comfyui(mod)  # idempotent
    q = torch.rand(1, 8, 16)
    assert mod.optimized_attention(q, q, q, 2).shape == (1, 8, 16)
    assert mod.optimized_attention(q, q, q, 2, mask=torch.zeros(8, 8)) == "original"  # masked -> the host's own
    hooks.uninstall_comfyui()
    assert mod.optimized_attention is original
    with pytest.raises(ImportError):
        hooks.install_comfyui()  # ComfyUI itself is not installed here


class _CrossAttention(torch.nn.Module):
    """The members sd-webui's hook point relies on (ldm.modules.attention.CrossAttention)."""

    def __init__(self, query_dim, context_dim, heads, dim_head):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.to_q = torch.nn.Linear(query_dim, inner, bias=False)
        self.to_k = torch.nn.Linear(context_dim, inner, bias=False)
        self.to_v = torch.nn.Linear(context_dim, inner, bias=False)
        self.to_out = torch.nn.Sequential(torch.nn.Linear(inner, query_dim), torch.nn.Dropout(0.0))

    def forward(self, x, context=None, mask=None):
        context = x if context is None else context
        q, k, v = (_split(t, self.heads) for t in (self.to_q(x), self.to_k(context), self.to_v(context)))
        o = _ref_attention(q, k, v).to(x.dtype).transpose(1, 2).reshape(x.shape[0], x.shape[1], -1)
        return self.to_out(o)


def test_install_webui_patches_cross_attention(stand_in):
    mod = types.ModuleType("fake_ldm_attention")
    mod.CrossAttention = type("CrossAttention", (_CrossAttention,), {})
    torch.manual_seed(0)
    m, mc = mod.CrossAttention(32, 32, 4, 8), mod.CrossAttention(32, 48, 4, 8)
    x, ctx = torch.rand(2, 10, 32), torch.rand(2, 5, 48)
    ref_self, ref_cross = m(x), mc(x, ctx)
    assert hooks.install_webui(mod) == [mod.CrossAttention]
    assert mod.CrossAttention.forward is hooks.webui_cross_attention_forward
    torch.testing.assert_close(m(x), ref_self, atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(mc(x, context=ctx), ref_cross, atol=1e-5, rtol=1e-5)
    assert stand_in[-1][2]  # went through the BNHD view
    n_calls = len(stand_in)
    torch.testing.assert_close(m(x, mask=torch.zeros(1)), ref_self, atol=1e-5, rtol=1e-5)  # masked -> original forward
    assert len(stand_in) == n_calls
    hooks.uninstall_webui()
    assert mod.CrossAttention.forward is not hooks.webui_cross_attention_forward


# ------------------------------------------------------------------------------------------------
# GPU: the real kernels at Stable-Diffusion shapes (README.md:104-154: SD 1.5 head dims 40 / 80 / 160, SDXL 64; text
# context of 77 tokens)
# ------------------------------------------------------------------------------------------------
SD_SHAPES = [  # batch, Nq, Nkv, heads, dim_head
    (2, 4096, 4096, 8, 40), (2, 4096, 77, 8, 40), (2, 1024, 1024, 8, 80), (2, 256, 77, 8, 160),
    (2, 1024, 1024, 20, 64), (2, 1024, 77, 20, 64), (1, 4096, 4096, 10, 64),
]


@pytest.mark.gpu
@pytest.mark.parametrize("shape", SD_SHAPES, ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16], ids=["f16", "bf16"])
def test_comfy_attention_matches_fp32_at_sd_shapes(shape, dtype):
    b, nq, nkv, heads, d = shape
    torch.manual_seed(0)
    q = torch.randn(b, nq, heads * d, dtype=dtype, device="cuda")
    k = torch.randn(b, nkv, heads * d, dtype=dtype, device="cuda")
    v = torch.randn(b, nkv, heads * d, dtype=dtype, device="cuda")
    ref = _ref_attention(_split(q, heads), _split(k, heads), _split(v, heads))
    o = hooks.comfy_attention(q, k, v, heads)
    assert o.shape == q.shape and o.dtype == dtype and o.is_contiguous()
    tol = 2e-3 if dtype == torch.float16 else 1.6e-2  # randn inputs: |o| up to ~1, one 16-bit rounding of the output + P
    torch.testing.assert_close(o.float(), ref.transpose(1, 2).reshape(b, nq, heads * d), atol=tol, rtol=1e-2)
    o4 = hooks.comfy_attention(_split(q, heads), _split(k, heads), _split(v, heads), heads, skip_reshape=True,
                               skip_output_reshape=True)
    torch.testing.assert_close(o4.float(), ref, atol=tol, rtol=1e-2)


@pytest.mark.gpu
def test_comfy_attention_on_slices_of_a_fused_qkv_projection():
    """q, k, v as chunks of one [B, N, 3 * inner] projection (strided rows): still no copies needed."""
    torch.manual_seed(1)
    b, n, heads, d = 2, 1024, 10, 64
    qkv = torch.randn(b, n, 3 * heads * d, dtype=torch.float16, device="cuda")
    q, k, v = qkv.chunk(3, dim=-1)
    ref = _ref_attention(_split(q.contiguous(), heads), _split(k.contiguous(), heads), _split(v.contiguous(), heads))
    o = hooks.comfy_attention(q, k, v, heads)
    torch.testing.assert_close(o.float(), ref.transpose(1, 2).reshape(b, n, heads * d), atol=2e-3, rtol=1e-2)


@pytest.mark.gpu
def test_webui_cross_attention_forward_and_backward():
    mod = types.ModuleType("fake_ldm_attention")
    mod.CrossAttention = type("CrossAttention", (_CrossAttention,), {})
    torch.manual_seed(0)
    m = mod.CrossAttention(320, 768, 8, 40).cuda().half()
    x = torch.randn(2, 1024, 320, dtype=torch.float16, device="cuda", requires_grad=True)
    ctx = torch.randn(2, 77, 768, dtype=torch.float16, device="cuda")
    ref = m(x, ctx)
    g_ref, = torch.autograd.grad(ref.float().square().sum(), x)
    hooks.install_webui(mod)
    try:
        out = m(x, context=ctx)
        g, = torch.autograd.grad(out.float().square().sum(), x)
    finally:
        hooks.uninstall_webui()
    torch.testing.assert_close(out.float(), ref.float(), atol=4e-3, rtol=2e-2)
    torch.testing.assert_close(g.float(), g_ref.float(), atol=2e-2 * g_ref.float().abs().max().item(), rtol=5e-2)
