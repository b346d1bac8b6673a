"""CPU oracle for the attention-forward hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module; the product path (``flash-attention-v2-rdna3-minimal_b200/``) never
does and has no CPU fallback.

It restates, on the CPU, the algorithm of the reference's own oracle and checkers:

* ``tiled_fa2_forward``  <- /root/reference/pure_torch_ver.py:24-90   (tiled FA2 forward, in the
                            input dtype, natural exp, -100 padding, -65500 causal fill)
* ``sdpa_math``          <- /root/reference/pure_torch_ver.py:181, precision_test.py:65
                            (math-backend ``F.scaled_dot_product_attention``: softmax(scale q k^T) v)
                            evaluated in fp32 / fp64 as the ground truth
* ``cpu_sdpa``           <- the same call in the input dtype on CPU tensors: the reference's "CPU SDPA
                            path" that bench.py times as ``cpu_baseline``
* ``lse_base2``          <- /root/reference/rocwmma_fattn/kernel_fp16.cu:541-542 (L = m + log2 l over
                            scores pre-multiplied by scale*log2(e), :827)
* ``tiled_fa2_backward`` <- /root/reference/pure_torch_ver.py:94-153  (tiled FA2 backward)
* ``sdpa_backward``      <- autograd through math SDPA, the reference's gradient check
                            (pure_torch_ver.py:192-205, precision_test.py:75-98)

Pinning: the reference ships no golden vectors, tolerances or seeds, so the
oracle is pinned against outputs of the reference's own ``pure_torch_ver.py`` run in the build
container: ``tests/golden/make_golden.py`` imports it from /root/reference and stores inputs and
outputs under ``tests/golden/*.npz``; ``tests/test_oracle.py`` checks this module against them.

Third-party arithmetic: everything below bottoms out in PyTorch CPU kernels (torch 2.11 here; the
reference README states PyTorch 2.2.1, README.md:33), as the reference's oracle does.
"""
from __future__ import annotations

import math

import torch

LOG2E = 1.4426950408889634

# tolerances (max-abs, U[0,1) inputs) adopted in SURVEY.md section 8c: one output rounding of the
# 16-bit type plus the reference oracle's own distance from fp32 SDPA
TOL_VS_FP32 = {torch.float16: 1e-3, torch.bfloat16: 8e-3}
# randn inputs: outputs are O(1) with larger dynamic range -> relative companion
RTOL_RANDN = {torch.float16: 1e-2, torch.bfloat16: 1e-2}
ATOL_RANDN = {torch.float16: 2e-3, torch.bfloat16: 1.6e-2}


def _causal_keep_mask(nq: int, nkv: int, device=None) -> torch.Tensor:
    """True where the score is kept.  Top-left aligned: key j visible to query i iff j <= i
    (reference masks ``col > row``, kernel_fp16.cu:403-411; identical to SDPA is_causal=True)."""
    i = torch.arange(nq, device=device).unsqueeze(1)
    j = torch.arange(nkv, device=device).unsqueeze(0)
    return j <= i


def sdpa_math(q, k, v, causal=False, scale=None, dtype=torch.float32):
    """Ground truth: softmax(scale * q k^T [+ causal mask]) v evaluated in ``dtype`` (fp32 or fp64)
    on [B,H,N,D] tensors.  Returns (o, lse2) with lse2 the base-2 log-sum-exp of the scaled scores
    (see ``lse_base2``)."""
    qf, kf, vf = q.to(dtype), k.to(dtype), v.to(dtype)
    d = q.shape[-1]
    if scale is None:
        scale = d ** -0.5
    s = torch.matmul(qf, kf.transpose(-1, -2)) * scale
    if causal:
        keep = _causal_keep_mask(q.shape[-2], k.shape[-2], device=q.device)
        s = s.masked_fill(~keep, float("-inf"))
    m = s.amax(dim=-1, keepdim=True)
    p = torch.exp(s - m)
    l = p.sum(dim=-1, keepdim=True)
    o = torch.matmul(p / l, vf)
    lse2 = (m.squeeze(-1) + torch.log(l.squeeze(-1))) * LOG2E
    return o, lse2


def lse_base2(q, k, causal=False, scale=None, dtype=torch.float64):
    """L as the reference's kernels store it: with s' = scale*log2(e)*q.k, L = max s' + log2 sum
    2^(s' - max) (kernel_fp16.cu:827 for the folded scale, :541-542 for L)."""
    d = q.shape[-1]
    if scale is None:
        scale = d ** -0.5
    s = torch.matmul(q.to(dtype), k.to(dtype).transpose(-1, -2)) * (scale * LOG2E)
    if causal:
        keep = _causal_keep_mask(q.shape[-2], k.shape[-2], device=q.device)
        s = s.masked_fill(~keep, float("-inf"))
    m = s.amax(dim=-1)
    return m + torch.log2(torch.exp2(s - m.unsqueeze(-1)).sum(dim=-1))


def cpu_sdpa(q, k, v, causal=False, scale=None):
    """The reference's CPU SDPA path: ``F.scaled_dot_product_attention`` in the input dtype
    (pure_torch_ver.py:181; on CPU tensors this is the fused CPU kernel / math path)."""
    return torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=causal, scale=scale)


def _pad_rows(t: torch.Tensor, multiple: int, value: float) -> torch.Tensor:
    """Append rows of ``value`` along dim -2 up to a multiple (pure_torch_ver.py:9-18)."""
    n = t.shape[-2]
    extra = (-n) % multiple
    if extra == 0:
        return t
    fill = torch.full((*t.shape[:-2], extra, t.shape[-1]), value, dtype=t.dtype, device=t.device)
    return torch.cat((t, fill), dim=-2)


def tiled_fa2_forward(q, k, v, causal=False, Br=64, Bc=256):
    """Restatement of the reference oracle's forward (pure_torch_ver.py:24-90) on [B,H,N,D]
    tensors, all arithmetic in the input dtype:

      pad Q, K rows with -100 and V rows with 0 (:34-36; only a valid mask for Q >= 0 inputs);
      for each Br-row Q tile, walk the Bc-row K/V tiles keeping (m, l, O):
        S = (scale Q_i) K_j^T                                   (:60-62)
        causal: fill -65500 above the diagonal of tiles that cross it   (:64-69)
        m' = max(m, rowmax S); P = exp(S - m'); l = l exp(m - m') + rowsum P;
        O = O exp(m - m') + P V_j                                (:71-79)
      O /= l; L = m + log l                                      (:81-85)

    Returns (o[B,H,N,D], L[B,H,N] fp32 natural-log LSE)."""
    B, H, N, D = q.shape
    scale = D ** -0.5
    dt = q.dtype
    qp = _pad_rows(q, Br, -100.0)
    kp = _pad_rows(k, Bc, -100.0)
    vp = _pad_rows(v, Bc, 0.0)
    Np, Nkp = qp.shape[-2], kp.shape[-2]
    o = torch.zeros_like(qp)
    L = torch.zeros((B, H, Np), dtype=torch.float32)
    neg_inf = float("-inf")

    for r0 in range(0, Np, Br):
        qi = qp[:, :, r0:r0 + Br, :] * scale
        m = torch.full((B, H, Br), neg_inf, dtype=dt)
        l = torch.zeros((B, H, Br), dtype=dt)
        acc = torch.zeros((B, H, Br, D), dtype=dt)
        for c0 in range(0, Nkp, Bc):
            kj = kp[:, :, c0:c0 + Bc, :]
            vj = vp[:, :, c0:c0 + Bc, :]
            s = torch.matmul(qi, kj.transpose(-1, -2))
            if causal and r0 < c0 + Bc - 1:
                rows = torch.arange(r0, r0 + Br).unsqueeze(1)
                cols = torch.arange(c0, c0 + Bc).unsqueeze(0)
                s = s.masked_fill(cols > rows, -65500.0)
            m_new = torch.maximum(s.amax(dim=-1), m)
            p = torch.exp(s - m_new.unsqueeze(-1))
            shrink = torch.exp(m - m_new)
            l = l * shrink + p.sum(dim=-1)
            acc = acc * shrink.unsqueeze(-1) + torch.matmul(p, vj)
            m = m_new
        o[:, :, r0:r0 + Br, :] = acc / l.unsqueeze(-1)
        L[:, :, r0:r0 + Br] = (m + torch.log(l)).to(torch.float32)
    return o[:, :, :N, :], L[:, :, :N]


def sdpa_backward(q, k, v, d_o, causal=False, scale=None, dtype=torch.float32):
    """Ground truth for the backward: autograd through the math definition
    softmax(scale q k^T [+ causal mask]) v evaluated in ``dtype`` (what the reference's
    ``o2.backward(dO)`` on math-SDPA gives, pure_torch_ver.py:192-205 / precision_test.py:75-98).
    Returns (dq, dk, dv) in ``dtype``."""
    qf = q.detach().to(dtype).requires_grad_(True)
    kf = k.detach().to(dtype).requires_grad_(True)
    vf = v.detach().to(dtype).requires_grad_(True)
    d = q.shape[-1]
    if scale is None:
        scale = d ** -0.5
    s = torch.matmul(qf, kf.transpose(-1, -2)) * scale
    if causal:
        keep = _causal_keep_mask(q.shape[-2], k.shape[-2], device=q.device)
        s = s.masked_fill(~keep, float("-inf"))
    o = torch.matmul(torch.softmax(s, dim=-1), vf)
    o.backward(d_o.to(dtype))
    return qf.grad, kf.grad, vf.grad


def tiled_fa2_backward(q, k, v, o, L, d_o, causal=False, Br=64, Bc=256):
    """Restatement of the reference oracle's backward (pure_torch_ver.py:94-153) for sequence
    lengths that are multiples of the tile sizes, arithmetic in the input dtype:

      for each Bc-row K/V tile j, for each Br-row Q tile i:
        S = scale Q_i K_j^T, causal fill -inf above the diagonal            (:137-144)
        P = exp(S - L_i) rounded to the input dtype                          (:146)
        dV_j += P^T dO_i;  dP = dO_i V_j^T;  D_i = rowsum(dO_i o O_i)        (:147-149)
        dS = scale P (dP - D_i);  dQ_i += dS K_j;  dK_j += dS^T Q_i          (:150-152)

    ``L`` is the natural-log LSE the oracle forward returns.  Returns (dq, dk, dv)."""
    B, H, N, D = q.shape
    Nkv = k.shape[-2]
    if N % Br or Nkv % Bc:
        raise ValueError("tiled_fa2_backward restates the aligned case only")
    scale = D ** -0.5
    dq, dk, dv = torch.zeros_like(q), torch.zeros_like(k), torch.zeros_like(v)
    for c0 in range(0, Nkv, Bc):
        kj, vj = k[:, :, c0:c0 + Bc, :], v[:, :, c0:c0 + Bc, :]
        for r0 in range(0, N, Br):
            qi, oi, doi = q[:, :, r0:r0 + Br, :], o[:, :, r0:r0 + Br, :], d_o[:, :, r0:r0 + Br, :]
            li = L[:, :, r0:r0 + Br]
            s = scale * torch.matmul(qi, kj.transpose(-1, -2))
            if causal and r0 < c0 + Bc - 1:
                rows = torch.arange(r0, r0 + Br).unsqueeze(1)
                cols = torch.arange(c0, c0 + Bc).unsqueeze(0)
                s = s.masked_fill(cols > rows, float("-inf"))
            p = torch.exp(s - li.unsqueeze(-1)).to(q.dtype)
            dv[:, :, c0:c0 + Bc, :] += torch.matmul(p.transpose(-1, -2), doi)
            dp = torch.matmul(doi, vj.transpose(-1, -2))
            di = torch.sum(doi * oi, -1)
            ds = scale * p * (dp - di.unsqueeze(-1))
            dq[:, :, r0:r0 + Br, :] += torch.matmul(ds, kj)
            dk[:, :, c0:c0 + Bc, :] += torch.matmul(ds.transpose(-1, -2), qi)
    return dq, dk, dv


# Backward gate.  Ground truth = fp32 autograd of math SDPA (``sdpa_backward``).  Every Flash-
# Attention backward, the reference's included, takes D_i = rowsum(dO o O) from the 16-bit O the
# forward stored (kernel_fp16.cu:605-631), so part of its distance from fp32 autograd is inherited
# from that rounding, not from the backward arithmetic (dQ of uniform data is a heavily cancelling
# sum: 1 % of max|dQ| in fp16, 13 % in bf16 at N = 512 for ANY implementation).  The gate is the
# reference's own procedure made quantitative: precision_test.py:65-98 compares the extension with
# math SDPA run in the SAME 16-bit dtype; we require
#     max|g - ref32| <= 2 * max|g16 - ref32| + 1e-5 + 1e-3 * max|ref32|
# where g16 = autograd of math SDPA evaluated in the 16-bit input dtype (``sdpa_backward(dtype=...)``),
# i.e. at most twice the error of the plain 16-bit PyTorch implementation (the criterion the public
# flash-attention test-suite uses), and the tests on the committed fixtures also require the
# kernels to be at least as close to fp32 as the reference's tiled oracle is.
def check_close_grad(g, ref_fp32, g16) -> tuple[bool, float, float]:
    """Returns (ok, err, bound) for one gradient tensor; ``g16`` is the 16-bit baseline gradient."""
    r64 = ref_fp32.double().cpu()
    err = (g.double().cpu() - r64).abs().max().item()
    base = (g16.double().cpu() - r64).abs().max().item()
    bound = 2.0 * base + 1e-5 + 1e-3 * r64.abs().max().item()
    return err <= bound, err, bound


def attention_flops(B, H, Nq, Nkv, D, causal=False) -> float:
    """The reference's FLOP count: 2 matmuls x 2 B H N^2 D, halved when causal
    (bench_with_sdpa.py:35-38)."""
    f = 4.0 * B * H * Nq * Nkv * D
    return f * 0.5 if causal else f


def attention_bytes(B, H, Nq, Nkv, D, elem_bytes=2, with_lse=False) -> float:
    """Algorithmic HBM bytes: read Q, K, V once, write O once."""
    b = elem_bytes * B * H * D * (2 * Nq + 2 * Nkv)
    return b + (4 * B * H * Nq if with_lse else 0)


def make_inputs(B, H, Nq, Nkv, D, dtype, seed=0, dist="rand", device="cpu", bnhd=False):
    """Seeded synthetic Q, K, V.  ``rand`` is U[0,1) like every script of the reference
    (bench_with_sdpa.py:207-209, precision_test.py:44-46); ``randn`` exercises signed scores."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    fn = torch.rand if dist == "rand" else torch.randn

    def one(n):
        shape = (B, n, H, D) if bnhd else (B, H, n, D)
        return fn(shape, generator=g, dtype=torch.float32).to(dtype).to(device)

    return one(Nq), one(Nkv), one(Nkv)


def max_abs_err(a: torch.Tensor, b: torch.Tensor) -> float:
    return (a.double().cpu() - b.double().cpu()).abs().max().item()


def check_close(o, ref_fp32, dtype, dist="rand") -> tuple[bool, float]:
    """Apply the stated tolerance: max-abs for U[0,1) inputs; atol+rtol for randn."""
    o64, r64 = o.double().cpu(), ref_fp32.double().cpu()
    err = (o64 - r64).abs()
    if dist == "rand":
        return bool(err.max().item() <= TOL_VS_FP32[dtype]), err.max().item()
    bound = ATOL_RANDN[dtype] + RTOL_RANDN[dtype] * r64.abs()
    return bool((err <= bound).all().item()), err.max().item()


def nan_free(t: torch.Tensor) -> bool:
    return bool(torch.isfinite(t.float()).all().item())


__all__ = [
    "sdpa_math", "lse_base2", "cpu_sdpa", "tiled_fa2_forward", "attention_flops",
    "sdpa_backward", "tiled_fa2_backward", "check_close_grad",
    "attention_bytes", "make_inputs", "max_abs_err", "check_close", "nan_free",
    "TOL_VS_FP32", "LOG2E",
]

if __name__ == "__main__":  # tiny self-demo
    q, k, v = make_inputs(1, 2, 128, 128, 64, torch.float16)
    o_ref, _ = sdpa_math(q, k, v)
    o_t, _ = tiled_fa2_forward(q, k, v)
    print("tiled vs fp32 sdpa:", max_abs_err(o_t, o_ref), " log2e", math.log2(math.e))
