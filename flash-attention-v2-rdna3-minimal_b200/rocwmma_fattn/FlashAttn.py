"""Drop-in for /root/reference/rocwmma_fattn/FlashAttn.py (forward and backward).

Same entry point and calling convention as the reference::

    from rocwmma_fattn.FlashAttn import FlashAttentionFunction
    o = FlashAttentionFunction.apply(q, k, v, mask, causal, scale, BNHD_fmt)

* ``q`` is ``[B,H,N,D]`` (or ``[B,N,H,D]`` with ``BNHD_fmt=True``), ``k``/``v`` likewise with ``Nkv``;
  the result has the shape, dtype and memory layout of ``q`` (reference: FlashAttn.py:49-76).
* ``mask`` is accepted and ignored, exactly as in the reference (FlashAttn.py:49,74).
* ``causal=None`` means False; the mask is top-left aligned (``col > row`` masked,
  kernel_fp16.cu:396-412), i.e. ``F.scaled_dot_product_attention(is_causal=True)``.
* ``scale`` defaults to ``D ** -0.5`` (FlashAttn.py:63-64).
* inputs that are neither fp16 nor bf16 are computed in bf16 (host.cpp:41-44).

Underneath, instead of the JIT-hipified rocWMMA extension, this calls the C-ABI library
``lib/libfa_fwd_sm100.so`` (include/fa_fwd_sm100.h) through ctypes with raw device pointers and
strides.  Tile sizes (the reference's Br/Bc, FlashAttn.py:56-67) are internal to the kernels.  There
is no CPU fallback: CPU tensors raise, and a missing library fails the import.
"""
from __future__ import annotations

import torch

from . import _capi

__all__ = [
    "FlashAttentionFunction",
    "flash_attn_forward",
    "flash_attn_forward_host",
    "flash_attn_host_sync",
    "flash_attn_wmma",
]

_TMA_D_ALIGN = 8  # head dim multiple of 16 bytes for the tensor-core (TMA) kernels
_TC_MAX_D = 256      # forward: ws / sk / tc1 kernels up to 128, the wide kernel up to 256
_TC_MAX_D_BWD = 256  # tcgen05 backward kernels (fa_bwd_ws up to 128, fa_bwd_wide up to 256); larger head dims run the generic CUDA-core backward


def _raw_stream(dev_index: int) -> int:
    """cudaStream_t of torch's current stream on that device (the launch stream)."""
    try:
        return torch._C._cuda_getCurrentRawStream(dev_index)
    except AttributeError:  # private API moved: the public spelling is just slower
        return torch.cuda.current_stream(dev_index).cuda_stream


def _dtype_code(dt: torch.dtype) -> int:
    if dt == torch.float16:
        return _capi.FA_DTYPE_F16
    if dt == torch.bfloat16:
        return _capi.FA_DTYPE_BF16
    raise TypeError(f"unsupported dtype {dt}")


def _logical_strides(t: torch.Tensor, bnhd: bool):
    """Element strides in logical (b, h, n, d) order; the identity the reference's kernels apply
    for ``permute_NH`` (kernel_fp16.cu:324-333, checked by test_arrange.py:23-30)."""
    s = t.stride()
    return (s[0], s[2], s[1], s[3]) if bnhd else (s[0], s[1], s[2], s[3])


def _logical_shape(t: torch.Tensor, bnhd: bool):
    sh = t.shape
    return (sh[0], sh[2], sh[1], sh[3]) if bnhd else (sh[0], sh[1], sh[2], sh[3])


def _prepare(t: torch.Tensor, d_pad: int) -> torch.Tensor:
    if d_pad:
        t = torch.nn.functional.pad(t, (0, d_pad))  # reference pads D too (kernel_fp16.cu:763-779)
    if t.stride(-1) != 1:
        t = t.contiguous()  # kernel_fp16.cu:780-787
    return t


def _tma_view(t: torch.Tensor) -> torch.Tensor:
    """`t` itself if TMA can address it (16-byte aligned base, strides multiples of 8 elements), else a
    contiguous copy."""
    if t.data_ptr() % 16 == 0 and all(st % 8 == 0 for st, n in zip(t.stride()[:3], t.shape[:3]) if n > 1):
        return t
    return t.contiguous()


def _forward(q, k, v, causal, scale, bnhd, want_lse):
    if q.dim() != 4 or k.dim() != 4 or v.dim() != 4:
        raise ValueError("q, k, v must be 4-D: [B,H,N,D] or [B,N,H,D] (BNHD_fmt=True)")
    if not q.is_cuda:
        raise RuntimeError(
            "rocwmma_fattn (B200 build) runs on CUDA tensors only; there is no CPU fallback"
        )
    if k.device != q.device or v.device != q.device:
        raise ValueError("q, k, v must be on the same device")
    if q.dtype not in (torch.float16, torch.bfloat16):
        # reference: host.cpp:41-44 casts anything else to bf16 and returns bf16
        q, k, v = q.to(torch.bfloat16), k.to(torch.bfloat16), v.to(torch.bfloat16)
    if k.dtype != q.dtype or v.dtype != q.dtype:
        raise TypeError("q, k, v must share one dtype")

    B, H, Nq, D = _logical_shape(q, bnhd)
    Bk, Hk, Nkv, Dk = _logical_shape(k, bnhd)
    if (Bk, Hk, Dk) != (B, H, D) or _logical_shape(v, bnhd) != (Bk, Hk, Nkv, Dk):
        raise ValueError(
            f"shape mismatch: q {tuple(q.shape)}, k {tuple(k.shape)}, v {tuple(v.shape)}"
        )
    if min(B, H, Nq, Nkv, D) < 1:
        raise ValueError("empty tensors are not supported (every dimension must be >= 1)")
    if scale is None:
        scale = D ** -0.5
    causal = bool(causal)

    # head dims that are not a multiple of 8 cannot be addressed by TMA: zero-pad them (zeros change
    # neither q.k nor the first D columns of p.v); D > 256 goes to the generic kernel unpadded.
    d_pad = (-D) % _TMA_D_ALIGN if D < _TC_MAX_D else 0
    qp, kp, vp = _prepare(q, d_pad), _prepare(k, d_pad), _prepare(v, d_pad)
    o_full = torch.empty_like(qp)
    lse = torch.empty((B, H, Nq), dtype=torch.float32, device=q.device) if want_lse else None

    dev_index = q.device.index
    args = (
        qp.data_ptr(), kp.data_ptr(), vp.data_ptr(), o_full.data_ptr(),
        lse.data_ptr() if lse is not None else None,
        B, H, Nq, Nkv, D + d_pad,
        _capi.strides4(_logical_strides(qp, bnhd)), _capi.strides4(_logical_strides(kp, bnhd)),
        _capi.strides4(_logical_strides(vp, bnhd)), _capi.strides4(_logical_strides(o_full, bnhd)),
        _dtype_code(qp.dtype), int(causal), float(scale),
    )
    # the call costs a few microseconds of Python on top of the launch; at N <= 1024 that is as long as
    # the kernel, so skip the device guard when q already lives on the current device
    if torch.cuda.current_device() == dev_index:
        rc = _capi.lib.fa_fwd_sm100(*args, _raw_stream(dev_index))
    else:
        with torch.cuda.device(dev_index):
            rc = _capi.lib.fa_fwd_sm100(*args, _raw_stream(dev_index))
    if rc:
        _capi.check(rc, "fa_fwd_sm100")
    o = o_full[..., :D] if d_pad else o_full
    return o, lse, (qp, kp, vp, o_full), (causal, float(scale), Nq, Nkv, D, bnhd)


def _backward(qp, kp, vp, o_full, d_o, lse, D, causal, scale, bnhd):
    """dQ, dK, dV on the tensors the forward saved (head dim padded to a multiple of 8).  Replaces
    backward_fp16 / backward_bf16 (kernel_fp16.cu:878-1028): the incoming gradient is zero-padded
    in the head dim like there (:903-917), the three gradients come back sliced to ``D``.  Head dims up to
    128 run the pipelined tcgen05 kernel (csrc/fa_bwd_ws.cuh), 129..256 the three-launch tcgen05 kernel
    (csrc/fa_bwd_wide.cuh), larger ones (the reference pads and serves any, :900) the generic CUDA-core
    backward (csrc/fa_bwd_simt.cuh)."""
    if not qp.is_cuda:
        raise RuntimeError("rocwmma_fattn (B200 build) runs on CUDA tensors only; there is no CPU fallback")
    DP = qp.shape[3]
    if d_o.dtype != qp.dtype:
        d_o = d_o.to(qp.dtype)  # host.cpp:49-57 dispatches on dO's dtype; ours follows q's
    d_o = _prepare(d_o, DP - d_o.shape[3])
    if DP <= _TC_MAX_D_BWD and DP % _TMA_D_ALIGN == 0:
        # The tcgen05 backward reads its operands through TMA: 16-byte aligned base pointers and strides.  The
        # forward serves other views (an odd offset into a packed qkv buffer, a row stride that is not a multiple
        # of 8) with its generic kernel; the backward makes them contiguous instead, as the reference does for
        # everything (kernel_fp16.cu:903-917), rather than dropping to the generic backward.
        qp, kp, vp, o_full, d_o = (_tma_view(t) for t in (qp, kp, vp, o_full, d_o))
    B, H, Nq, _ = _logical_shape(qp, bnhd)
    Nkv = _logical_shape(kp, bnhd)[2]
    dq, dk, dv = torch.empty_like(qp), torch.empty_like(kp), torch.empty_like(vp)
    # fp32 dQ accumulator of the head-dim <= 128 kernel (the other kernels write dQ directly: a token buffer)
    dq_acc = torch.empty((B, H, Nq, DP) if DP <= 128 else (8,), dtype=torch.float32, device=qp.device)
    delta = torch.empty((B, H, Nq), dtype=torch.float32, device=qp.device)
    st = lambda t: _capi.strides4(_logical_strides(t, bnhd))  # noqa: E731
    with torch.cuda.device(qp.device):
        stream = torch.cuda.current_stream(qp.device).cuda_stream
        rc = _capi.lib.fa_bwd_sm100(
            qp.data_ptr(), kp.data_ptr(), vp.data_ptr(), o_full.data_ptr(), d_o.data_ptr(),
            lse.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), dq_acc.data_ptr(),
            delta.data_ptr(), B, H, Nq, Nkv, DP,
            st(qp), st(kp), st(vp), st(o_full), st(d_o), st(dq), st(dk), st(dv),
            _dtype_code(qp.dtype), int(causal), float(scale), stream,
        )
    _capi.check(rc, "fa_bwd_sm100")
    if DP != D:
        dq, dk, dv = dq[..., :D], dk[..., :D], dv[..., :D]
    return dq, dk, dv


def flash_attn_forward(q, k, v, causal=False, scale=None, BNHD_fmt=False, return_lse=False):
    """Functional form of the forward.  With ``return_lse`` also returns the base-2 log-sum-exp of
    the scaled scores, fp32 ``[B,H,Nq]`` (what the reference stores in ``L``,
    kernel_fp16.cu:541-542)."""
    o, lse, _, _ = _forward(q, k, v, causal, scale, BNHD_fmt, return_lse)
    return (o, lse) if return_lse else o


def flash_attn_forward_host(q, k, v, causal=False, scale=None, out=None, return_lse=False, wait=True):
    """Forward on HOST tensors (contiguous ``[B,H,N,D]``, ideally pinned): the C-ABI stages the
    inputs to the current CUDA device, runs the same kernels and copies the result back, overlapping
    the copies with compute.  Returns a host tensor (``out`` if given).  This is the end-to-end
    path ``bench.py`` reports as ``e2e``.

    ``wait=False`` returns as soon as the work is enqueued (``fa_fwd_sm100_host_async``): the inputs and
    the returned tensor must not be touched until ``flash_attn_host_sync()``; consecutive calls then
    pipeline (the uploads of one call run under the tail of the previous one)."""
    for t in (q, k, v):
        if t.is_cuda or t.dim() != 4 or not t.is_contiguous():
            raise ValueError("flash_attn_forward_host expects contiguous 4-D host tensors")
    if q.dtype not in (torch.float16, torch.bfloat16) or k.dtype != q.dtype or v.dtype != q.dtype:
        raise TypeError("q, k, v must be fp16 or bf16 and share one dtype")
    B, H, Nq, D = q.shape
    Nkv = k.shape[2]
    if k.shape != (B, H, Nkv, D) or v.shape != k.shape:
        raise ValueError("shape mismatch between q, k, v")
    if scale is None:
        scale = D ** -0.5
    if out is None:
        out = torch.empty_like(q, pin_memory=q.is_pinned())
    elif out.shape != q.shape or out.dtype != q.dtype or out.is_cuda or not out.is_contiguous():
        raise ValueError("out must be a contiguous host tensor shaped and typed like q")
    lse = torch.empty((B, H, Nq), dtype=torch.float32, pin_memory=q.is_pinned()) if return_lse else None
    entry = _capi.lib.fa_fwd_sm100_host if wait else _capi.lib.fa_fwd_sm100_host_async
    rc = entry(
        q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(),
        lse.data_ptr() if lse is not None else None,
        B, H, Nq, Nkv, D, _dtype_code(q.dtype), int(bool(causal)), float(scale),
    )
    _capi.check(rc, "fa_fwd_sm100_host" if wait else "fa_fwd_sm100_host_async")
    return (out, lse) if return_lse else out


def flash_attn_host_sync():
    """Wait for every ``flash_attn_forward_host(..., wait=False)`` issued on the current device."""
    _capi.check(_capi.lib.fa_host_sync(), "fa_host_sync")


class _NativeModule:
    """Stand-in for the reference's pybind module object ``flash_attn_wmma`` (FlashAttn.py:23,
    host.cpp:60-64): ``forward(q,k,v,Br,Bc,causal,scale,permute_NH)`` returns the same six tensors
    ``[O_view, q_pad, k_pad, v_pad, O_pad, L]`` (kernel_fp16.cu:875).  Br/Bc are accepted for
    signature compatibility and ignored: tile sizes are internal to the sm_100a kernels."""

    @staticmethod
    def forward(q, k, v, Br, Bc, causal, scale, permute_NH):
        o, lse, (qp, kp, vp, o_full), _ = _forward(q, k, v, causal, scale, permute_NH, True)
        return [o, qp, kp, vp, o_full, lse]

    @staticmethod
    def backward(Q, K, V, O, dO, L, act_n, act_nkv, act_d, Br, Bc, causal, scale, permute_NH):
        """Reference signature (host.cpp:9-22,47-58): the tensors saved by ``forward`` (head dim
        already padded), the incoming gradient, the actual sizes, the tile sizes (ignored) -> the
        three gradients sliced to ``act_d`` (kernel_fp16.cu:1000-1027)."""
        return list(_backward(Q, K, V, O, dO, L, act_d, bool(causal), float(scale), bool(permute_NH)))


flash_attn_wmma = _NativeModule()


class FlashAttentionFunction(torch.autograd.Function):
    """Reference: /root/reference/rocwmma_fattn/FlashAttn.py:45-92."""

    @staticmethod
    @torch.no_grad()
    def forward(ctx, q, k, v, mask=None, causal=None, scale=None, BNHD_fmt=False, *args, **kwargs):
        need_grad = q.requires_grad or k.requires_grad or v.requires_grad
        o, lse, saved, meta = _forward(q, k, v, causal, scale, BNHD_fmt, need_grad)
        if need_grad:
            causal_, scale_, N, Nkv, D, bnhd = meta
            ctx.args = (causal_, scale_, mask, N, Nkv, D, bnhd)
            ctx.save_for_backward(*saved, lse)
        return o

    @staticmethod
    @torch.no_grad()
    def backward(ctx, do):
        """Reference: FlashAttn.py:80-92 (Br = Bc = 128 there; tile sizes are internal here)."""
        causal, scale, mask, N, Nkv, D, bnhd = ctx.args
        q, k, v, o, L = ctx.saved_tensors
        dQ, dK, dV = flash_attn_wmma.backward(q, k, v, o, do, L, N, Nkv, D, 128, 128, causal, scale, bnhd)
        return dQ, dK, dV, None, None, None, None
