"""GPU parity tests of the BACKWARD path (run on a B200: ``pytest -m gpu``).

The gradients come from ``FlashAttentionFunction.apply(...).backward(dO)`` (autograd, exactly how the
reference's scripts drive it, pure_torch_ver.py:192-199 / precision_test.py:75-94) or from the
reference-shaped native call ``flash_attn_wmma.backward`` -> ctypes -> ``fa_bwd_sm100``; they are
checked against fp32 autograd of math SDPA (``oracle.sdpa_backward``) with the gate stated in
oracle/fa_oracle.py:  max|g - ref32| <= 2 max|g16 - ref32| + 1e-5 + 1e-3 max|ref32|, g16 = autograd of
math SDPA in the 16-bit input dtype (the reference's own comparator, precision_test.py:65-98), and
against the committed reference-generated fixtures (tests/golden/bwd_*.npz).
"""
import pytest
import torch

import fa_oracle as orc
from golden_util import golden_bwd_names, load_golden_bwd
from rocwmma_fattn import _capi
from rocwmma_fattn.FlashAttn import FlashAttentionFunction, flash_attn_wmma

pytestmark = pytest.mark.gpu

DEV = "cuda"
F16, BF16 = torch.float16, torch.bfloat16


def grads(q, k, v, d_o, causal=False, scale=None, bnhd=False):
    q, k, v = (t.detach().clone().requires_grad_(True) for t in (q, k, v))
    o = FlashAttentionFunction.apply(q, k, v, None, causal, scale, bnhd)
    o.backward(d_o)
    torch.cuda.synchronize()
    return o.detach(), q.grad, k.grad, v.grad


def assert_grads(got, ref, base16, dtype, what, atol=0.0):
    for nm, g, r, b16 in zip(("dq", "dk", "dv"), got, ref, base16):
        assert g.dtype == dtype and g.shape == r.shape, f"{what} {nm}: {g.dtype} {tuple(g.shape)}"
        assert orc.nan_free(g), f"{what} {nm}: non-finite"
        ok, err, bound = orc.check_close_grad(g, r, b16)
        assert ok or err <= atol, (f"{what} {nm}: max abs err {err:.3e} > bound {bound:.3e} "
                                   f"(max|ref| {r.abs().max().item():.3e})")


def truth(q, k, v, d_o, causal=False, scale=None):
    """(fp32 autograd, 16-bit autograd) of math SDPA on CPU copies."""
    q, k, v, d_o = (t.detach().cpu() for t in (q, k, v, d_o))
    return (orc.sdpa_backward(q, k, v, d_o, causal=causal, scale=scale),
            orc.sdpa_backward(q, k, v, d_o, causal=causal, scale=scale, dtype=q.dtype))


@pytest.mark.parametrize("bnhd", [False, True])
@pytest.mark.parametrize("name", golden_bwd_names())
def test_backward_golden_vectors(name, bnhd):
    g = load_golden_bwd(name)
    q, k, v, d_o = (g[x].to(DEV) for x in ("q", "k", "v", "d_o"))
    if bnhd:
        q, k, v, d_o = (t.transpose(1, 2).contiguous() for t in (q, k, v, d_o))
    _, dq, dk, dv = grads(q, k, v, d_o, g["causal"], None, bnhd)
    if bnhd:
        dq, dk, dv = (t.transpose(1, 2) for t in (dq, dk, dv))
    ref = (g["dq_f32"], g["dk_f32"], g["dv_f32"])
    assert_grads((dq, dk, dv), ref, (g["dq_sdpa16"], g["dk_sdpa16"], g["dv_sdpa16"]), g["dtype"], name)
    # at least as close to fp32 autograd as the reference's own tiled oracle
    for nm, ours, theirs, r in zip(("dq", "dk", "dv"), (dq, dk, dv),
                                   (g["dq_ref_tiled"], g["dk_ref_tiled"], g["dv_ref_tiled"]), ref):
        e_ours, e_ref = orc.max_abs_err(ours, r), orc.max_abs_err(theirs, r)
        assert e_ours <= e_ref * 1.05 + 1e-6, f"{name} {nm}: ours {e_ours:.3e} vs reference oracle {e_ref:.3e}"


SHAPES = [
    # B, H, Nq, Nkv, D
    (1, 2, 128, 128, 64),     # BASELINE config 1
    (2, 3, 256, 256, 128),
    (1, 4, 512, 512, 128),
    (1, 2, 1000, 1000, 128),  # unaligned
    (2, 2, 200, 77, 64),      # cross-attention style Nkv = 77
    (1, 3, 193, 150, 112),    # head dim padded to 112 -> 128 column kernel
    (1, 2, 129, 1, 128),
    (1, 2, 77, 300, 40),      # causal with more keys than queries: trailing keys get zero grads
    (3, 7, 1537, 1234, 111),  # precision_test.py:34-39 (odd head dim: padded to 112)
]


@pytest.fixture
def bwd_kernel(request):
    """Force a backward kernel for one test (fa_set_bwd_kernel, test-hook header) and restore the default."""
    prev = _capi.set_bwd_kernel(request.param)
    yield request.param
    _capi.set_bwd_kernel(prev)


@pytest.mark.parametrize("bwd_kernel", [_capi.FA_BWD_KERNEL_WS, _capi.FA_BWD_KERNEL_TC], indirect=True,
                         ids=["pipelined", "serial"])
@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("dtype", [F16, BF16])
@pytest.mark.parametrize("shape", SHAPES)
def test_backward_matches_fp32_autograd(shape, dtype, causal, bwd_kernel):
    B, H, Nq, Nkv, D = shape
    q, k, v = orc.make_inputs(B, H, Nq, Nkv, D, dtype, seed=sum(shape) + 7 * int(causal))
    g = torch.Generator().manual_seed(99 + sum(shape))
    d_o = torch.rand((B, H, Nq, D), generator=g).to(dtype)
    ref, b16 = truth(q, k, v, d_o, causal)
    _, dq, dk, dv = grads(q.to(DEV), k.to(DEV), v.to(DEV), d_o.to(DEV), causal)
    # a single key: softmax is identically 1, the exact dQ and dK are 0 and plain PyTorch returns
    # exactly 0, so the relative gate degenerates; what remains is (dP - rowsum(dO o O16)) summed over
    # the query rows, i.e. the 16-bit rounding of the stored O (any FA2 backward has it)
    atol = 2e-4 if Nkv == 1 else 0.0
    assert_grads((dq, dk, dv), ref, b16, dtype, f"{shape} {dtype} causal={causal}", atol)


def test_default_backward_kernel_is_the_pipelined_one_and_agrees_with_the_serial_kernel():
    """dK and dV do not depend on any summation order across CTAs, so the two kernels must agree to rounding of
    their 16-bit P / dS intermediates; dQ is an fp32 reduce-add in both."""
    assert _capi.set_bwd_kernel(_capi.FA_BWD_KERNEL_AUTO) == _capi.FA_BWD_KERNEL_AUTO  # nothing forced by default
    B, H, N, D = 2, 4, 1024, 128
    q, k, v = orc.make_inputs(B, H, N, N, D, F16, seed=5)
    d_o = torch.rand((B, H, N, D), generator=torch.Generator().manual_seed(6)).to(F16)
    args = (q.to(DEV), k.to(DEV), v.to(DEV), d_o.to(DEV))
    n0 = _capi.launch_count()
    auto = grads(*args, True)[1:]
    assert _capi.launch_count() - n0 == 4  # forward, delta pre-pass, main kernel, dQ conversion
    prev = _capi.set_bwd_kernel(_capi.FA_BWD_KERNEL_TC)
    try:
        serial = grads(*args, True)[1:]
    finally:
        _capi.set_bwd_kernel(prev)
    for nm, a, b in zip(("dq", "dk", "dv"), auto, serial):
        err = (a.float() - b.float()).abs().max().item()
        assert err <= 2e-3 * max(1e-3, b.float().abs().max().item()) + 1e-5, f"{nm}: kernels disagree by {err:.3e}"


def test_backward_randn_inputs_and_custom_scale():
    B, H, N, D = 1, 3, 384, 64
    q, k, v = orc.make_inputs(B, H, N, N, D, F16, seed=17, dist="randn")
    d_o = torch.randn((B, H, N, D), generator=torch.Generator().manual_seed(5)).to(F16)
    for causal in (False, True):
        ref, b16 = truth(q, k, v, d_o, causal, 0.07)
        _, dq, dk, dv = grads(q.to(DEV), k.to(DEV), v.to(DEV), d_o.to(DEV), causal, 0.07)
        assert_grads((dq, dk, dv), ref, b16, F16, f"randn causal={causal}")


def test_backward_strided_views_and_bnhd():
    """q, k, v as slices of one packed [B,N,3,H,D] buffer (BNHD layout, no copies) and a
    non-contiguous incoming gradient."""
    B, H, N, D = 2, 4, 320, 128
    g = torch.Generator().manual_seed(3)
    qkv = torch.rand((B, N, 3, H, D), generator=g).to(BF16).to(DEV)
    d_o = torch.rand((B, H, N, D), generator=g).to(BF16).to(DEV).transpose(1, 2)  # [B,N,H,D] view
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    _, dq, dk, dv = grads(q, k, v, d_o, True, None, True)
    qc, kc, vc, dc = (t.transpose(1, 2).cpu() for t in (q, k, v, d_o))
    ref, b16 = truth(qc, kc, vc, dc, True)
    got = tuple(t.transpose(1, 2) for t in (dq, dk, dv))
    assert_grads(got, ref, b16, BF16, "packed qkv, BNHD")


def test_native_module_backward_signature():
    """flash_attn_wmma.backward(Q,K,V,O,dO,L,N,Nkv,D,Br,Bc,causal,scale,permute_NH) -> [dQ,dK,dV]
    (host.cpp:9-22,47-58) on the tensors forward() returned."""
    B, H, N, D = 1, 2, 256, 64
    q, k, v = (t.to(DEV) for t in orc.make_inputs(B, H, N, N, D, F16, seed=2))
    d_o = torch.rand((B, H, N, D), generator=torch.Generator().manual_seed(8)).to(F16).to(DEV)
    o, qp, kp, vp, o_pad, L = flash_attn_wmma.forward(q, k, v, 64, 128, False, D ** -0.5, False)
    n0 = _capi.launch_count()
    out = flash_attn_wmma.backward(qp, kp, vp, o_pad, d_o, L, N, N, D, 128, 128, False, D ** -0.5, False)
    torch.cuda.synchronize()
    assert _capi.launch_count() == n0 + 3  # pre-pass, main kernel, dQ conversion
    assert isinstance(out, list) and len(out) == 3
    ref, b16 = truth(q, k, v, d_o)
    assert_grads(tuple(out), ref, b16, F16, "native backward")


def test_backward_is_deterministic_up_to_atomic_order_and_dq_linear_in_do():
    """dK and dV are reduced inside one CTA: bit-identical across runs.  dQ is summed across key
    tiles with fp32 atomics: equal up to fp32 reassociation.  All three are linear in dO."""
    B, H, N, D = 1, 2, 640, 128
    q, k, v = (t.to(DEV) for t in orc.make_inputs(B, H, N, N, D, F16, seed=11))
    d_o = torch.rand((B, H, N, D), generator=torch.Generator().manual_seed(1)).to(F16).to(DEV)
    _, dq1, dk1, dv1 = grads(q, k, v, d_o, True)
    _, dq2, dk2, dv2 = grads(q, k, v, d_o, True)
    assert torch.equal(dk1, dk2) and torch.equal(dv1, dv2)
    assert orc.max_abs_err(dq1, dq2) <= 2e-3 * dq1.float().abs().max().item()
    _, dq3, dk3, dv3 = grads(q, k, v, d_o * 2, True)
    for a, b in ((dq1, dq3), (dk1, dk3), (dv1, dv3)):
        assert orc.max_abs_err(a.float() * 2, b) <= 4e-3 * b.float().abs().max().item()


def test_backward_baseline_shape_sampled_rows():
    """BASELINE sweep shape (B=1, H=16, D=128) at N=2048 against fp32 autograd on the device."""
    B, H, N, D = 1, 16, 2048, 128
    q, k, v = (t.to(DEV) for t in orc.make_inputs(B, H, N, N, D, BF16, seed=4))
    d_o = torch.rand((B, H, N, D), generator=torch.Generator().manual_seed(6)).to(BF16).to(DEV)
    _, dq, dk, dv = grads(q, k, v, d_o, False)
    ref = orc.sdpa_backward(q, k, v, d_o)  # fp32 on the GPU
    b16 = orc.sdpa_backward(q, k, v, d_o, dtype=BF16)  # plain 16-bit PyTorch on the GPU
    assert_grads((dq, dk, dv), tuple(r.cpu() for r in ref), tuple(b.cpu() for b in b16), BF16,
                 "sweep shape N=2048")


WIDE_BWD_SHAPES = [
    # B, H, Nq, Nkv, D: head dims above 128 (the reference pads and serves any, kernel_fp16.cu:900).  129..256 run the
    # three-launch tcgen05 kernel (csrc/fa_bwd_wide.cuh), larger ones the generic CUDA-core backward (csrc/fa_bwd_simt.cuh)
    (1, 2, 300, 300, 160),   # SD 1.5 head dim
    (2, 1, 257, 130, 192),
    (1, 2, 128, 384, 256),
    (1, 2, 640, 640, 256),   # several streamed tiles per CTA: both stages of the ring wrap, single-stage modes too
    (2, 3, 1000, 700, 160),  # unaligned, Nq != Nkv, more queries than keys
    (1, 2, 130, 1000, 232),  # causal with keys beyond the last query: whole key tiles get zero gradients
    (1, 1, 200, 77, 264),    # above 256: generic forward AND backward
    (1, 2, 193, 150, 135),   # odd head dim, padded to 136
]


def _bwd_launches(D):
    """Launches of one backward call: pre-pass + (dV/dK fused + dQ at head dims 129..192 | dV + dK + dQ at 193..256 of the
    three-launch tcgen05 kernel | dQ kernel + dK/dV kernel of the generic path above 256 or off the 8-element grid)."""
    DP = -(-D // 8) * 8
    return 1 + (2 if 128 < DP <= 192 else 3 if 192 < DP <= 256 else 2)


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("dtype", [F16, BF16])
@pytest.mark.parametrize("shape", WIDE_BWD_SHAPES)
def test_backward_head_dims_above_128_match_fp32_autograd(shape, dtype, causal):
    B, H, Nq, Nkv, D = shape
    q, k, v = orc.make_inputs(B, H, Nq, Nkv, D, dtype, seed=sum(shape) + int(causal))
    g = torch.Generator().manual_seed(5 + sum(shape))
    d_o = torch.rand((B, H, Nq, D), generator=g).to(dtype)
    ref, b16 = truth(q, k, v, d_o, causal)
    n0 = _capi.launch_count()
    _, dq, dk, dv = grads(q.to(DEV), k.to(DEV), v.to(DEV), d_o.to(DEV), causal)
    assert _capi.launch_count() == n0 + 1 + _bwd_launches(D)  # forward + backward
    assert_grads((dq, dk, dv), ref, b16, dtype, f"{shape} {dtype} causal={causal}")


def test_wide_backward_bnhd_layout_and_packed_qkv_views():
    """[B,N,H,D] tensors that are views into one packed [B,N,3,H,D] buffer: the 64-row and 128-row tensor maps of the
    head-dim 129..256 backward take the strides as they are (no copies), gradients come back in the same layout."""
    B, H, N, D = 2, 3, 320, 160
    qkv = torch.rand((B, N, 3, H, D), generator=torch.Generator().manual_seed(21)).to(F16).to(DEV)
    q, k, v = (qkv[:, :, i] for i in range(3))  # BNHD views, row stride 3*H*D
    d_o = torch.rand((B, N, H, D), generator=torch.Generator().manual_seed(22)).to(F16).to(DEV)
    n0 = _capi.launch_count()
    _, dq, dk, dv = grads(q, k, v, d_o, True, None, True)
    assert _capi.launch_count() == n0 + 1 + _bwd_launches(D)
    bhnd = lambda t: t.transpose(1, 2)  # noqa: E731
    ref, b16 = truth(bhnd(q), bhnd(k), bhnd(v), bhnd(d_o), True)
    assert_grads(tuple(bhnd(t) for t in (dq, dk, dv)), ref, b16, F16, "BNHD packed views, D=160")


@pytest.mark.parametrize("causal", [False, True])
def test_wide_backward_agrees_with_the_generic_cuda_core_backward(causal):
    """fa_set_bwd_kernel(3) forces the CUDA-core kernels at head dims 129..256 too: two independent implementations
    of the same gradients (fp32 math there, 16-bit P / dS on the tensor cores here)."""
    B, H, Nq, Nkv, D = 1, 2, 384, 320, 192
    q, k, v = orc.make_inputs(B, H, Nq, Nkv, D, F16, seed=11)
    d_o = torch.rand((B, H, Nq, D), generator=torch.Generator().manual_seed(12)).to(F16)
    args = (q.to(DEV), k.to(DEV), v.to(DEV), d_o.to(DEV))
    tc = grads(*args, causal)[1:]
    prev = _capi.set_bwd_kernel(_capi.FA_BWD_KERNEL_SIMT_ABOVE_128)
    try:
        n0 = _capi.launch_count()
        simt = grads(*args, causal)[1:]
        assert _capi.launch_count() == n0 + 1 + 3
    finally:
        _capi.set_bwd_kernel(prev)
    for nm, a, b in zip(("dq", "dk", "dv"), tc, simt):
        err = (a.float() - b.float()).abs().max().item()
        assert err <= 4e-3 * max(1e-3, b.float().abs().max().item()) + 1e-5, f"{nm}: kernels disagree by {err:.3e}"


def test_backward_of_unaligned_views_takes_the_tensor_core_kernel_on_contiguous_copies():
    """the forward serves views TMA cannot address (an odd element offset into a packed buffer)
    with its generic kernel; the backward must not fail on the tensors it saved."""
    B, H, N, D = 1, 2, 256, 64
    base = torch.rand((3, B, H, N, D + 4), generator=torch.Generator().manual_seed(3)).to(F16).to(DEV)
    q, k, v = (base[i, ..., 2:2 + D] for i in range(3))  # last stride 1, row stride 68, base offset 4 bytes
    assert q.data_ptr() % 16 != 0 and q.stride(2) % 8 != 0
    d_o = torch.rand((B, H, N, D), generator=torch.Generator().manual_seed(4)).to(F16).to(DEV)
    _, dq, dk, dv = grads(q, k, v, d_o, True)
    ref, b16 = truth(q, k, v, d_o, True)
    assert_grads((dq, dk, dv), ref, b16, F16, "unaligned views")
