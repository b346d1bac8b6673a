"""CPU tests: pin oracle/fa_oracle.py against the golden vectors produced by the reference's own
pure_torch_ver.py (tests/golden/make_golden.py), and check its internal consistency."""
import pytest
import torch

import fa_oracle as orc
from golden_util import golden_bwd_names, golden_names, load_golden, load_golden_bwd


@pytest.mark.parametrize("name", golden_names())
def test_inputs_reproducible_from_seed(name):
    g = load_golden(name)
    B, H, Nq, D = g["q"].shape
    Nkv = g["k"].shape[2]
    q, k, v = orc.make_inputs(B, H, Nq, Nkv, D, g["dtype"], seed=g["seed"], dist=g["dist"])
    assert torch.equal(q, g["q"]) and torch.equal(k, g["k"]) and torch.equal(v, g["v"])


@pytest.mark.parametrize("name", golden_names())
def test_tiled_restatement_matches_reference_tiled_oracle(name):
    """oracle.tiled_fa2_forward restates pure_torch_ver.py:24-90; same dtype arithmetic, so it must
    sit within a few ulps of the reference's stored output (matmul vs einsum summation order)."""
    g = load_golden(name)
    o, L = orc.tiled_fa2_forward(g["q"], g["k"], g["v"], causal=g["causal"])
    tol = 2e-3 if g["dtype"] == torch.float16 else 1.6e-2  # <= 4 ulp at |o| ~ 0.5
    assert orc.max_abs_err(o, g["o_ref_tiled"]) <= tol
    assert o.shape == g["q"].shape and L.shape == g["q"].shape[:3]


@pytest.mark.parametrize("name", golden_names())
def test_sdpa_math_matches_reference_fp32_sdpa(name):
    g = load_golden(name)
    o, _ = orc.sdpa_math(g["q"], g["k"], g["v"], causal=g["causal"])
    assert orc.max_abs_err(o, g["o_ref_f32"]) <= 2e-6


@pytest.mark.parametrize("name", golden_names())
def test_cpu_sdpa_matches_reference_sdpa(name):
    g = load_golden(name)
    o = orc.cpu_sdpa(g["q"], g["k"], g["v"], causal=g["causal"])
    # same torch call the generator made; allow one ulp for thread-count dependent summation
    tol = 5e-4 if g["dtype"] == torch.float16 else 4e-3
    assert orc.max_abs_err(o, g["o_ref_sdpa"]) <= tol


@pytest.mark.parametrize("name", golden_names())
def test_reference_outputs_within_adopted_tolerance(name):
    """The tolerance the GPU parity tests use must admit the reference's own 16-bit
    SDPA output, otherwise it would be tighter than the reference is to itself."""
    g = load_golden(name)
    ok, err = orc.check_close(g["o_ref_sdpa"], g["o_ref_f32"], g["dtype"], g["dist"])
    assert ok, err


def test_lse_consistency_and_base2():
    q, k, v = orc.make_inputs(2, 3, 70, 45, 32, torch.float16, seed=5, dist="randn")
    for causal in (False, True):
        _, lse2 = orc.sdpa_math(q, k, v, causal=causal, dtype=torch.float64)
        l2 = orc.lse_base2(q, k, causal=causal)
        assert (lse2 - l2).abs().max().item() < 1e-9
        # natural-log LSE of the tiled oracle (pure_torch_ver.py:84) * log2(e) == base-2 LSE
        # (on non-negative inputs: the tiled oracle's -100 padding needs q >= 0)
        qn, kn, vn = (t.abs() for t in (q, k, v))
        _, Ln = orc.tiled_fa2_forward(qn.float(), kn.float(), vn.float(), causal=causal, Br=32, Bc=16)
        assert (Ln.double() * orc.LOG2E - orc.lse_base2(qn, kn, causal=causal)).abs().max().item() < 1e-3


@pytest.mark.parametrize("nq,nkv", [(5, 9), (9, 5), (64, 64), (1, 7)])
def test_causal_is_top_left_aligned_like_sdpa(nq, nkv):
    q, k, v = orc.make_inputs(1, 2, nq, nkv, 16, torch.float32, seed=nq * 31 + nkv, dist="randn")
    o, _ = orc.sdpa_math(q, k, v, causal=True)
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=True)
    assert orc.max_abs_err(o, ref) < 1e-5
    # row 0 only sees key 0
    assert orc.max_abs_err(o[:, :, 0], v[:, :, 0]) < 1e-6


def test_tiled_oracle_pad_trick_only_valid_for_nonnegative_q():
    """Documents SURVEY section 4: the reference oracle's -100 padding masks correctly only for
    q >= 0, which is why signed-input parity is anchored on sdpa_math instead."""
    q, k, v = orc.make_inputs(1, 1, 100, 77, 32, torch.float32, seed=3, dist="rand")
    o_t, _ = orc.tiled_fa2_forward(q, k, v)
    o_r, _ = orc.sdpa_math(q, k, v)
    assert orc.max_abs_err(o_t, o_r) < 1e-4


def test_flop_and_byte_counts_match_baseline_table():
    assert orc.attention_flops(1, 16, 512, 512, 128) == pytest.approx(2.147e9, rel=1e-3)
    assert orc.attention_flops(1, 16, 16384, 16384, 128) == pytest.approx(2.199e12, rel=1e-3)
    assert orc.attention_flops(1, 16, 4096, 4096, 128, causal=True) == pytest.approx(0.5 * 1.374e11, rel=1e-3)
    assert orc.attention_bytes(1, 16, 512, 512, 128) == 16384 * 512
    assert orc.attention_flops(1, 2, 128, 128, 64) == pytest.approx(8.39e6, rel=1e-3)


# ---------------------------------------------------------------------------------------------
# backward oracle (pure_torch_ver.py:94-153) against the reference-generated fixtures
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", golden_bwd_names())
def test_backward_ground_truth_matches_stored_fp32_autograd(name):
    g = load_golden_bwd(name)
    dq, dk, dv = orc.sdpa_backward(g["q"], g["k"], g["v"], g["d_o"], causal=g["causal"])
    for a, b in ((dq, g["dq_f32"]), (dk, g["dk_f32"]), (dv, g["dv_f32"])):
        assert orc.max_abs_err(a, b) <= 1e-5 * max(1.0, b.abs().max().item())


@pytest.mark.parametrize("name", golden_bwd_names())
def test_tiled_backward_restatement_matches_reference_tiled_oracle(name):
    """oracle.tiled_fa2_backward restates pure_torch_ver.py:94-153 in the same 16-bit arithmetic: it
    must land within a few ulps (of the largest entry) of what the reference's oracle produced."""
    g = load_golden_bwd(name)
    o, L = orc.tiled_fa2_forward(g["q"], g["k"], g["v"], causal=g["causal"])
    dq, dk, dv = orc.tiled_fa2_backward(g["q"], g["k"], g["v"], o, L.to(g["dtype"]), g["d_o"], causal=g["causal"])
    ulp = 2.0 ** -10 if g["dtype"] == torch.float16 else 2.0 ** -7
    for nm, a, b in (("dq", dq, g["dq_ref_tiled"]), ("dk", dk, g["dk_ref_tiled"]), ("dv", dv, g["dv_ref_tiled"])):
        assert orc.max_abs_err(a, b) <= 8 * ulp * b.float().abs().max().item() + 1e-4, nm


@pytest.mark.parametrize("name", golden_bwd_names())
def test_backward_16bit_baseline_reproducible_and_gate_admits_it(name):
    """The 16-bit math-SDPA gradients the gate is built on are reproducible here, and the gate
    trivially admits them (so it is never tighter than plain PyTorch is to fp32)."""
    g = load_golden_bwd(name)
    got = orc.sdpa_backward(g["q"], g["k"], g["v"], g["d_o"], causal=g["causal"], dtype=g["dtype"])
    for nm, a in zip(("dq", "dk", "dv"), got):
        b, r = g[nm + "_sdpa16"], g[nm + "_f32"]
        assert orc.max_abs_err(a, b) <= 4e-3 * r.abs().max().item() + 1e-6, nm
        ok, err, bound = orc.check_close_grad(b, r, b)
        assert ok, (nm, err, bound)
