"""CPU tests of the drop-in boundary: the C-ABI library loads, exports exactly what
include/fa_fwd_sm100.h (the boundary) and include/fa_fwd_sm100_test.h (test hooks) declare, validates
arguments, and fails loudly (never falls back) when no sm_100 device is present.  No compute is launched here."""
import ctypes
import os
import re

import pytest
import torch

from rocwmma_fattn import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fa_fwd_sm100.h")
TEST_HEADER = os.path.join(ROOT, "include", "fa_fwd_sm100_test.h")


def _declared_functions(path):
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fa_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported():
    names = _declared_functions(HEADER)
    assert len(names) >= 9
    for n in names:
        assert hasattr(_capi.lib, n), f"{n} declared in include/fa_fwd_sm100.h but not exported"
    assert sorted(_capi.EXPORTED_SYMBOLS) == names
    # the boundary header carries no test hooks: they live in their own header
    assert not [n for n in names if n.startswith("fa_set_") or n.endswith("_selftest")]


def test_test_hook_header_symbols_all_exported():
    names = _declared_functions(TEST_HEADER)
    for n in names:
        assert hasattr(_capi.lib, n), f"{n} declared in include/fa_fwd_sm100_test.h but not exported"
    assert sorted(_capi.TEST_HOOK_SYMBOLS) == names


def test_abi_version_and_error_string():
    assert _capi.lib.fa_abi_version() == _capi.FA_ABI_VERSION == 5
    assert isinstance(_capi.last_error(), str)


def test_library_is_in_tree_and_native():
    assert os.path.dirname(_capi.LIB_PATH).startswith(ROOT)
    with open(_capi.LIB_PATH, "rb") as fh:
        assert fh.read(4) == b"\x7fELF"


def _st(B, H, N, D):
    return (H * N * D, N * D, D, 1)


@pytest.mark.parametrize(
    "shape,expected",
    [
        ((1, 16, 8192, 8192, 128), _capi.FA_KERNEL_SK),       # BASELINE sweep point: 3.46 rounds of work, 4 one-shot rounds
        ((1, 16, 4096, 4096, 128), _capi.FA_KERNEL_WS),       # 1.73 rounds: the split's L2 round trip eats the gain (model: tie)
        ((1, 16, 16384, 16384, 128), _capi.FA_KERNEL_WS),     # 6.92 rounds: 1 % to gain, one-shot kernel (ws2 retired from auto)
        ((64, 16, 4096, 4096, 128), _capi.FA_KERNEL_SK),      # BASELINE config 5: 110.7 rounds, cheap unit boundaries add up
        ((8, 16, 4096, 4096, 128), _capi.FA_KERNEL_SK),       # config 5 shard on 8 GPUs: 13.84 rounds
        ((1, 37, 4096, 4096, 128), _capi.FA_KERNEL_WS),       # 592 units = 4 x 148: nothing to balance
        ((1, 2, 128, 128, 64), _capi.FA_KERNEL_TC1),          # BASELINE config 1 shape
        ((3, 7, 1537, 1234, 112), _capi.FA_KERNEL_WS),        # precision_test.py after D pad
        ((1, 16, 1024, 1024, 128), _capi.FA_KERNEL_WIDE),     # sweep point: 128 tiles <= 148 SMs, one round of one-tile CTAs
        ((1, 16, 512, 512, 128), _capi.FA_KERNEL_WIDE),
        ((1, 16, 2048, 2048, 128), _capi.FA_KERNEL_WS),       # 256 tiles: two rounds -> two-tile kernel
        ((3, 7, 1537, 1234, 111), _capi.FA_KERNEL_SIMT),      # unpadded odd head dim
        ((1, 2, 300, 300, 256), _capi.FA_KERNEL_WIDE),        # head dim 129..256: one Q tile, two S buffers
        ((1, 16, 4096, 4096, 160), _capi.FA_KERNEL_WIDE),     # bench_with_sdpa.py:259-261 sweep point D = 16 * 10
        ((1, 2, 300, 300, 264), _capi.FA_KERNEL_SIMT),        # head dim > 256
        ((2, 8, 4096, 4096, 40), _capi.FA_KERNEL_SK),         # SD 1.5 head dim 40, 256 units = 1.73 rounds of pairs: persistent kernel (94.9 vs 100.0 us)
        ((2, 8, 16384, 16384, 40), _capi.FA_KERNEL_WS3),      # 1024 units x 128 tiles, 6.9 rounds of pairs: the early-S kernel on CTA pairs (1283 vs 1330 us)
        ((1, 16, 16384, 16384, 40), _capi.FA_KERNEL_WS3),
        ((2, 10, 4096, 4096, 64), _capi.FA_KERNEL_SK),        # SDXL: 320 units = 2.16 rounds -> persistent kernel (123 vs 147 us)
        ((2, 20, 1024, 1024, 64), _capi.FA_KERNEL_SK),        # SDXL 32x32 level: 160 units x 8 tiles (24.9 vs 31.4 us)
        ((2, 10, 4096, 77, 64), _capi.FA_KERNEL_SK),          # SDXL cross-attention, 320 one-tile units: unit boundaries beat CTA turnover (13.2 vs 16.2 us)
        ((2, 10, 4096, 77, 128), _capi.FA_KERNEL_SK),         # ... at head dim 128 too (14.8 vs 20.0 us)
        ((3, 10, 1536, 300, 64), _capi.FA_KERNEL_WS),         # 180 units x 3 tiles: the split would cost more than the half-empty round
        ((1, 4, 4096, 77, 64), _capi.FA_KERNEL_WIDE),         # 128 tiles: one round of one-tile CTAs, nothing for a persistent CTA to gain
        ((1, 8, 4096, 77, 64), _capi.FA_KERNEL_WS),           # 128 units: one round of two-tile CTAs
        ((1, 16, 2048, 2048, 64), _capi.FA_KERNEL_WS3),       # 128 units <= 148 SMs: never the persistent kernel at head dims <= 64
        ((1, 8, 1024, 1024, 64), _capi.FA_KERNEL_WIDE),       # 64 tiles: one round of one-tile CTAs still wins
        ((2, 4, 77, 300, 40), _capi.FA_KERNEL_TC1),
    ],
)
def test_kernel_selection_is_pure_host_logic(shape, expected):
    B, H, Nq, Nkv, D = shape
    k = _capi.select_kernel(B, H, Nq, Nkv, D, _st(B, H, Nq, D), _st(B, H, Nkv, D), _st(B, H, Nkv, D),
                            _st(B, H, Nq, D), _capi.FA_DTYPE_F16, False, D ** -0.5)
    assert k == expected


def test_kernel_selection_causal_head_dim_64():
    st = _st(1, 16, 16384, 64)
    assert _capi.select_kernel(1, 16, 16384, 16384, 64, st, st, st, st, _capi.FA_DTYPE_F16, True, 0.125) == _capi.FA_KERNEL_WS3
    st = _st(1, 16, 4096, 64)  # 256 blocks: more than one round, fewer than two - the two-tile kernel (LPT order)
    assert _capi.select_kernel(1, 16, 4096, 4096, 64, st, st, st, st, _capi.FA_DTYPE_F16, True, 0.125) == _capi.FA_KERNEL_WS
    st = _st(1, 16, 2048, 64)  # 128 blocks: one round - the one-tile kernel's 128-row grain
    assert _capi.select_kernel(1, 16, 2048, 2048, 64, st, st, st, st, _capi.FA_DTYPE_F16, True, 0.125) == _capi.FA_KERNEL_WIDE


@pytest.mark.parametrize(
    "n,expected",
    [
        (2048, _capi.FA_KERNEL_WIDE),   # BASELINE config 4 point: 128 blocks, one round -> 128-row scheduling grain
        (4096, _capi.FA_KERNEL_WS),     # 256 blocks, longest first across heads: the two-tile kernel (62 vs 72 us)
        (8192, _capi.FA_KERNEL_WS),
        (16384, _capi.FA_KERNEL_WS),
        (128, _capi.FA_KERNEL_TC1),
    ],
)
def test_kernel_selection_causal(n, expected):
    st = _st(1, 16, n, 128)
    assert _capi.select_kernel(1, 16, n, n, 128, st, st, st, st, _capi.FA_DTYPE_F16, True, 128 ** -0.5) == expected


def test_kernel_selection_bnhd_strides_and_bad_strides():
    B, H, N, D = 2, 8, 512, 128
    bnhd = (N * H * D, D, H * D, 1)  # logical (b,h,n,d) strides of a [B,N,H,D] tensor
    # a tensor-core kernel serves BNHD views in place (64 tiles <= 148 SMs: the one-tile arrangement)
    assert _capi.select_kernel(B, H, N, N, D, bnhd, bnhd, bnhd, bnhd, _capi.FA_DTYPE_BF16, False,
                               0.1) == _capi.FA_KERNEL_WIDE
    odd = (H * N * (D + 4), N * (D + 4), D + 4, 1)  # row stride not a multiple of 16 bytes
    assert _capi.select_kernel(B, H, N, N, D, odd, odd, odd, odd, _capi.FA_DTYPE_F16, False,
                               0.1) == _capi.FA_KERNEL_SIMT
    with pytest.raises(_capi.FlashAttnError) as ei:
        _capi.select_kernel(B, H, N, N, D, (1, 1, 1, 2), bnhd, bnhd, bnhd, _capi.FA_DTYPE_F16, False, 0.1)
    assert ei.value.code == _capi.FA_ERR_INVALID_ARG and "stride" in str(ei.value)
    # non-positive scale: handled by the generic kernel, not the tensor-core one
    assert _capi.select_kernel(B, H, N, N, D, _st(B, H, N, D), _st(B, H, N, D), _st(B, H, N, D),
                               _st(B, H, N, D), _capi.FA_DTYPE_F16, False, -1.0) == _capi.FA_KERNEL_SIMT


@pytest.mark.parametrize(
    "kw,code",
    [
        (dict(B=0), _capi.FA_ERR_INVALID_ARG),
        (dict(Nkv=0), _capi.FA_ERR_INVALID_ARG),
        (dict(dtype=7), _capi.FA_ERR_INVALID_ARG),
        (dict(D=2048), _capi.FA_ERR_UNSUPPORTED),
        (dict(scale=float("nan")), _capi.FA_ERR_INVALID_ARG),
    ],
)
def test_argument_validation(kw, code):
    a = dict(B=1, H=2, Nq=128, Nkv=128, D=64, dtype=_capi.FA_DTYPE_F16, scale=0.125)
    a.update(kw)
    st = _capi.strides4(_st(max(a["B"], 1), a["H"], a["Nq"], a["D"]))
    rc = _capi.lib.fa_select_kernel(a["B"], a["H"], a["Nq"], a["Nkv"], a["D"], st, st, st, st,
                                    a["dtype"], 0, a["scale"])
    assert rc == -code
    assert _capi.last_error() != ""


def test_null_pointers_rejected_before_any_device_work():
    st = _capi.strides4(_st(1, 2, 128, 64))
    rc = _capi.lib.fa_fwd_sm100(None, None, None, None, None, 1, 2, 128, 128, 64, st, st, st, st, 0, 0,
                                0.125, None)
    assert rc == _capi.FA_ERR_INVALID_ARG
    rc = _capi.lib.fa_fwd_sm100_host(None, None, None, None, None, 1, 2, 128, 128, 64, 0, 0, 0.125)
    assert rc == _capi.FA_ERR_INVALID_ARG


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_device_fails_loudly_no_cpu_fallback():
    buf = (ctypes.c_uint16 * (2 * 128 * 64))()
    p = ctypes.addressof(buf)
    st = _capi.strides4(_st(1, 2, 128, 64))
    rc = _capi.lib.fa_fwd_sm100(p, p, p, p, None, 1, 2, 128, 128, 64, st, st, st, st, 0, 0, 0.125, None)
    assert rc == _capi.FA_ERR_NO_DEVICE
    assert "device" in _capi.last_error().lower()
    assert _capi.launch_count() == 0


def test_set_kernel_roundtrip():
    prev = _capi.set_kernel(_capi.FA_KERNEL_SIMT)
    try:
        assert _capi.set_kernel(_capi.FA_KERNEL_AUTO) == _capi.FA_KERNEL_SIMT
        with pytest.raises(ValueError):
            _capi.set_kernel(99)
        for retired in (3, 8):  # FA_KERNEL_TC1_PSMEM / FA_KERNEL_QUAD2 (ABI 4)
            with pytest.raises(ValueError):
                _capi.set_kernel(retired)
    finally:
        _capi.set_kernel(prev)
