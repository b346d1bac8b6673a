"""CPU tests of the Python host layer that mirrors /root/reference/rocwmma_fattn/FlashAttn.py:
argument handling that needs no device, and the loud failure without one."""
import inspect

import pytest
import torch

from rocwmma_fattn import FlashAttn
from rocwmma_fattn.FlashAttn import FlashAttentionFunction, _logical_shape, _logical_strides


def test_signature_matches_reference_entry_point():
    """FlashAttn.py:49 — forward(ctx, q, k, v, mask=None, causal=None, scale=None, BNHD_fmt=False, *args, **kwargs)"""
    sig = inspect.signature(FlashAttentionFunction.forward)
    names = list(sig.parameters)
    assert names[:8] == ["ctx", "q", "k", "v", "mask", "causal", "scale", "BNHD_fmt"]
    assert sig.parameters["mask"].default is None
    assert sig.parameters["causal"].default is None
    assert sig.parameters["scale"].default is None
    assert sig.parameters["BNHD_fmt"].default is False
    assert issubclass(FlashAttentionFunction, torch.autograd.Function)
    assert hasattr(FlashAttn, "flash_attn_wmma") and hasattr(FlashAttn.flash_attn_wmma, "forward")


def test_bnhd_stride_identity():
    """The identity /root/reference/test_arrange.py:23-30 checks: element (b,h,n,d) of a [B,N,H,D]
    tensor addressed through logical (b,h,n,d) strides equals the BHND-contiguous fetch."""
    B, N, H, D = 2, 5, 3, 4
    t = torch.arange(B * N * H * D).reshape(B, N, H, D)
    st = _logical_strides(t, True)
    assert _logical_shape(t, True) == (B, H, N, D)
    bhnd = t.permute(0, 2, 1, 3).contiguous()
    flat = t.reshape(-1)
    for b, h, n, d in [(0, 0, 0, 0), (1, 2, 4, 3), (0, 1, 3, 2), (1, 0, 2, 1)]:
        assert flat[b * st[0] + h * st[1] + n * st[2] + d * st[3]] == bhnd[b, h, n, d]
    assert _logical_strides(bhnd, False) == bhnd.stride()


def test_cpu_tensors_raise_no_fallback():
    q = torch.rand(1, 2, 16, 8, dtype=torch.float16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        FlashAttentionFunction.apply(q, q, q)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        FlashAttn.flash_attn_forward(q, q, q)


def test_bad_rank_raises():
    q = torch.rand(2, 16, 8)
    with pytest.raises(ValueError):
        FlashAttn.flash_attn_forward(q, q, q)


def test_host_path_argument_checks():
    q = torch.rand(1, 2, 16, 8, dtype=torch.float32)
    with pytest.raises(TypeError):
        FlashAttn.flash_attn_forward_host(q, q, q)
    h = q.half()
    with pytest.raises(ValueError):
        FlashAttn.flash_attn_forward_host(h, h[:, :, :8], h[:, :1])
    with pytest.raises(ValueError):
        FlashAttn.flash_attn_forward_host(h.transpose(1, 2), h, h)  # not contiguous


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_host_path_fails_loudly_without_gpu():
    h = torch.rand(1, 2, 16, 8).half()
    with pytest.raises(RuntimeError, match="device"):
        FlashAttn.flash_attn_forward_host(h, h, h)


def test_backward_entry_has_the_reference_signature_and_no_cpu_fallback():
    """host.cpp:9-22: backward(Q,K,V,O,dO,L,act_n,act_nkv,act_d,Br,Bc,causal,scale,permute_NH)."""
    import inspect

    names = list(inspect.signature(FlashAttn.flash_attn_wmma.backward).parameters)
    assert names == ["Q", "K", "V", "O", "dO", "L", "act_n", "act_nkv", "act_d", "Br", "Bc", "causal",
                     "scale", "permute_NH"]
    t = torch.zeros(1, 1, 8, 8, dtype=torch.float16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        FlashAttn.flash_attn_wmma.backward(t, t, t, t, t, torch.zeros(1, 1, 8), 8, 8, 8, 128, 128, False,
                                           0.35, False)


def test_sdpa_front_end_is_strict_and_installable():
    """rocwmma_fattn.sdpa: torch-SDPA-shaped adapter; unsupported calls raise unless
    the caller passes a fallback explicitly; install() / uninstall() swap torch's function."""
    from rocwmma_fattn import sdpa

    q = torch.zeros(1, 2, 8, 16, dtype=torch.float16)
    with pytest.raises(NotImplementedError, match="CUDA"):
        sdpa.scaled_dot_product_attention(q, q, q)
    assert "attn_mask" in sdpa.supported(q, q, q, attn_mask=torch.ones(8, 8, dtype=torch.bool))
    assert "dropout" in sdpa.supported(q, q, q, dropout_p=0.1)
    calls = []

    def fb(query, key, value, **kw):
        calls.append(kw)
        return query

    assert sdpa.scaled_dot_product_attention(q, q, q, is_causal=True, fallback=fb) is q
    assert calls and calls[0]["is_causal"] is True
    stock = torch.nn.functional.scaled_dot_product_attention
    sdpa.install()
    try:
        assert torch.nn.functional.scaled_dot_product_attention is not stock
        with pytest.raises(NotImplementedError):
            torch.nn.functional.scaled_dot_product_attention(q, q, q)
    finally:
        sdpa.uninstall()
    assert torch.nn.functional.scaled_dot_product_attention is stock


def test_host_chunk_planner_covers_every_head_and_follows_its_cost_model(monkeypatch):
    """fa_fwd_sm100_host() pipelines H2D -> kernel -> D2H over chunks of whole heads; the planner is pure
    host logic (fa_host_plan_chunks).  Sizes must sum to B*H, the count grows with the work that can be
    hidden, and only the tail is shortened."""
    from rocwmma_fattn import _capi

    monkeypatch.delenv("FA_HOST_CHUNKS", raising=False)
    counts = {}
    for n in (512, 1024, 2048, 4096, 8192, 16384):
        c = _capi.host_plan_chunks(1, 16, n, n, 128)
        assert sum(c) == 16 and min(c) >= 1
        assert c == sorted(c, reverse=True)  # uniform chunks, then a shrinking tail
        counts[n] = len(c)
    assert counts[512] == 1                       # launch-bound: one chunk, no pipeline overhead
    assert counts[16384] >= 8                     # PCIe-bound: enough chunks to hide kernel and D2H
    assert all(counts[a] <= counts[b] for a, b in zip((512, 1024, 2048, 4096, 8192), (1024, 2048, 4096, 8192, 16384)))
    # ragged head counts and a single head
    for heads in (1, 3, 5, 7, 1024):
        c = _capi.host_plan_chunks(1, heads, 4096, 4096, 128, causal=True)
        assert sum(c) == heads and min(c) >= 1
    # explicit override: uniform, no tail split
    monkeypatch.setenv("FA_HOST_CHUNKS", "4")
    assert _capi.host_plan_chunks(1, 16, 16384, 16384, 128) == [4, 4, 4, 4]
    monkeypatch.setenv("FA_HOST_CHUNKS", "64")
    assert _capi.host_plan_chunks(1, 16, 512, 512, 128) == [1] * 16
    with pytest.raises(_capi.FlashAttnError):
        _capi.host_plan_chunks(0, 16, 512, 512, 128)


def test_device_local_cpus_parses_sysfs(tmp_path):
    """shard.device_local_cpus reads the GPU's PCI device node (local_cpulist) to find its NUMA-local CPUs."""
    import shard

    d = tmp_path / "0000:1b:00.0"
    d.mkdir()
    (d / "local_cpulist").write_text("0-3,32-35\n")
    assert shard.device_local_cpus(0, 0x1B, 0, sysfs=str(tmp_path)) == [0, 1, 2, 3, 32, 33, 34, 35]
    assert shard.device_local_cpus(0, 0x1C, 0, sysfs=str(tmp_path)) == []
    (d / "local_cpulist").write_text("7\n")
    assert shard.device_local_cpus(0, 0x1B, 0, sysfs=str(tmp_path)) == [7]
