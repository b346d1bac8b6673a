"""Golden fixtures for the BACKWARD path, generated from the UNMODIFIED reference oracle.

    python tests/golden/make_golden_bwd.py        (build container only: needs /root/reference)

For every case: seeded q, k, v, dO; forward + backward through
/root/reference/pure_torch_ver.py::FlashAttentionFunction (the reference's tiled oracle,
:24-153) and through autograd of math SDPA in fp32 and in the input dtype (the reference's own check,
:192-205 / precision_test.py:65-98).  Stored as
``bwd_<case>.npz``.  Sequence lengths are multiples of the oracle's tiles (Br=64, Bc=256): its
padding path mutates the saved K in place (:101) and is not a usable pin.
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("FA_REFERENCE_DIR", "/root/reference")

# name, B, H, N, D, dtype, causal
CASES = [
    ("bwd_f16_b1h2n256d64", 1, 2, 256, 64, torch.float16, False),
    ("bwd_f16_b1h2n256d64_causal", 1, 2, 256, 64, torch.float16, True),
    ("bwd_f16_b1h1n512d128", 1, 1, 512, 128, torch.float16, False),
    ("bwd_bf16_b1h1n512d128_causal", 1, 1, 512, 128, torch.bfloat16, True),
]


def _load_reference():
    spec = importlib.util.spec_from_file_location("pure_torch_ver", os.path.join(REF, "pure_torch_ver.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _bits(t: torch.Tensor) -> np.ndarray:
    return t.contiguous().view(torch.int16).numpy().view(np.uint16)


def main() -> None:
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))
    import fa_oracle as orc

    ref = _load_reference()
    for i, (name, B, H, N, D, dtype, causal) in enumerate(CASES):
        q, k, v = orc.make_inputs(B, H, N, N, D, dtype, seed=2000 + i)
        g = torch.Generator().manual_seed(3000 + i)
        d_o = torch.rand((B, H, N, D), generator=g, dtype=torch.float32).to(dtype)
        q1, k1, v1 = (t.clone().requires_grad_(True) for t in (q, k, v))
        o1 = ref.FlashAttentionFunction.apply(q1, k1, v1, None, causal)
        o1.backward(d_o)
        dq32, dk32, dv32 = orc.sdpa_backward(q, k, v, d_o, causal=causal)
        dq16, dk16, dv16 = orc.sdpa_backward(q, k, v, d_o, causal=causal, dtype=dtype)
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            q=_bits(q), k=_bits(k), v=_bits(v), d_o=_bits(d_o), o_ref_tiled=_bits(o1.detach()),
            dq_ref_tiled=_bits(q1.grad), dk_ref_tiled=_bits(k1.grad), dv_ref_tiled=_bits(v1.grad),
            dq_f32=dq32.numpy(), dk_f32=dk32.numpy(), dv_f32=dv32.numpy(),
            dq_sdpa16=_bits(dq16), dk_sdpa16=_bits(dk16), dv_sdpa16=_bits(dv16),
            dtype=np.array("float16" if dtype == torch.float16 else "bfloat16"),
            causal=np.array(causal), seed=np.array(2000 + i),
        )
        msg = []
        for nm, a, b, c16 in (("dq", q1.grad, dq32, dq16), ("dk", k1.grad, dk32, dk16), ("dv", v1.grad, dv32, dv16)):
            ok, err, bound = orc.check_close_grad(a, b, c16)
            msg.append(f"{nm} tiled-oracle err {err:.3e} bound {bound:.3e} max|ref| {b.abs().max().item():.3e} ok={ok}")
        print(name, "; ".join(msg))


if __name__ == "__main__":
    main()
