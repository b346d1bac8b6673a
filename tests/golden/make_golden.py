"""Generate the golden fixtures in this directory from the UNMODIFIED reference oracle.

Run in the build container (where /root/reference is mounted):

    python tests/golden/make_golden.py

For every case it draws seeded inputs, runs
  * /root/reference/pure_torch_ver.py::FlashAttentionFunction.apply   (the reference's tiled oracle)
  * torch.nn.functional.scaled_dot_product_attention                  (the reference's SDPA check,
    pure_torch_ver.py:181 / precision_test.py:65) in the input dtype and in fp32
and stores inputs and outputs as ``<case>.npz`` (16-bit tensors as uint16 bit patterns).
The GPU box has no /root/reference: tests only read the committed .npz files.
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("FA_REFERENCE_DIR", "/root/reference")

# name, B, H, Nq, Nkv, D, dtype, causal, dist
CASES = [
    # BASELINE config 1: fp16 fwd B=1 H=2 N=128 D=64 non-causal (the reference's CPU-runnable case)
    ("c1_f16_b1h2n128d64", 1, 2, 128, 128, 64, torch.float16, False, "rand"),
    ("c1_f16_b1h2n128d64_causal", 1, 2, 128, 128, 64, torch.float16, True, "rand"),
    ("bf16_b1h2n128d64", 1, 2, 128, 128, 64, torch.bfloat16, False, "rand"),
    # head dim of the sweep configs, several KV tiles
    ("f16_b1h1n320d128", 1, 1, 320, 320, 128, torch.float16, False, "rand"),
    ("f16_b1h1n320d128_causal", 1, 1, 320, 320, 128, torch.float16, True, "rand"),
    ("bf16_b1h1n320d128_causal", 1, 1, 320, 320, 128, torch.bfloat16, True, "rand"),
    # unaligned sequence lengths, Nq != Nkv (precision_test.py:34-39 in miniature)
    ("f16_b2h3n200k77d64", 2, 3, 200, 77, 64, torch.float16, False, "rand"),
    ("bf16_b1h3n193k150d112", 1, 3, 193, 150, 112, torch.bfloat16, False, "rand"),
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
    from fa_oracle import make_inputs  # seeded generator shared with the tests

    ref = _load_reference()
    for i, (name, B, H, Nq, Nkv, D, dtype, causal, dist) in enumerate(CASES):
        q, k, v = make_inputs(B, H, Nq, Nkv, D, dtype, seed=1000 + i, dist=dist)
        o_tiled = ref.FlashAttentionFunction.apply(q, k, v, None, causal)
        o_sdpa = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=causal)
        o_f32 = torch.nn.functional.scaled_dot_product_attention(
            q.float(), k.float(), v.float(), is_causal=causal)
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            q=_bits(q), k=_bits(k), v=_bits(v),
            o_ref_tiled=_bits(o_tiled), o_ref_sdpa=_bits(o_sdpa), o_ref_f32=o_f32.numpy(),
            dtype=np.array("float16" if dtype == torch.float16 else "bfloat16"),
            causal=np.array(causal), seed=np.array(1000 + i), dist=np.array(dist),
        )
        print(f"{name}: tiled-vs-f32 {(o_tiled.float() - o_f32).abs().max().item():.3e}  "
              f"sdpa-vs-f32 {(o_sdpa.float() - o_f32).abs().max().item():.3e}")


if __name__ == "__main__":
    main()
