"""Reader for tests/golden/*.npz (written by tests/golden/make_golden.py)."""
import glob
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def _from_bits(a: np.ndarray, dtype: torch.dtype) -> torch.Tensor:
    return torch.from_numpy(a.view(np.int16).copy()).view(dtype)


def load_golden(name: str) -> dict:
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    dtype = torch.float16 if str(z["dtype"]) == "float16" else torch.bfloat16
    out = {k: _from_bits(z[k], dtype) for k in ("q", "k", "v", "o_ref_tiled", "o_ref_sdpa")}
    out["o_ref_f32"] = torch.from_numpy(z["o_ref_f32"].copy())
    out["dtype"] = dtype
    out["causal"] = bool(z["causal"])
    out["seed"] = int(z["seed"])
    out["dist"] = str(z["dist"])
    return out
