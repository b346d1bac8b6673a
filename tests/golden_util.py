"""Reader for tests/golden/*.npz (written by tests/golden/make_golden.py)."""
import glob
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def golden_names():
    """Forward fixtures (tests/golden/make_golden.py)."""
    return [n for n in _names() if not n.startswith("bwd_")]


def golden_bwd_names():
    """Backward fixtures (tests/golden/make_golden_bwd.py)."""
    return [n for n in _names() if n.startswith("bwd_")]


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


def load_golden_bwd(name: str) -> dict:
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    dtype = torch.float16 if str(z["dtype"]) == "float16" else torch.bfloat16
    out = {k: _from_bits(z[k], dtype) for k in ("q", "k", "v", "d_o", "o_ref_tiled", "dq_ref_tiled",
                                                "dk_ref_tiled", "dv_ref_tiled", "dq_sdpa16", "dk_sdpa16",
                                                "dv_sdpa16")}
    for k in ("dq_f32", "dk_f32", "dv_f32"):
        out[k] = torch.from_numpy(z[k].copy())
    out["dtype"] = dtype
    out["causal"] = bool(z["causal"])
    out["seed"] = int(z["seed"])
    return out
