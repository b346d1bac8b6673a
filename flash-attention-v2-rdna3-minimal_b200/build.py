"""Ahead-of-time build of the C-ABI library (``lib/libfa_fwd_sm100.so``) with nvcc for sm_100a.

Counterpart of the JIT build the reference does at import time
(/root/reference/rocwmma_fattn/FlashAttn.py:16-41, ``torch.utils.cpp_extension.load`` + hipify):
here the library has no torch / pybind dependency, so a single nvcc command is enough, and the
result ships in-tree to the GPU box.

Usage:  python build.py [--force] [--verbose]
"""
from __future__ import annotations

import argparse
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libfa_fwd_sm100.so")
STAMP_PATH = os.path.join(LIB_DIR, "libfa_fwd_sm100.stamp")
TRACE_LIB_PATH = os.path.join(LIB_DIR, "libfa_fwd_sm100_trace.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC or put /usr/local/cuda/bin on PATH)")


def have_nvcc() -> bool:
    try:
        _nvcc()
        return True
    except RuntimeError:
        return False


def source_files() -> list[str]:
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))]
    files.append(os.path.join(INCLUDE, "fa_fwd_sm100.h"))
    files.append(os.path.join(INCLUDE, "fa_fwd_sm100_test.h"))
    return files


def source_hash() -> str:
    h = hashlib.sha256()
    for f in source_files():
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    if not (os.path.exists(LIB_PATH) and os.path.exists(STAMP_PATH)):
        return False
    with open(STAMP_PATH) as fh:
        return fh.read().strip() == source_hash()


def _compile(out_path: str, extra: list[str], verbose: bool) -> None:
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [_nvcc(), *NVCC_FLAGS, *extra]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += ["-o", out_path, os.path.join(CSRC, "fa_capi.cu")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/fa_capi.cu (which includes every kernel) into LIB_PATH; returns the path."""
    if not force and is_current():
        return LIB_PATH
    _compile(LIB_PATH, [], verbose)
    with open(STAMP_PATH, "w") as fh:
        fh.write(source_hash())
    return LIB_PATH


def build_variant(name: str, defines: list[str], verbose: bool = False) -> str:
    """Experimental build lib/libfa_fwd_sm100_<name>.so with extra -D flags (A/B runs through
    FA_FWD_SM100_LIB; see tools/ab_bench.py).  Not used by the package."""
    out = os.path.join(LIB_DIR, f"libfa_fwd_sm100_{name}.so")
    _compile(out, [f"-D{d}" for d in defines], verbose)
    return out


def build_trace(verbose: bool = False) -> str:
    """Debug build with -DFA_TRACE (clock64 timeline of one CTA, tools/trace_ws.py).  Never loaded
    by the package unless FA_FWD_SM100_LIB points at it."""
    _compile(TRACE_LIB_PATH, ["-DFA_TRACE"], verbose)
    return TRACE_LIB_PATH


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--trace", action="store_true", help="also build the FA_TRACE debug library")
    ap.add_argument("--variant", action="append", default=[], metavar="NAME:DEF[,DEF...]",
                    help="also build lib/libfa_fwd_sm100_NAME.so with the given -D defines")
    a = ap.parse_args()
    print(build(force=a.force, verbose=a.verbose))
    if a.trace:
        print(build_trace(verbose=a.verbose))
    for v in a.variant:
        name, _, defs = v.partition(":")
        print(build_variant(name, [d for d in defs.split(",") if d], verbose=a.verbose))
