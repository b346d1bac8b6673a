"""The drop-in boundary from plain C: tests/c_abi/consumer.c includes only include/fa_fwd_sm100.h and
links only the library (no CUDA headers, no torch).  CPU part: it compiles with -Wall -Wextra -Werror
as C11, links, and its --abi mode agrees with the header; GPU part: it runs three forwards through
fa_fwd_sm100_host() and checks them against its own double-precision attention."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402


def _binary():
    entry.build()
    return entry.build_c_consumer()


def test_c_consumer_compiles_links_and_reports_the_abi():
    out = subprocess.run([_binary(), "--abi"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    fields = out.stdout.split()
    from rocwmma_fattn import _capi

    assert fields[0] == "abi" and int(fields[1]) == _capi.FA_ABI_VERSION
    assert int(fields[3]) == _capi.FA_KERNEL_WS  # the N=16384 sweep point, selected without a GPU


@pytest.mark.gpu
def test_c_consumer_runs_the_forward_and_matches_its_own_reference():
    out = subprocess.run([_binary()], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.strip().endswith("OK")
    assert out.stdout.count("max|o-ref|") == 3
