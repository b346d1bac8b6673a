"""Randomised parity over the kernels the library picks by itself (tools/fuzz_sk.py): non-causal problems made of whole
256-row query blocks - where the cost model may hand the call to the persistent stream-K kernel - against fp32 attention
on the device, every shape launched twice (same bits) and cross-checked against the one-shot kernel."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [11, 12])
def test_auto_selected_kernels_on_random_whole_block_shapes(seed):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_sk.py"), "60", str(seed)],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    last = res.stdout.strip().splitlines()[-1]
    assert last.startswith("ok: 60 shapes") and "'sk'" in last, last
