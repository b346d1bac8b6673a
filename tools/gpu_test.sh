#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider ${PYTEST_K:+-k "$PYTEST_K"} > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest.log
grep -E "^(FAILED|ERROR)|AssertionError:|passed|failed" gpurun_out/pytest.log | head -60
if [ -n "$RUN_BWD_BENCH" ]; then timeout 600 python tools/bench_bwd.py; fi
