#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -k "$PYTEST_K" > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest.log
grep -E "^(FAILED|ERROR|[0-9]+ (passed|failed))|AssertionError:|passed|failed" gpurun_out/pytest.log | head -80
