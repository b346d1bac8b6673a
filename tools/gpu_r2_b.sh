#!/bin/bash
# round 2, GPU call B: tests, then the differential experiment for the persistent kernel's boundary cost
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -5 gpurun_out/r2b_pytest.log
rm -f gpurun_out/sweep_kernels.json
# whole units only (units = k x 148): no splits, k-1 unit boundaries per CTA
timeout 300 python tools/sweep_kernels.py --heads 37 --ns 1024,2048,4096 --kernels ws,sk --pdl 1 2>&1 | tail -3
timeout 300 python tools/sweep_kernels.py --heads 74 --ns 1024,2048 --kernels ws,sk --pdl 1 2>&1 | tail -2
# the sweep points
timeout 300 python tools/sweep_kernels.py --ns 1024,2048,4096,8192,16384 --kernels ws,sk --pdl 1 2>&1 | tail -5
