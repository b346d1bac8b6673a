#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -3 gpurun_out/r2d_pytest.log
rm -f gpurun_out/sweep_kernels.json
timeout 300 python tools/sweep_kernels.py --lib --causal --kernels auto,ws,wide --pdl 1 2>&1 | tail -6
timeout 300 python tools/sweep_kernels.py --causal --dim 64 --kernels auto,ws,wide,ws3 --pdl 1 --ns 1024,2048,4096,8192,16384 2>&1 | tail -5
timeout 200 python tools/dtype_power_probe.py 2>&1 | tail -5
