#!/bin/bash
# round 2, GPU call A: full GPU test suite, then the per-N kernel A/B sweeps (fp16, bf16, causal) with cuDNN beside them
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
rm -f gpurun_out/sweep_kernels.json
timeout 600 python tools/sweep_kernels.py --lib > gpurun_out/r2a_sweep_f16.log 2>&1
timeout 300 python tools/sweep_kernels.py --lib --dtype bf16 --kernels auto,sk,ws2 --pdl 1 > gpurun_out/r2a_sweep_bf16.log 2>&1
timeout 300 python tools/sweep_kernels.py --lib --causal --kernels auto,ws,wide --pdl 1,0 > gpurun_out/r2a_sweep_causal.log 2>&1
tail -8 gpurun_out/r2a_sweep_f16.log
