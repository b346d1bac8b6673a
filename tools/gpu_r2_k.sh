#!/bin/bash
# round 2, GPU call K: early S issue in the persistent kernel at head dim 64 - parity, sanitizer, D=64 sweep vs the early-S kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2k_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
tail -3 gpurun_out/r2k_pytest.log
timeout 400 python tools/sweep_kernels.py --dim 64 --kernels auto,ws3,sk --pdl 1 --ns 2048,4096,8192,16384 --lib --out gpurun_out/sweep_d64_sk_early_s.json 2>&1 | tail -5
timeout 400 python tools/sweep_kernels.py --dim 64 --dtype bf16 --kernels auto,ws3,sk --pdl 1 --ns 4096,8192 --out gpurun_out/sweep_d64_sk_early_s_bf16.json 2>&1 | tail -3
bash tools/gpu_sanitize_sk.sh 2>&1 | grep -E "^==|SUMMARY|^ok"
grep "Race reported" -A1 gpurun_out/sanitizer_sk_racecheck.log | grep -o "[a-z_0-9]*.cuh:[0-9]*" | sort | uniq -c
