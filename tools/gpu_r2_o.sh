#!/bin/bash
# A/B: have_o a compile-time constant in the non-first softmax steps + condition-free MMA loop (default) vs the committed tree (prev)
mkdir -p gpurun_out; rm -f gpurun_out/ab_bench.json
AB_VARIANTS='default prev default prev' AB_NS='1024 2048 4096 8192 16384' bash tools/gpu_ab.sh
FA_CAUSAL=1 AB_VARIANTS='default prev' AB_NS='4096 16384' bash tools/gpu_ab.sh
L=$PWD/flash-attention-v2-rdna3-minimal_b200/lib
for v in default prev; do
  lib=$L/libfa_fwd_sm100.so; [ $v = prev ] && lib=$L/libfa_fwd_sm100_prev.so
  echo "== $v"
  FA_FWD_SM100_LIB=$lib timeout 300 python tools/sweep_kernels.py --batch 8 --ns 4096 --kernels auto --pdl 1 --out gpurun_out/sweep_c5_$v.json 2>&1 | tail -1
  FA_FWD_SM100_LIB=$lib SD_KERNELS=auto timeout 300 python tools/bench_sd_shapes.py 2>&1 | cut -c1-75 | head -7
done
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2o_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2o_pytest.log
tail -3 gpurun_out/r2o_pytest.log
