#!/bin/bash
# A/B of the peeled loop in the early-S kernel on pairs (default) vs the committed tree (prev), then the GPU suite
mkdir -p gpurun_out
L=$PWD/flash-attention-v2-rdna3-minimal_b200/lib
for rep in 1 2; do
for v in default prev; do
  lib=$L/libfa_fwd_sm100.so; [ $v = prev ] && lib=$L/libfa_fwd_sm100_prev.so
  echo "== $v"
  FA_FWD_SM100_LIB=$lib timeout 300 python tools/sweep_kernels.py --dim 64 --ns 2048,4096,16384 --kernels ws3 --pdl 1 --out gpurun_out/sweep_ws3_$v.json 2>&1 | tail -3
  FA_FWD_SM100_LIB=$lib SD_SHAPES="2,8,16384,16384,40" SD_KERNELS=auto timeout 300 python tools/bench_sd_shapes.py 2>&1 | cut -c1-75
done
done
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2n_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2n_pytest.log
tail -3 gpurun_out/r2n_pytest.log
