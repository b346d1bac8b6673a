#!/bin/bash
# A/B of the mask-free interior step in the persistent kernel (default) vs the build without it (nomaskpeel), then the GPU suite
mkdir -p gpurun_out; rm -f gpurun_out/ab_bench.json
AB_VARIANTS='default nomaskpeel default nomaskpeel' AB_NS='8192' bash tools/gpu_ab.sh
L=$PWD/flash-attention-v2-rdna3-minimal_b200/lib
for v in default nomaskpeel; do
  lib=$L/libfa_fwd_sm100.so; [ $v = nomaskpeel ] && lib=$L/libfa_fwd_sm100_nomaskpeel.so
  echo "== $v"
  FA_FWD_SM100_LIB=$lib timeout 300 python tools/sweep_kernels.py --batch 8 --ns 4096 --kernels auto --pdl 1 --out gpurun_out/sweep_c5_$v.json 2>&1 | tail -1
  FA_FWD_SM100_LIB=$lib timeout 300 python tools/sweep_kernels.py --dim 64 --ns 4096,8192 --kernels auto --pdl 1 --out gpurun_out/sweep_d64_$v.json 2>&1 | tail -2
  FA_FWD_SM100_LIB=$lib SD_KERNELS=auto timeout 300 python tools/bench_sd_shapes.py 2>&1 | cut -c1-75 | head -6
done
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2m_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log
tail -3 gpurun_out/r2m_pytest.log
