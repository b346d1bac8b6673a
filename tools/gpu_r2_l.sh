#!/bin/bash
# A/B: first KV tile of a pass peeled into its own softmax-step instantiation (default) vs -DFA_PEEL_MASK=0; then the whole GPU suite
mkdir -p gpurun_out; rm -f gpurun_out/ab_bench.json
AB_VARIANTS='default nomaskpeel default nomaskpeel' AB_NS='1024 2048 4096 8192 16384' bash tools/gpu_ab.sh
L=$PWD/flash-attention-v2-rdna3-minimal_b200/lib
for v in default nomaskpeel; do
  lib=$L/libfa_fwd_sm100.so; [ $v = nomaskpeel ] && lib=$L/libfa_fwd_sm100_nomaskpeel.so
  echo "== $v"; FA_FWD_SM100_LIB=$lib SD_KERNELS=auto timeout 300 python tools/bench_sd_shapes.py 2>&1 | cut -c1-75
done
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/r2l_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log
tail -3 gpurun_out/r2l_pytest.log
