#!/bin/bash
# A/B of the shared-space row-max exchange in the early-S kernel (head dim 64): default build vs -DFA_MAX_XCHG_SHARED=0
mkdir -p gpurun_out
L=flash-attention-v2-rdna3-minimal_b200/lib
for rep in 1 2; do
  for v in default xs0; do
    lib=$L/libfa_fwd_sm100.so; [ $v = xs0 ] && lib=$L/libfa_fwd_sm100_xs0.so
    echo "== $v"
    FA_FWD_SM100_LIB=$PWD/$lib timeout 300 python tools/sweep_kernels.py --dim 64 --kernels ws3 --pdl 1 --ns 4096,16384 --out gpurun_out/sweep_ws3_$v.json 2>&1 | tail -3
  done
done
