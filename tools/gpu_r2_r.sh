#!/bin/bash
# A/B of the backward's mask-free interior iterations (default) vs -DFA_BWD_PEEL_MASK=0 (bwdgen), then the backward GPU tests
mkdir -p gpurun_out
L=$PWD/flash-attention-v2-rdna3-minimal_b200/lib
for rep in 1 2; do
for v in default bwdgen; do
  lib=$L/libfa_fwd_sm100.so; [ $v = bwdgen ] && lib=$L/libfa_fwd_sm100_bwdgen.so
  echo "== $v"; FA_FWD_SM100_LIB=$lib FA_BWD_QUICK=1 timeout 300 python tools/bench_bwd.py 2048 4096 16384 2>&1 | tail -4
done
done
timeout 1500 python -m pytest tests/test_gpu_backward.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -2
