#!/bin/bash
# A/B of the peeled loop in the CTA-pair wide kernel (head dims 193..256; default) vs the committed tree (prev), then its tests
mkdir -p gpurun_out
L=$PWD/flash-attention-v2-rdna3-minimal_b200/lib
for rep in 1 2; do
for v in default prev; do
  lib=$L/libfa_fwd_sm100.so; [ $v = prev ] && lib=$L/libfa_fwd_sm100_prev.so
  echo "== $v"; FA_FWD_SM100_LIB=$lib BENCH_D=256,224 BENCH_N=4096,16384 timeout 300 python tools/bench_wide.py 2>&1 | grep -v "^$" | tail -10
done
done
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider -k "wide or head_dim or pairs or 256" 2>&1 | tail -2
