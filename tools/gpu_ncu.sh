#!/bin/bash
# one full ncu capture (with SASS/source-level sampling) of the ws kernel at N=$1 (default 8192)
mkdir -p gpurun_out
N=${1:-8192}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fa_fwd_ws -s 2 -c 1 -f \
    -o gpurun_out/prof_ws_n$N python tools/prof_kernel.py $N x f16 4 > gpurun_out/prof.log 2>&1
tail -5 gpurun_out/prof.log
ls -la gpurun_out/*.ncu-rep
