#!/bin/bash
# ncu: launch list of the bench command + one full capture of the dominant kernel
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fa_fwd_ws -s 1 -c 1 -f -o gpurun_out/prof_ws_n8192 \
    python tools/prof_kernel.py 8192 > gpurun_out/prof.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fa_fwd_ws -s 1 -c 1 -f -o gpurun_out/prof_ws_n16384 \
    python tools/prof_kernel.py 16384 >> gpurun_out/prof.log 2>&1
ls -la gpurun_out
