#!/bin/bash
# round 2, GPU call E: ncu captures (full set) of the dominant kernels + the launch list of the bench command
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fa_fwd_ws_kernel -s 2 -c 1 -f \
    -o gpurun_out/r02_prof_ws_n16384 python tools/prof_kernel.py 16384 x f16 4 ws > gpurun_out/r02_prof.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fa_fwd_sk_kernel -s 2 -c 1 -f \
    -o gpurun_out/r02_prof_sk_n8192 python tools/prof_kernel.py 8192 x f16 4 sk >> gpurun_out/r02_prof.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fa_fwd_ws_kernel -s 2 -c 1 -f \
    -o gpurun_out/r02_prof_ws_causal_n16384 python tools/prof_kernel.py 16384 causal f16 4 ws >> gpurun_out/r02_prof.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r02_bench_under_ncu.log 2>&1
tail -3 gpurun_out/r02_prof.log
ls -la gpurun_out/*.ncu-rep | tail -4
wc -l gpurun_out/r02_launches_bench.csv
