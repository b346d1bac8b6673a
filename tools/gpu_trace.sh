#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/trace_ws.py 16384 > gpurun_out/trace16k.log 2>&1
timeout 300 python tools/trace_ws.py 4096 > gpurun_out/trace4k.log 2>&1
nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap --format=csv,noheader -lms 50 > gpurun_out/clocks_loop.csv &
SMI=$!
timeout 300 python tools/prof_kernel.py 16384 x f16 1500 > gpurun_out/loop.log 2>&1
kill $SMI
tail -40 gpurun_out/trace16k.log
sort gpurun_out/clocks_loop.csv | uniq -c | sort -rn | head
