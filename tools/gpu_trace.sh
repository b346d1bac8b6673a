#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/trace_ws.py 16384 x 2 > gpurun_out/trace16k.log 2>&1
grep -A60 '^{' gpurun_out/trace16k.log | head -60
