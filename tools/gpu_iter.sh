#!/bin/bash
# quick iteration: parity tests, A/B variants, timeline trace (cold and under sustained load)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -3 gpurun_out/pytest.log
timeout 900 python tools/ab_bench.py $AB_VARIANTS
timeout 300 python tools/trace_ws.py 16384 x 2 > gpurun_out/trace16k.log 2>&1
tail -34 gpurun_out/trace16k.log
timeout 300 python tools/trace_ws.py 16384 x 600 > gpurun_out/trace16k_hot.log 2>&1
tail -5 gpurun_out/trace16k_hot.log
