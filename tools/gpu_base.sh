#!/bin/bash
# baseline of HEAD: parity tests, A/B of exp2-emulation variants, timeline traces, short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -3 gpurun_out/pytest.log
timeout 900 python tools/ab_bench.py $AB_VARIANTS
timeout 300 python tools/trace_ws.py 16384 > gpurun_out/trace16k.log 2>&1
tail -26 gpurun_out/trace16k.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json
