#!/bin/bash
# first GPU contact: low-level probes, then the parity tests, then a short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
rm -f gpurun_out/debug.jsonl
timeout 600 python tools/gpu_debug.py > gpurun_out/debug.log 2>&1
echo "debug rc=$?" >> gpurun_out/debug.log
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?" >> gpurun_out/bench.err
tail -5 gpurun_out/pytest.log
cat gpurun_out/bench.json
