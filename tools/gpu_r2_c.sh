#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -4 gpurun_out/r2c_pytest.log
timeout 600 python bench.py > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2c_bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2c_bench_ref.json 2>> gpurun_out/r2c_bench.err
cut -c1-600 gpurun_out/r2c_bench_ref.json
