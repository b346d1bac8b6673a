#!/bin/bash
# ncu: launch list of the bench command + one full capture of the dominant kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fa_fwd_ws -s 1 -c 1 -f -o gpurun_out/prof_ws_n16384 \
    python tools/prof_kernel.py 16384 > gpurun_out/prof.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fa_fwd_ws -s 1 -c 1 -f -o gpurun_out/prof_ws_n4096 \
    python tools/prof_kernel.py 4096 >> gpurun_out/prof.log 2>&1
tail -3 gpurun_out/pytest.log
cat gpurun_out/bench.json
ls -la gpurun_out
