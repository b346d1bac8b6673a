#!/bin/bash
# bench + ncu evidence: launch list of the bench command (our kernels only) and one full capture of
# the dominant kernel (N=16384) with source-level sampling
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fa_fwd -c 60 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:fa_fwd_(ws|sk)" -s 2 -c 1 -f \
    -o gpurun_out/prof_ws_n16384 python tools/prof_kernel.py 16384 x f16 4 > gpurun_out/prof.log 2>&1
tail -2 gpurun_out/bench.err
cat gpurun_out/bench.json
ls -la gpurun_out
