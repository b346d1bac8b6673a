#!/bin/bash
# round 2, GPU call J (third session): ncu captures (full set) on the final tree - the dominant kernel, the persistent kernel at
# the N=8192 sweep point and at an SDXL shape through the hook - + the launch list of the bench command
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -s 2 -c 1 -f"
timeout 600 $NCU -k regex:fa_fwd_ws_kernel -o gpurun_out/r02s3_prof_ws_n16384 python tools/prof_kernel.py 16384 x f16 4 ws > gpurun_out/r02s3_prof.log 2>&1
timeout 600 $NCU -k regex:fa_fwd_sk_kernel -o gpurun_out/r02s3_prof_sk_n8192 python tools/prof_kernel.py 8192 x f16 4 sk >> gpurun_out/r02s3_prof.log 2>&1
timeout 600 $NCU -k regex:fa_fwd_sk_kernel -o gpurun_out/r02s3_prof_sk_sdxl python tools/prof_sd_shape.py 2 4096 4096 10 64 4 >> gpurun_out/r02s3_prof.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r02s3_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/r02s3_bench_under_ncu.log 2>&1
tail -3 gpurun_out/r02s3_prof.log
for n in ws_n16384 sk_n8192 sk_sdxl; do python tools/ncu_summary.py gpurun_out/r02s3_prof_$n.ncu-rep gpurun_out/r02s3_ncu_full_$n.json > /dev/null 2>&1; done
rm -f gpurun_out/r02s3_prof_sk_n8192.ncu-rep gpurun_out/r02s3_prof_sk_sdxl.ncu-rep  # (64 MiB limit on what comes back; the ws capture is kept)
ls -la gpurun_out/r02s3_* | tail -8
wc -l gpurun_out/r02s3_launches_bench.csv
