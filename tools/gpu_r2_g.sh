#!/bin/bash
# round 2, GPU call G: new kernel selection for Stable-Diffusion shapes (persistent kernel at head dims <= 64)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_hooks.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider \
    -k "hooks or stable_diffusion or stream_k or comfy or webui or sdpa_front" > gpurun_out/pytest_sd.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_sd.log
grep -E "^(FAILED|ERROR)|Error|passed|failed" gpurun_out/pytest_sd.log | head -40
timeout 600 python tools/bench_sd_unet.py --json gpurun_out/bench_sd_unet.json 2>&1 | tail -8
SD_KERNELS=auto,ws,ws3,sk,wide timeout 400 python tools/bench_sd_shapes.py 2>&1 | tee gpurun_out/bench_sd_shapes_kernels.txt
