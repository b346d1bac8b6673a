#!/bin/bash
# round 2, GPU call F: the ComfyUI / sd-webui hooks (parity at Stable-Diffusion shapes) + the attention stack of one UNet step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hooks.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_hooks.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_hooks.log
grep -E "^(FAILED|ERROR)|Error|passed|failed" gpurun_out/pytest_hooks.log | head -40
timeout 600 python tools/bench_sd_unet.py --json gpurun_out/bench_sd_unet.json 2>&1 | tail -8
timeout 300 python tools/bench_sd_shapes.py 2>&1 | tee gpurun_out/bench_sd_shapes.txt | tail -10
