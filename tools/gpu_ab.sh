#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/ab_bench.py $AB_VARIANTS ${AB_NS:+-- $AB_NS}
