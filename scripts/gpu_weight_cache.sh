#!/bin/bash
# packed weight cache on the device: the GPU test + start-up timings at SD-1.5 size
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
L=gpurun_out/r2_weight_cache.log
: > $L
timeout 300 python -m pytest tests/test_unet_gpu.py -m gpu -x -q -k "packed_checkpoint" 2>&1 | tail -5 | tee -a $L
timeout 400 python tests/gpu_diag.py time_weight_cache 2>&1 | grep -v "^W\|Warning" | tail -6 | cut -c1-900 | tee -a $L
