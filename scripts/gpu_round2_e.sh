#!/bin/bash
# 1 GPU: PDL on the per-rank workload of config B at 8 GPUs (2 frames x 2 CFG branches = 4 images) and 4 GPUs (4 frames)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
L=gpurun_out/r2_e.log
: > $L
for fr in 2 4; do
for pdl in 1 0 1 0; do
  ( MDK_PDL=$pdl timeout 300 python bench.py --frames $fr --steps 20 --warmup 5 --skip-cpu-baseline --skip-reference-unet --skip-profile 2>> gpurun_out/r2_bench_stderr.log \
     | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1])
print('frames=$fr MDK_PDL=$pdl ms/step', round(d['ms_per_step'],3), 'clk', d['clocks']['sm_mhz'])
" ) 2>&1 | tee -a $L
done
done
