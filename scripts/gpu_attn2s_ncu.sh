#!/bin/bash
# parity of the two-stream kernel + ncu --set full of it and of the one-stream kernel (L0 self-attention, 2 images)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
L=gpurun_out/r2_attn2s_parity.log
: > $L
for env in "MDK_ATTN_2S=1" "MDK_ATTN_2S=1 MDK_ATTN_POLY=1"; do
  echo "== parity $env" | tee -a $L
  ( env $env timeout 200 python tests/gpu_diag.py attn 2>&1 | grep -E "FAIL|PASS|EXC|rel" | tail -40 ) | tee -a $L
done
for v in 1 0; do
  MDK_ATTN_2S=$v timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn -s 1 -c 1 \
      -f -o gpurun_out/r2_ncu_attn_2s$v python tests/gpu_diag.py ncu_attn > gpurun_out/r2_ncu_attn_2s$v.log 2>&1
  tail -3 gpurun_out/r2_ncu_attn_2s$v.log
done
