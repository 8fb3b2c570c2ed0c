#!/bin/bash
# Two-stream attention kernels: parity of every attention case, A/B timing on the L0 shape.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
L=gpurun_out/r2_attn2s.log
: > $L
for env in "MDK_ATTN_2S=3" "MDK_ATTN_2S=3 MDK_ATTN_POLY=0" "MDK_ATTN_2S=3 MDK_ATTN_POLY=2"; do
  echo "== parity $env" | tee -a $L
  ( env $env timeout 200 python tests/gpu_diag.py attn 2>&1 | grep -E "FAIL|PASS|EXC|mdk" | tail -30 ) | tee -a $L
done
echo "== A/B on the L0 self-attention shape" | tee -a $L
( MDK_TEST_UNVALIDATED=1 timeout 120 python tests/gpu_diag.py ab_attn_switches 2>&1 | grep -E "^perf|PASS|FAIL|EXC" ) | tee -a $L
