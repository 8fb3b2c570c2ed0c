#!/bin/bash
# B-stationary GEMM: parity (forced on all eligible shapes + the level-0 shapes), then perf of the L0 linears and step.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
L=gpurun_out/r2_bs.log
: > $L
( MDK_GEMM_BS=2 MDK_GEMM_CG=1 timeout 300 python tests/gpu_diag.py gemm_basic gemm_epilogue gemm_bs_l0 2>&1 | grep -E "FAIL|PASS|EXC|bs gemm|mdk" | tail -20 ) | tee -a $L
for bs in 1 0; do
  ( MDK_GEMM_BS=$bs timeout 200 python tests/gpu_diag.py perf_gemm_small 2>&1 | grep -E "^perf" | sed "s/^/[BS=$bs] /" | head -12 ) | tee -a $L
done
for bs in 1 0; do
  ( MDK_GEMM_BS=$bs timeout 300 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-reference-unet 2>> gpurun_out/r2_bench_stderr.log \
     | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1])
print('MDK_GEMM_BS=$bs ms/step', round(d['ms_per_step'],2), 'clk', d['clocks']['sm_mhz'], 'gemm ms', round(d['kernels']['gemm_tc']['ms'],2))
for s in d['top_shapes']:
    if 'K=320' in s['shape']: print('   ', s)
" ) 2>&1 | tee -a $L
done
