#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
L=gpurun_out/r2_vt.log
: > $L
( timeout 300 python tests/gpu_diag.py gemm_epilogue 2>&1 | grep -E "FAIL|PASS|EXC|seg" | tail -20 ) | tee -a $L
( timeout 600 python -m pytest tests/test_unet_gpu.py tests/test_refunet_gpu.py tests/test_clip_gpu.py -q -x 2>&1 | tail -3 ) | tee -a $L
( timeout 300 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-reference-unet 2>> gpurun_out/r2_bench_stderr.log \
   | tee gpurun_out/r2_bench_vt.json | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1])
print('ms/step', round(d['ms_per_step'],2), 'clk', d['clocks']['sm_mhz'], 'gemm ms', round(d['kernels']['gemm_tc']['ms'],2), 'roofline', round(d['roofline']['frac'],3))
for s in d['top_shapes']:
    if ' T' in s['shape'] or 'attn' in s['shape']: print('   ', s)
" ) 2>&1 | tee -a $L
