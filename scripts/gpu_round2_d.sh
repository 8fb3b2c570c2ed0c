#!/bin/bash
# 1 GPU: programmatic dependent launch — full GPU test suite with PDL on, then step time with MDK_PDL=1 / 0.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
L=gpurun_out/r2_d.log
: > $L
echo "== pytest -m gpu (PDL on)" | tee -a $L
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) | tee -a $L
for pdl in 1 0 1 0; do
  ( MDK_PDL=$pdl timeout 300 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline --skip-reference-unet --skip-profile 2>> gpurun_out/r2_bench_stderr.log \
     | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1])
print('MDK_PDL=$pdl ms/step', round(d['ms_per_step'],2), 'frames/s', round(d['value'],3), 'clk', d['clocks']['sm_mhz'], 'roofline', round(d['roofline']['frac'],3))
" ) 2>&1 | tee -a $L
done
for pdl in 1 0; do
  ( MDK_PDL=$pdl timeout 300 python bench.py --config A --steps 20 --warmup 3 --skip-cpu-baseline --skip-reference-unet --skip-profile 2>> gpurun_out/r2_bench_stderr.log \
     | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1])
print('config A (launch-bound) MDK_PDL=$pdl ms/step', round(d['ms_per_step'],3))
" ) 2>&1 | tee -a $L
done
