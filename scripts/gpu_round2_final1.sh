#!/bin/bash
# 1 GPU, end-of-round validation: pytest -m gpu, smoke(), both bench arms exactly as the driver calls them, config E.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
L=gpurun_out/r2_final1.log
: > $L
echo "== pytest -m gpu" | tee -a $L
( timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "unet parity|passed|failed|error|Error|FAIL" | tail -12 ) | tee -a $L
echo "== smoke" | tee -a $L
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 ) | tee -a $L
echo "== bench --impl reference --gpus 1 --steps 20 --warmup 5" | tee -a $L
SECONDS=0
( timeout 1200 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2> gpurun_out/r2_ref_stderr.log | tee gpurun_out/r2_bench_reference.json | cut -c1-1500 ) | tee -a $L
echo "reference arm wall: ${SECONDS}s" | tee -a $L
echo "== bench --gpus 1 --steps 20 --warmup 5" | tee -a $L
( timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 2> gpurun_out/r2_bench_stderr.log | tee gpurun_out/r2_bench_n1.json | cut -c1-1800 ) | tee -a $L
echo "== bench --config E" | tee -a $L
( timeout 900 python bench.py --config E --steps 5 --warmup 3 --skip-cpu-baseline --skip-reference-unet 2>> gpurun_out/r2_bench_stderr.log | tee gpurun_out/r2_bench_configE.json | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1])
print('config E ms/step', round(d['ms_per_step'],2), 'frames/s', round(d['value'],3), 'roofline', round(d['roofline']['frac'],3), 'clk', d['clocks']['sm_mhz'])
for s in d['top_shapes'][:6]: print('   ', s)
" ) 2>&1 | tee -a $L
