#!/bin/bash
# 1 GPU: everything changed since the last green run — full -m gpu suite (incl. the config-B-shape UNet test with the
# torch-fp16 yardstick, exchange-layout kernels), smoke(), and both bench arms exactly as the driver calls them.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
L=gpurun_out/r2_b.log
: > $L
echo "== pytest -m gpu" | tee -a $L
( timeout 1200 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "unet parity|passed|failed|error|Error|FAIL|assert" | tail -30 ) | tee -a $L
echo "== smoke" | tee -a $L
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4 ) | tee -a $L
echo "== bench --impl reference (driver flags)" | tee -a $L
( /usr/bin/time -v timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2> gpurun_out/r2_ref_time.log | tee gpurun_out/r2_bench_reference.json | cut -c1-1200 ) | tee -a $L
grep -E "Elapsed|Maximum resident" gpurun_out/r2_ref_time.log | tee -a $L
echo "== bench (driver flags)" | tee -a $L
( timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 2> gpurun_out/r2_bench_stderr.log | tee gpurun_out/r2_bench_n1.json | cut -c1-1500 ) | tee -a $L
