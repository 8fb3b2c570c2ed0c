#!/bin/bash
# per-chunk P hand-off in the attention kernel: parity, perf (with / without polynomial exp2), suite, bench, ncu
mkdir -p gpurun_out
make -j8 >/dev/null 2>&1 || echo "MAKE FAILED"
export PYTHONUNBUFFERED=1
L=gpurun_out/ab.log
: > $L
echo "== kernels [default]" | tee -a $L
( timeout 600 python -m pytest tests/test_kernels_gpu.py -q 2>&1 | tail -25 ) | tee gpurun_out/k_default.log | tail -6 | tee -a $L
for cfg in "MDK_X=0" "MDK_ATTN_POLY=1" "MDK_ATTN_BKV=64" ; do
  echo "== perf_attn [$cfg]" | tee -a $L
  ( env $cfg timeout 200 python tests/gpu_diag.py perf_attn 2>&1 | grep -E "^perf" ) | tee -a $L
done
echo "== full suite" | tee -a $L
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 ) | tee gpurun_out/pytest_gpu.log | tee -a $L
for cfg in "MDK_X=0" "MDK_ATTN_POLY=1" ; do
  echo "== bench [$cfg]" | tee -a $L
  ( env $cfg timeout 900 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline 2> gpurun_out/bench_stderr.log | tee gpurun_out/bench_${cfg//[= ]/_}.json ) | cut -c1-230 | tee -a $L
done
cp gpurun_out/bench_MDK_X_0.json gpurun_out/bench.json
( timeout 300 ncu --set full --import-source on --clock-control none -k regex:attn_ -c 1 -o gpurun_out/prof_attn_r5 -f python tests/gpu_diag.py ncu_attn > gpurun_out/ncu_attn_r5.log 2>&1 ; tail -2 gpurun_out/ncu_attn_r5.log )
