#!/bin/bash
# A/B of the mbarrier try_wait time limit (MDK_WAIT_NS) on attention / GEMM / whole step, with the
# alternative attention kernels re-measured under it; then suite + bench.
mkdir -p gpurun_out
make -j8 >/dev/null 2>&1 || echo "MAKE FAILED"
export PYTHONUNBUFFERED=1
L=gpurun_out/ab.log
: > $L
echo "== kernels [default]" | tee -a $L
( timeout 600 python -m pytest tests/test_kernels_gpu.py -q 2>&1 | tail -15 ) | tee gpurun_out/k_default.log | tail -4 | tee -a $L
for cfg in "MDK_X=0" "MDK_WAIT_NS=0" "MDK_WAIT_NS=20000" "MDK_ATTN_SK=1" "MDK_ATTN_SK=1 MDK_WAIT_NS=0" "MDK_ATTN_PP=3" ; do
  echo "== perf_attn [$cfg]" | tee -a $L
  ( env $cfg timeout 200 python tests/gpu_diag.py perf_attn 2>&1 | grep -E "^perf" ) | tee -a $L
done
for cfg in "MDK_X=0" "MDK_WAIT_NS=0" ; do
  echo "== perf_gemm [$cfg]" | tee -a $L
  ( env $cfg timeout 200 python tests/gpu_diag.py perf_gemm 2>&1 | grep -E "^perf" ) | tee -a $L
done
echo "== full suite" | tee -a $L
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 ) | tee gpurun_out/pytest_gpu.log | tee -a $L
for cfg in "MDK_X=0" "MDK_WAIT_NS=0" ; do
  echo "== bench [$cfg]" | tee -a $L
  ( env $cfg timeout 900 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline 2> gpurun_out/bench_stderr.log | tee gpurun_out/bench_${cfg//[= ]/_}.json ) | cut -c1-230 | tee -a $L
done
cp gpurun_out/bench_MDK_X_0.json gpurun_out/bench.json
tail -3 gpurun_out/bench_stderr.log
