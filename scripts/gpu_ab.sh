#!/bin/bash
# A/B validation of the CTA-pair GEMM (MDK_GEMM_CG) and the ping-pong attention (MDK_ATTN_PP):
# kernel parity for each switch setting, micro-benchmarks, then the full test suite + bench on the
# best passing configuration.
mkdir -p gpurun_out
make -j8 >/dev/null 2>&1 || echo "MAKE FAILED"
export PYTHONUNBUFFERED=1
L=gpurun_out/ab.log
: > $L
run_k() {  # $1 = label, rest = env assignments
  local label=$1; shift
  echo "== kernels [$label]" | tee -a $L
  ( env "$@" timeout 400 python -m pytest tests/test_kernels_gpu.py -q -x 2>&1 | tail -15 ) > gpurun_out/k_$label.log 2>&1
  tail -3 gpurun_out/k_$label.log | tee -a $L
  grep -q " passed" gpurun_out/k_$label.log && ! grep -q "failed\|error" gpurun_out/k_$label.log
}
OK_DEF=0; OK_CG1=0; OK_PP0=0
run_k default MDK_X=0 && OK_DEF=1
if [ $OK_DEF = 0 ]; then
  run_k cg1 MDK_GEMM_CG=1 && OK_CG1=1
  run_k pp0 MDK_ATTN_PP=0 && OK_PP0=1
  grep -E "FAIL|rel_l2|Error|error|timed out" gpurun_out/k_default.log | head -20 | tee -a $L
fi
echo "OK_DEF=$OK_DEF OK_CG1=$OK_CG1 OK_PP0=$OK_PP0" | tee -a $L
for cfg in "MDK_X=0" "MDK_GEMM_CG=1" ; do
  echo "== perf_gemm [$cfg]" | tee -a $L
  ( env $cfg timeout 200 python tests/gpu_diag.py perf_gemm perf_gemm_small 2>&1 | grep -E "^perf" ) | tee -a $L
done
for cfg in "MDK_X=0" "MDK_ATTN_PP=0" "MDK_ATTN_PP=1" ; do
  echo "== perf_attn [$cfg]" | tee -a $L
  ( env $cfg timeout 200 python tests/gpu_diag.py perf_attn 2>&1 | grep -E "^perf" ) | tee -a $L
done
# full suite + bench on the configuration that passed
CFG="MDK_X=0"
if [ $OK_DEF = 0 ]; then
  CFG=""
  [ $OK_CG1 = 1 ] && CFG="MDK_GEMM_CG=1"
  [ $OK_PP0 = 1 ] && CFG="MDK_ATTN_PP=0"
  [ -z "$CFG" ] && CFG="MDK_GEMM_CG=1 MDK_ATTN_PP=0"
fi
echo "== full suite + bench with [$CFG]" | tee -a $L
( env $CFG timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 ) | tee gpurun_out/pytest_gpu.log | tee -a $L
( env $CFG timeout 900 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline 2> gpurun_out/bench_stderr.log | tee gpurun_out/bench.json ) | cut -c1-260 | tee -a $L
tail -3 gpurun_out/bench_stderr.log
