#!/bin/bash
# Round of A/B measurements: split-key attention (MDK_ATTN_SK), 64-key x 3 CTA attention, GroupNorm L2
# chunking (MDK_GN_CHUNK_MB), lanes-per-row LayerNorm (MDK_LN_LPR), CG heuristics; then suite + bench + ncu.
mkdir -p gpurun_out
make -j8 >/dev/null 2>&1 || echo "MAKE FAILED"
export PYTHONUNBUFFERED=1
L=gpurun_out/ab.log
: > $L
run_k() {
  local label=$1; shift
  echo "== kernels [$label]" | tee -a $L
  ( env "$@" timeout 400 python -m pytest tests/test_kernels_gpu.py -q -x 2>&1 | tail -25 ) > gpurun_out/k_$label.log 2>&1
  tail -3 gpurun_out/k_$label.log | tee -a $L
  grep -q " passed" gpurun_out/k_$label.log && ! grep -q "failed\|error" gpurun_out/k_$label.log
}
OK_DEF=0; OK_SK0=0
run_k default MDK_X=0 && OK_DEF=1
run_k gnchunk1 MDK_GN_CHUNK_MB=1
if [ $OK_DEF = 0 ]; then
  grep -E "FAIL|rel_l2|Error|error|timed out" gpurun_out/k_default.log | head -20 | tee -a $L
  run_k sk0 MDK_ATTN_SK=0 && OK_SK0=1
fi
echo "OK_DEF=$OK_DEF OK_SK0=$OK_SK0" | tee -a $L
for cfg in "MDK_X=0" "MDK_ATTN_SK=0" "MDK_ATTN_SK=0 MDK_ATTN_BKV=64" ; do
  echo "== perf_attn [$cfg]" | tee -a $L
  ( env $cfg timeout 200 python tests/gpu_diag.py perf_attn 2>&1 | grep -E "^perf" ) | tee -a $L
done
for cfg in "MDK_X=0" "MDK_GN_CHUNK_MB=0 MDK_LN_LPR=0" "MDK_GN_CHUNK_MB=24" ; do
  echo "== perf_misc [$cfg]" | tee -a $L
  ( env $cfg timeout 200 python tests/gpu_diag.py perf_misc 2>&1 | grep -E "^perf (group|layer)" ) | tee -a $L
done
echo "== perf_gemm [default heuristics]" | tee -a $L
( timeout 200 python tests/gpu_diag.py perf_gemm 2>&1 | grep -E "^perf" ) | tee -a $L
CFG="MDK_X=0"
if [ $OK_DEF = 0 ]; then CFG="MDK_ATTN_SK=0"; fi
echo "== full suite + bench with [$CFG]" | tee -a $L
( env $CFG timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 ) | tee gpurun_out/pytest_gpu.log | tee -a $L
( env $CFG timeout 900 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline 2> gpurun_out/bench_stderr.log | tee gpurun_out/bench.json ) | cut -c1-260 | tee -a $L
tail -3 gpurun_out/bench_stderr.log
# source-level profile of the L0 self-attention kernel in the configuration that runs
( env $CFG timeout 300 ncu --set full --import-source on --clock-control none -k regex:attn_ -c 1 -o gpurun_out/prof_attn_r3 -f python tests/gpu_diag.py ncu_attn > gpurun_out/ncu_attn_r3.log 2>&1 ; tail -2 gpurun_out/ncu_attn_r3.log )
ls -la gpurun_out | tail -5
