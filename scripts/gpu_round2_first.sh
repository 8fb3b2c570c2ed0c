#!/bin/bash
# First GPU call of the next round (one B200, ~8 min of box time): runs everything round 1 prepared but could
# not measure, cheapest first, each step under its own timeout so one hang cannot eat the call.
#   gpurun --timeout 420 -- 'bash scripts/gpu_round2_first.sh'
# Results land in gpurun_out/r2_first.log (+ JSON lines); PERF.md "Next experiments" says what each answers.
mkdir -p gpurun_out
make -j8 >/dev/null 2>&1 || echo "MAKE FAILED"
export PYTHONUNBUFFERED=1
L=gpurun_out/r2_first.log
: > $L
echo "== 1. parity of the kernels written without a GPU (split K / V^T rings)" | tee -a $L
( MDK_TEST_UNVALIDATED=1 timeout 120 python -m pytest tests/test_kernels_gpu.py -q -x -k "split_kv" 2>&1 | tail -6 ) | tee -a $L
echo "== 2. every attention switch on the L0 self-attention shape (incl. MDK_ATTN_SPLITKV)" | tee -a $L
( MDK_TEST_UNVALIDATED=1 timeout 60 python tests/gpu_diag.py ab_attn_switches 2>&1 | grep -E "^perf|PASS|FAIL|EXC" ) | tee -a $L
echo "== 2b. per-tile timeline of one CTA (which wait sets the period)" | tee -a $L
( MDK_TEST_UNVALIDATED=1 timeout 60 python tests/gpu_diag.py trace_attn 2>&1 | grep -E "trace|softmax warp|MMA warp|TMA warp|PASS|FAIL|EXC" ) | tee -a $L
echo "== 3. softmax inner-loop ceiling (pure instruction mix)" | tee -a $L
( timeout 20 ./build/softmax_loop_bench 2>&1 | tail -28 ) | tee -a $L
echo "== 4. step time with / without the split rings (only meaningful if 1. passed)" | tee -a $L
for cfg in "MDK_X=0" "MDK_ATTN_SPLITKV=1" ; do
  ( env $cfg timeout 150 python bench.py --steps 8 --warmup 3 --skip-cpu-baseline --skip-reference-unet 2>> gpurun_out/r2_bench_stderr.log \
      | tee gpurun_out/r2_bench_${cfg//[= ]/_}.json ) | cut -c1-220 | tee -a $L
done
echo "== 5. native CLIP image encoder (written without a GPU)" | tee -a $L
( MDK_TEST_UNVALIDATED=1 timeout 120 python -m pytest tests/test_clip_gpu.py -q 2>&1 | tail -6 ) | tee -a $L
echo "== 6. native VAE (written without a GPU)" | tee -a $L
( MDK_TEST_UNVALIDATED=1 timeout 180 python -m pytest tests/test_vae_gpu.py -q 2>&1 | tail -6 ) | tee -a $L
echo "== 7. the pipeline with every model stage native" | tee -a $L
( MDK_TEST_UNVALIDATED=1 timeout 180 python -m pytest tests/test_native_pipeline_gpu.py -q 2>&1 | tail -6 ) | tee -a $L
echo "== 8. VAE / CLIP timings at the bench resolution (only meaningful if 5 / 6 passed)" | tee -a $L
( timeout 120 python tests/gpu_diag.py perf_vae_clip 2>&1 | grep -E "^perf|PASS|FAIL|EXC" ) | tee -a $L
echo "== 9. (separate call, 2 GPUs)  gpurun --gpus 2 --timeout 600 -- 'MDK_CFG_SPLIT=1 bash scripts/gpu_multi.sh 2'" | tee -a $L
