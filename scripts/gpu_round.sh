#!/bin/bash
# One GPU call: attention v2 check + perf, GEMM epilogue triage, ncu capture of the GEMM, quick bench.
mkdir -p gpurun_out
make -j8 >/dev/null 2>&1 || echo "MAKE FAILED"
export PYTHONUNBUFFERED=1
( timeout 300 python tests/gpu_diag.py attn perf_attn 2>&1 | tail -30 ) | tee gpurun_out/attn_v2.log
for f in 0 1 3; do
  ( MDK_GEMM_DEBUG=$f timeout 200 python tests/gpu_diag.py perf_gemm_small 2>&1 | grep perf ) | tee -a gpurun_out/gemm_triage.log
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 2 -o gpurun_out/prof_gemm python tests/gpu_diag.py ncu_gemm > gpurun_out/ncu_gemm.log 2>&1; tail -3 gpurun_out/ncu_gemm.log
( timeout 900 python bench.py --steps 5 --warmup 3 --skip-cpu-baseline 2> gpurun_out/bench_stderr.log | tee gpurun_out/bench_v2.json ) | cut -c1-600
tail -5 gpurun_out/bench_stderr.log
ls -la gpurun_out
