#!/bin/bash
# One GPU call: microbenchmarks with warm clocks, the first bench.py line, ncu captures.
mkdir -p gpurun_out
make -j8 >/dev/null 2>&1 || echo "MAKE FAILED"
export PYTHONUNBUFFERED=1
( timeout 300 python tests/gpu_diag.py perf_gemm perf_attn 2>&1 | tail -40 ) | tee gpurun_out/perf_micro.log
( timeout 900 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench_stderr.log | tee gpurun_out/bench_first.json ) | cut -c1-3000
tail -5 gpurun_out/bench_stderr.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 6 -c 2 -o gpurun_out/prof_gemm python tests/gpu_diag.py ncu_gemm > gpurun_out/ncu_gemm.log 2>&1; tail -3 gpurun_out/ncu_gemm.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 1 -c 1 -o gpurun_out/prof_attn python tests/gpu_diag.py ncu_attn > gpurun_out/ncu_attn.log 2>&1; tail -3 gpurun_out/ncu_attn.log
ls -la gpurun_out
