#!/bin/bash
mkdir -p gpurun_out
make -j8 >/dev/null 2>&1 || echo "MAKE FAILED"
export PYTHONUNBUFFERED=1
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 ) | tee gpurun_out/pytest_gpu.log
( timeout 300 python tests/gpu_diag.py perf_attn perf_gemm_small 2>&1 | grep -E "perf|==" ) | tee gpurun_out/perf_micro.log
( timeout 900 python bench.py --steps 10 --warmup 3 --skip-cpu-baseline 2> gpurun_out/bench_stderr.log | tee gpurun_out/bench.json ) | cut -c1-300
tail -5 gpurun_out/bench_stderr.log
ls -la gpurun_out
