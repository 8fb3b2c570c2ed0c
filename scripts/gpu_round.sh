#!/bin/bash
# One GPU call: full GPU test suite, microbenchmarks, bench.py, ncu launch list + attention capture.
mkdir -p gpurun_out
make -j8 >/dev/null 2>&1 || echo "MAKE FAILED"
export PYTHONUNBUFFERED=1
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) | tee gpurun_out/pytest_gpu.log
( timeout 300 python tests/gpu_diag.py perf_gemm_small perf_gemm perf_attn 2>&1 | grep -E "perf|==" ) | tee gpurun_out/perf_micro.log
( timeout 900 python bench.py --steps 10 --warmup 3 2> gpurun_out/bench_stderr.log | tee gpurun_out/bench.json ) | cut -c1-700
tail -5 gpurun_out/bench_stderr.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-graph --skip-profile --skip-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; tail -2 gpurun_out/bench_under_ncu.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 1 -c 1 -o gpurun_out/prof_attn_v2 python tests/gpu_diag.py ncu_attn > gpurun_out/ncu_attn.log 2>&1; tail -2 gpurun_out/ncu_attn.log
ls -la gpurun_out
