#!/bin/bash
# default bench line (with the bounded CPU baseline) and the reference arm, as the driver launches them
mkdir -p gpurun_out
make -j8 >/dev/null 2>&1 || echo "MAKE FAILED"
export PYTHONUNBUFFERED=1
( time timeout 900 python bench.py 2> gpurun_out/bench_stderr.log | tee gpurun_out/bench.json ) 2>&1 | cut -c1-300
tail -3 gpurun_out/bench_stderr.log
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 3 2> gpurun_out/bench_ref_stderr.log | tee gpurun_out/bench_reference.json ) 2>&1 | cut -c1-900
tail -3 gpurun_out/bench_ref_stderr.log
