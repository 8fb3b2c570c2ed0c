#!/bin/bash
# Run each GPU diagnostic in its own process under a timeout so that a trapped / hung kernel
# cannot take the following checks (or the box) with it.  usage: run_diag.sh check1 check2 ...
mkdir -p gpurun_out
make -j8 >/dev/null 2>&1 || echo "MAKE FAILED"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee gpurun_out/diag.log
for c in "$@"; do
  echo "######## $c" | tee -a gpurun_out/diag.log
  timeout 300 python tests/gpu_diag.py "$c" 2>&1 | tail -60 | tee -a gpurun_out/diag.log
done
