#!/bin/bash
# Round-end evidence run on one B200: full GPU test suite, smoke(), the default bench line (with the CPU
# baseline), the reference arm, the ncu launch list of one step and full ncu captures of the top kernels.
mkdir -p gpurun_out
make -j8 >/dev/null 2>&1 || echo "MAKE FAILED"
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv | tee gpurun_out/final_gpu.log
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 ) | tee gpurun_out/pytest_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -5 ) | tee gpurun_out/smoke.log
( timeout 900 python bench.py 2> gpurun_out/bench_stderr.log | tee gpurun_out/bench.json ) | cut -c1-300
tail -3 gpurun_out/bench_stderr.log
( timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2> gpurun_out/bench_ref_stderr.log | tee gpurun_out/bench_reference.json ) | cut -c1-400
( timeout 900 ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/launches_step.csv \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
    python bench.py --ncu-step --skip-cpu-baseline --skip-profile > gpurun_out/ncu_step.log 2>&1 ; tail -2 gpurun_out/ncu_step.log )
python scripts/ncu_summarise.py gpurun_out/launches_step.csv gpurun_out/ncu_step_summary.txt gpurun_out/ncu_step_traffic.json && head -24 gpurun_out/ncu_step_summary.txt
( timeout 300 ncu --set full --import-source on --clock-control none -k regex:attn_tc -c 1 -o gpurun_out/prof_attn_final -f python tests/gpu_diag.py ncu_attn > gpurun_out/ncu_attn_final.log 2>&1 ; tail -1 gpurun_out/ncu_attn_final.log )
( timeout 300 ncu --set full --import-source on --clock-control none -k regex:gemm_tc -s 4 -c 2 -o gpurun_out/prof_gemm_final -f python tests/gpu_diag.py ncu_gemm > gpurun_out/ncu_gemm_final.log 2>&1 ; tail -1 gpurun_out/ncu_gemm_final.log )
ls -la gpurun_out | tail -12
