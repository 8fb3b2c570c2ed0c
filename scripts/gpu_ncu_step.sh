#!/bin/bash
# ncu launch list (gpu__time_duration + dram bytes + tensor pipe) of ONE eager step: config B (16 frames) and the
# per-rank workload of 8 GPUs (2 frames), summarised per kernel; plus one --set full capture of the top kernel.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for fr in 16 2; do
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
      --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_f$fr.csv \
      python bench.py --ncu-step --frames $fr > gpurun_out/r2_ncu_step_f$fr.log 2>&1
  python scripts/ncu_summarise.py gpurun_out/r2_launches_f$fr.csv gpurun_out/r2_ncu_step_summary_f$fr.txt gpurun_out/r2_ncu_step_traffic_f$fr.json
  head -30 gpurun_out/r2_ncu_step_summary_f$fr.txt
done
rm -f gpurun_out/r2_launches_f2.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_2s -s 1 -c 1 -f -o gpurun_out/r2_ncu_attn2s_final \
    python tests/gpu_diag.py ncu_attn > gpurun_out/r2_ncu_attn2s_final.log 2>&1
tail -2 gpurun_out/r2_ncu_attn2s_final.log
