#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"layernorm|gn_" -c 12 -f -o gpurun_out/r2_ncu_norms \
    python tests/gpu_diag.py perf_misc > gpurun_out/r2_ncu_norms.log 2>&1
tail -2 gpurun_out/r2_ncu_norms.log
ncu -i gpurun_out/r2_ncu_norms.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr=rows[0]
want=['Kernel Name','gpu__time_duration.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__block_size','dram__bytes_read.sum','dram__bytes_write.sum','smsp__issue_active.avg.pct','launch__occupancy_limit_registers','launch__waves_per_multiprocessor','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio']
ix=[i for i,h in enumerate(hdr) if any(h==w or h.startswith(w+' ') for w in want)]
for r in rows[2:14]:
    print(' | '.join(f'{hdr[i][:38]}={r[i]}' for i in ix))
" | tee gpurun_out/r2_ncu_norms_metrics.txt | cut -c1-900
