#!/bin/bash
# step time A/B: bench.py under different env switch sets (each line: "NAME ENV=.. ENV=..")
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
L=gpurun_out/r2_bench_ab.log
: > $L
while read -r name envs; do
  [ -z "$name" ] && continue
  ( env $envs timeout 200 python bench.py --steps 6 --warmup 3 --skip-cpu-baseline --skip-reference-unet 2>> gpurun_out/r2_bench_stderr.log \
      | tee gpurun_out/r2_bench_$name.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('$name', 'ms/step', round(d['ms_per_step'],2), 'frames/s', round(d['value'],3), 'clk', d.get('clocks',{}).get('sm_mhz'))
ts=d.get('top_shapes') or d.get('profile',{}).get('top_shapes')
k=d.get('kernels') or d.get('profile',{}).get('kernels')
if k: print('   kernels', json.dumps(k)[:900])
" ) 2>&1 | tee -a $L
done <<'EOT'
base MDK_X=0
a2s MDK_ATTN_2S=1 MDK_ATTN_POLY=1
EOT
