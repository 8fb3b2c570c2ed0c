#!/bin/bash
# N-GPU validation: NCCL parity (pytest: CFG split and plain frame sharding, even + uneven windows, all-gather mode),
# then the frame-sharded bench line with and without the CFG split.   usage: gpu_multi.sh N [extra bench args]
N=${1:-2}
shift
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
L=gpurun_out/r2_multi_n$N.log
: > $L
nvidia-smi --query-gpu=index,name --format=csv | head -12 | tee -a $L
for split in ${SPLITS:-1 0}; do
  echo "== tests/multigpu_check.py N=$N MDK_CFG_SPLIT=$split" | tee -a $L
  MDK_CFG_SPLIT=$split timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29510 + split)) \
      tests/multigpu_check.py > gpurun_out/r2_multigpu_check_n${N}_split$split.log 2>&1
  grep -E "rel_l2|MULTIGPU|Error|error|Traceback|File \"/root|assert" gpurun_out/r2_multigpu_check_n${N}_split$split.log | head -30 | cut -c1-400 | tee -a $L
done
for split in ${SPLITS:-1 0}; do
  echo "== bench N=$N MDK_CFG_SPLIT=$split" | tee -a $L
  ( MDK_CFG_SPLIT=$split timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520 + split)) \
      bench.py --gpus $N --steps 10 --warmup 3 --skip-cpu-baseline "$@" 2> gpurun_out/r2_bench_multi_stderr.log \
      | tee gpurun_out/r2_bench_n${N}_split$split.json | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1])
print('N=$N split=$split ms/step', round(d['ms_per_step'],2), 'frames/s', round(d['value'],3), 'clk', d.get('clocks',{}).get('sm_mhz'), 'parity', d.get('parity_vs_single'))
k=d.get('kernels')
if k: print('   kernels', {n:(round(v['ms'],2),v['n']) for n,v in k.items()})
" ) 2>&1 | tee -a $L
  grep -v "^W\|warn\|^$" gpurun_out/r2_bench_multi_stderr.log | tail -4 | cut -c1-300 | tee -a $L
done
if [ -n "$EXTRA_CONFIG" ]; then
  echo "== bench N=$N --config $EXTRA_CONFIG" | tee -a $L
  ( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29530 \
      bench.py --gpus $N --config $EXTRA_CONFIG --steps 5 --warmup 3 --skip-cpu-baseline 2> gpurun_out/r2_bench_multi_stderr.log \
      | tee gpurun_out/r2_bench_n${N}_config$EXTRA_CONFIG.json | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1])
print('N=$N config $EXTRA_CONFIG ms/step', round(d['ms_per_step'],2), 'frames/s', round(d['value'],3), 'windows', d['config']['windows'], 'roofline', round(d['roofline']['frac'],3), 'parity', d.get('parity_vs_single'))
" ) 2>&1 | tee -a $L
  grep -v "^W\|warn\|^$\|OMP\|^\*" gpurun_out/r2_bench_multi_stderr.log | tail -4 | cut -c1-300 | tee -a $L
fi
