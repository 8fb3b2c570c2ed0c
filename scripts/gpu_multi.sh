#!/bin/bash
# N-GPU validation: sharded-vs-single parity, then the frame-sharded bench line.   usage: gpu_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
make -j8 >/dev/null 2>&1 || echo "MAKE FAILED"
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=index,name --format=csv | head -12
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py > gpurun_out/multigpu_check_$N.log 2>&1
grep -E "rel_l2|MULTIGPU|ran |Error|error|Traceback" gpurun_out/multigpu_check_$N.log | head -30
( timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --skip-cpu-baseline 2> gpurun_out/bench_multi_stderr.log | tee gpurun_out/bench_n$N.json ) | cut -c1-300
( MDK_SHARD_MODE=allgather timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --skip-cpu-baseline --skip-profile 2>> gpurun_out/bench_multi_stderr.log | tee gpurun_out/bench_n${N}_allgather.json ) | cut -c1-300
grep -v "^W\|warn" gpurun_out/bench_multi_stderr.log | tail -6 | cut -c1-300
