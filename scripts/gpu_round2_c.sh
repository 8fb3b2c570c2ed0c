#!/bin/bash
# 1 GPU: GEGLU gelu rewrite, GroupNorm apply prologue, pipelined temporal attention — kernel parity, then step time.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
L=gpurun_out/r2_c.log
: > $L
( timeout 600 python tests/gpu_diag.py gemm_epilogue norms temporal 2>&1 | grep -E "FAIL|PASS|EXC|geglu|exchange" | tail -30 ) | tee -a $L
( timeout 300 python tests/gpu_diag.py perf_misc 2>&1 | grep -E "^perf|PASS|FAIL|EXC" | tail -30 ) | tee -a $L
( MDK_TATTN_PIPE=0 timeout 300 python tests/gpu_diag.py perf_misc 2>&1 | grep -E "^perf.*temporal" | sed 's/^/[pipe off] /' | tail -10 ) | tee -a $L
bash scripts/gpu_bench_ab.sh 2>&1 | grep -v "^   kernels" | tee -a $L
python - <<'PY' | tee -a $L
import json
for n in ("base", "a2s"):
    d = json.load(open(f"gpurun_out/r2_bench_{n}.json"))
    print(n, {k: (round(v["ms"], 2), v["n"]) for k, v in d["kernels"].items()})
    for s in d["top_shapes"][:12]:
        print("   ", s)
PY
