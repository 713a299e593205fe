#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== group mode, all widths"
TRACE_W=8 TRACE_K=4 REPS=5 SEEDS=0,1 timeout 900 python tools/multi_gpu_check.py group
} > gpurun_out/r7_mg.log 2>&1
grep -v "^\s*$" gpurun_out/r7_mg.log | tail -70
