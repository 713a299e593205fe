#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_persist.py tests/test_gpu_stress.py -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r5_tests.log 2>&1
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -12 ) > gpurun_out/r5_smoke.log 2>&1
( timeout 900 python bench.py --steps 20 --warmup 3 2>&1 | tail -3 ) > gpurun_out/r5_bench.log 2>&1
tail -5 gpurun_out/r5_tests.log; cat gpurun_out/r5_smoke.log; cat gpurun_out/r5_bench.log
