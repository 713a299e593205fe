#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/p3_tests.log 2>&1
( MRC_B200_PERSIST=0 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/p3_tests_burst.log 2>&1
{
echo "== C3 seeds 0,1,3"
for o in "persist=0,walk_check=0" "persist=0,walk_check=1" "persist=1,walk_check=1"; do
  NO_PASSES=1 SEEDS=0,1,3 OPTS=$o timeout 600 python tools/perf_probe.py
done
echo "== n=100k m=2M"
for o in "persist=0,walk_check=1" "persist=1,walk_check=1"; do
  NO_PASSES=1 N=100000 M=2000000 SEEDS=5 OPTS=$o timeout 300 python tools/perf_probe.py
done
} > gpurun_out/p3_perf.log 2>&1
tail -6 gpurun_out/p3_tests.log; tail -6 gpurun_out/p3_tests_burst.log; cat gpurun_out/p3_perf.log
