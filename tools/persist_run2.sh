#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== C3 seeds 0,1,3"
for o in "persist=0,qstart=0" "persist=0,qstart=1" "persist=1,qstart=1"; do
  NO_PASSES=1 SEEDS=0,1,3 OPTS=$o timeout 600 python tools/perf_probe.py
done
echo "== n=100k m=2M"
for o in "persist=0,qstart=1" "persist=1,qstart=1"; do
  NO_PASSES=1 N=100000 M=2000000 SEEDS=5 OPTS=$o timeout 300 python tools/perf_probe.py
done
echo "== n=100k m=1M"
for o in "persist=0,qstart=1" "persist=1,qstart=1"; do
  NO_PASSES=1 N=100000 M=1000000 SEEDS=2 OPTS=$o timeout 300 python tools/perf_probe.py
done
} > gpurun_out/p2_perf.log 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 ) > gpurun_out/p2_tests.log 2>&1
cat gpurun_out/p2_perf.log; tail -6 gpurun_out/p2_tests.log
