#!/bin/bash
# first GPU run of the persistent probe kernel: quick correctness, then timing of build variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -12 ) > gpurun_out/p1_tests.log 2>&1
L=min_ratio_cycle_b200
{
NO_PASSES=1 OPTS=persist=0 timeout 300 python tools/perf_probe.py
timeout 300 python tools/perf_probe.py
for v in mb3 inl pf12 pf12l2 l2pf1 l2pf2; do
  MRC_B200_LIB=$PWD/$L/libmrc_b200.$v.so timeout 300 python tools/perf_probe.py
done
} > gpurun_out/p1_perf.log 2>&1
tail -5 gpurun_out/p1_tests.log
cat gpurun_out/p1_perf.log
