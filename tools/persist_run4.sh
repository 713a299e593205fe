#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/dbg_sparse.py > gpurun_out/p4_dbg.log 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/p4_tests.log 2>&1
( MRC_B200_PERSIST=1 timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 ) > gpurun_out/p4_tests_p1.log 2>&1
cat gpurun_out/p4_dbg.log; tail -6 gpurun_out/p4_tests.log; tail -8 gpurun_out/p4_tests_p1.log
