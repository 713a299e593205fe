#!/bin/bash
cd "$(dirname "$0")/.."
NO_PASSES=1 SEEDS=0,1 MODES=1:0,1:1,4:0 timeout 600 python tools/perf_probe.py 2>&1 | cut -c1-250
MRC_B200_LIB=$PWD/min_ratio_cycle_b200/libmrc_b200.ppf0.so NO_PASSES=1 SEEDS=0,1 MODES=1:0,1:1,4:0 timeout 600 python tools/perf_probe.py 2>&1 | cut -c1-250
