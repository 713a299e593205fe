#!/bin/bash
cd "$(dirname "$0")/.."
for o in "k_after_cycle=0" "k_after_cycle=1"; do
  NO_PASSES=1 SEEDS=0,1,3 MODES=4:0,4:1 OPTS=$o timeout 600 python tools/perf_probe.py 2>&1 | cut -c1-330
done
