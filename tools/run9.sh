#!/bin/bash
cd "$(dirname "$0")/.."
export MRC_B200_LIB=$PWD/min_ratio_cycle_b200/libmrc_b200.dbg.so
for i in 1 2; do
  echo "=== try $i"
  W=2 KLOCAL=4 MRC_B200_DPROBE_GROUP=4 MRC_B200_DPROBE_MAXPASS=1200 timeout 120 python tools/dbg_mg.py 2>&1 | grep -v "^ *tick" | tail -150
done
