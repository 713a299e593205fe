#!/bin/bash
cd "$(dirname "$0")/.."
for G in 4 3 2; do
  echo "=== G=$G"
  W=2 KLOCAL=4 MRC_B200_DPROBE_GROUP=$G MRC_B200_DPROBE_MAXPASS=600 timeout 120 python tools/dbg_mg.py 2>&1 | grep -v "^ *tick" | tail -12
done
