for bps in 1 2 4 8; do
  echo "== sparse blocks/SM $bps"
  MRC_B200_SPARSE_BPS=$bps MRC_SPARSE=1 MRC_SEEDS=0,1 MRC_KS=4,1 MRC_QUIET=1 python tools/trace_solve.py 2>&1 | grep "mrc_solve" | cut -c1-220
  MRC_B200_SPARSE_BPS=$bps python tools/explore_c5.py 2>&1 | grep scenarios | cut -c1-200
done
