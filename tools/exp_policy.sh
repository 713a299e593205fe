MRC_B200_SPARSE=1 timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for rep in 1 2 3 4; do
  MRC_SPARSE_LOOK=1 MRC_SPARSE=1 MRC_SEEDS=1,0 MRC_KS=1,4 MRC_QUIET=1 python tools/trace_solve.py 2>&1 | grep "mrc_solve\|total wall" | cut -c1-220
done
