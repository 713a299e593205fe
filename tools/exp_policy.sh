timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
for sp in 1 0; do echo "== sparse $sp"; MRC_SPARSE=$sp MRC_SEEDS=0,1,2 MRC_KS=4,1 MRC_QUIET=1 python tools/trace_solve.py 2>&1 | grep "mrc_solve\|total wall" | cut -c1-220; done
MRC_SEEDS=0 MRC_KS=4 python tools/trace_solve.py 2>&1 | cut -c1-360 | head -12
