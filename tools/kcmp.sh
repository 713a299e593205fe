# K=4 multisection against K=1 Newton over several graph families (worklist passes on); profiles/r01e_policy.txt section 11
for cfg in "1000000 10000000 0,1,2,3,4,5,6" "100000 2000000 1,2,3,4" "300000 3000000 1,2,3" "30000 300000 1,2,3"; do set -- $cfg
  MRC_SPARSE=1 MRC_N=$1 MRC_M=$2 MRC_SEEDS=$3 MRC_KS=4,1 MRC_QUIET=1 python tools/trace_solve.py 2>&1 | grep "mrc_solve" | cut -c1-60 | sed "s/^/n=$1 /"
done
