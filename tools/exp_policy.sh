python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
run() { # label env...
  label=$1; shift
  for cfg in "1000000 10000000 0,1,2" "100000 2000000 1,2,3" "10000 50000 1,2,3"; do set -- $cfg "$@"; n=$1; m=$2; seeds=$3; shift 3
    echo "== $label n=$n m=$m"; env "$@" MRC_QUIET=1 MRC_N=$n MRC_M=$m MRC_SEEDS=$seeds MRC_KS=4 python tools/trace_solve.py 2>&1 | grep "total wall\|mrc_solve" | cut -c1-200
  done
}
run gate0 MRC_GATE=0 MRC_BURST_MAX=8
run gate1_b8 MRC_BURST_MAX=8
run gate1_b16 MRC_BURST_MAX=16
run gate1_b32 MRC_BURST_MAX=32
run patience_b16 MRC_PATIENCE=1 MRC_BURST_MAX=16
