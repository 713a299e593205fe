python -m pytest tests/test_gpu_parity.py tests/test_gpu_stress.py -x -q -m gpu 2>&1 | tail -3
MRC_SEEDS=0,1,2 MRC_KS=4,1 python tools/trace_solve.py 2>&1 | cut -c1-330
