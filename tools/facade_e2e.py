"""Public-API end-to-end at C3: add_edges_arrays + solve() from insertion-order NumPy arrays."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from min_ratio_cycle_b200 import synthetic
from min_ratio_cycle_b200.solver import MinRatioCycleSolver, SolverConfig, LogLevel
n, U, V, C, T = synthetic.random_float(1_000_000, 10_000_000, seed=0)
for it in range(3):
    t0 = time.perf_counter()
    s = MinRatioCycleSolver(n, SolverConfig(log_level=LogLevel.NONE))
    s.add_edges_arrays(U, V, C, T); t1 = time.perf_counter()
    s._preprocess_graph(); t2 = time.perf_counter()
    s._build_arrays_if_needed(); t3 = time.perf_counter()
    r = s.solve(); t4 = time.perf_counter()
    print(f"add {1e3*(t1-t0):.0f} ms  preprocess {1e3*(t2-t1):.0f}  build+upload {1e3*(t3-t2):.0f}  solve() {1e3*(t4-t3):.0f}  total {1e3*(t4-t0):.0f} ms  ratio {r.ratio!r}")
    s._dev.close()
