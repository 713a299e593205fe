"""Public API end to end at C3: add_edges_arrays + solve() from insertion-order NumPy arrays, with a breakdown."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import min_ratio_cycle_b200.solver as S
from min_ratio_cycle_b200 import synthetic
n, U, V, C, T = synthetic.random_float(int(os.environ.get("N", 1_000_000)), int(os.environ.get("M", 10_000_000)), seed=0)
for rep in range(4):
    t0 = time.perf_counter()
    s = S.MinRatioCycleSolver(n, S.SolverConfig(log_level=S.LogLevel.NONE))
    t1 = time.perf_counter()
    s.add_edges_arrays(U, V, C, T)
    t2 = time.perf_counter()
    s._preprocess_graph()
    t3 = time.perf_counter()
    s._build_arrays_if_needed()
    t4 = time.perf_counter()
    r = s.solve()
    t5 = time.perf_counter()
    print(f"rep {rep}: total {1e3*(t5-t0):.1f} ms = add {1e3*(t2-t1):.1f} + preprocess/build {1e3*(t3-t2):.1f} + {1e3*(t4-t3):.1f} + solve() {1e3*(t5-t4):.1f}  "
          f"ratio {r.ratio!r} mirrors_pending {s._mirrors_pending}", flush=True)
    t0 = time.perf_counter(); _ = s._U; print(f"   mirrors on demand: {1e3*(time.perf_counter()-t0):.1f} ms")
    del s
