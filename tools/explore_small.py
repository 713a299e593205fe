"""Small graphs: large-graph path (host-driven launches) vs the one-CTA on-device solver (batch kernel, B=1)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from min_ratio_cycle_b200 import _native as N, synthetic
for (n, m) in [(64, 1000), (1000, 5000), (300, 1800), (500, 8000)]:
    _, U, V, C, T = synthetic.random_float(n, m, seed=0)
    o = np.argsort(V, kind="stable"); U, V, C, T = U[o], V[o], C[o], T[o]
    g = N.DeviceGraph(n, U, V, C, T); lo, hi = g.bounds()
    g.solve_numeric(lo, hi, K=4)
    t0 = time.perf_counter()
    for _ in range(5): a = g.solve_numeric(lo, hi, K=4)
    t_big = (time.perf_counter() - t0) / 5
    off = np.array([0, m], np.int64)
    N.batch_solve_numeric([n], off, U, V, C, T, K=4)
    t0 = time.perf_counter()
    for _ in range(5): b = N.batch_solve_numeric([n], off, U, V, C, T, K=4)
    t_one = (time.perf_counter() - t0) / 5
    print(f"n={n} m={m}: host-driven {1e3*t_big:.2f} ms (rounds {a['iterations']}), one-CTA {1e3*t_one:.2f} ms (kernel {b['stats']['relax_ms']:.2f} ms, rounds {int(b['iterations'][0])}); "
          f"ratios {a['ratio']!r} {float(b['ratio'][0])!r}")
    g.close()
