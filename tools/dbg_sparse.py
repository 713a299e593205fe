import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from min_ratio_cycle_b200 import _native as N, synthetic
n, U, V, C, T = synthetic.random_float(50_000, 500_000, seed=9)
o = np.argsort(V, kind="stable")
args = (n, U[o].astype(np.int32), V[o].astype(np.int32), C[o], T[o])
g = N.DeviceGraph(*args); lo, hi = g.bounds(); want = g.solve_numeric(lo, hi, K=4)["ratio"]; g.close()
for persist in (0, 1):
    g = N.DeviceGraph(*args)
    for k, v in (("sparse", 1), ("sparse_min_edges", 0), ("sparse_look", 1), ("sparse_div", 2), ("persist", persist)):
        g.set_option(k, v)
    for lam in (want - 1e-6, want + 1e-4, want - 1e-6):
        g.set_option("sparse_poison", 1)
        g.reset_stats()
        p = g.probe_numeric([lam])[-1]
        st = g.stats()
        print(persist, lam - want, p["verdict"], p["passes"], {k: st[k] for k in ("relax_launches", "sparse_passes", "check_launches", "probe_launches", "kernel_launches")}, flush=True)
    g.close()
