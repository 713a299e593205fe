"""Print the decisions of the asynchronous multisection driver (MRC_B200_TRACE=1) for one solve."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from min_ratio_cycle_b200 import _native as N, synthetic
n, m, seed = int(os.environ.get("MRC_N", 1000000)), int(os.environ.get("MRC_M", 10000000)), int(os.environ.get("SEED", 0))
n, U, V, C, T = synthetic.random_float(n, m, seed=seed)
o = np.argsort(V, kind="stable")
g = N.DeviceGraph(n, U[o], V[o], C[o], T[o])
g.set_option("async", int(os.environ.get("ASYNC", 1)))
lo, hi = g.bounds()
out = g.solve_numeric(lo, hi, K=int(os.environ.get("K", 4)))
st = g.stats()
print("ratio", out["ratio"], "iters", out["iterations"], "relax", st["relax_launches"], round(st["relax_ms"], 2), "check", st["check_launches"], round(st["check_ms"], 2))
