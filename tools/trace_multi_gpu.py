"""Timeline of ONE multi-GPU solve (single process, one host thread per device): W=<ranks> KLOCAL=<candidates per rank> python tools/trace_multi_gpu.py
prints the ticks of the first round, every distributed probe with its device-time breakdown per rank, and the cycle hand-over (MRC_B200_TRACE=1)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from min_ratio_cycle_b200 import _native as N, synthetic
n = int(os.environ.get("N", 1_000_000)); m = int(os.environ.get("M", 10_000_000))
W = int(os.environ.get("W", 2)); kl = int(os.environ.get("KLOCAL", 8))
_, U, V, C, T = synthetic.random_float(n, m, seed=0)
o = np.argsort(V, kind="stable")
U, V, C, T = U[o].astype(np.int32), V[o].astype(np.int32), C[o], T[o]
g = N.DeviceGraph(n, U, V, C, T); lo, hi = g.bounds(); ref = g.solve_numeric(lo, hi, K=4); print("ref", ref["ratio"], ref["cycle"], flush=True)
comm = N.Comm.group(list(range(W)))
mg = N.MultiGraph(comm, n, U, V, C, T)
os.environ["MRC_B200_TRACE"] = "1"
for rep in range(2):
    t0 = time.perf_counter()
    out = mg.solve_numeric(lo, hi, K_local=kl)
    print("rep", rep, "rc", out["rc"], "ratio", out["ratio"], "ticks", out["ticks"], "ms", 1e3 * (time.perf_counter() - t0), flush=True)
mg.close(); comm.close(); g.close()
