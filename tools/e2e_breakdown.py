"""Where does the e2e step (host buffers -> result) spend its time?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from min_ratio_cycle_b200 import _native as N
z = np.load("/tmp/c3.npz") if os.path.exists("/tmp/c3.npz") else None
if z is None:
    from min_ratio_cycle_b200 import synthetic
    n, U, V, C, T = synthetic.random_float(1_000_000, 10_000_000, seed=0)
    o = np.argsort(V, kind="stable"); U, V, C, T = U[o], V[o], C[o], T[o]
else:
    U, V, C, T = z["U"], z["V"], z["C"], z["T"]
n = 1_000_000
pin = [torch.from_numpy(a).pin_memory() for a in (U.astype(np.int32), V.astype(np.int32), C, T)]
pU, pV, pC, pT = [p.numpy() for p in pin]
for it in range(5):
    t0 = time.perf_counter(); g = N.DeviceGraph(n, pU, pV, pC, pT); t1 = time.perf_counter()
    lo, hi = g.bounds(); t2 = time.perf_counter()
    out = g.solve_numeric(lo, hi, K=4); t3 = time.perf_counter()
    g.close(); t4 = time.perf_counter()
    print(f"create {1e3*(t1-t0):.2f} ms  bounds {1e3*(t2-t1):.2f}  solve {1e3*(t3-t2):.2f}  destroy {1e3*(t4-t3):.2f}  total {1e3*(t4-t0):.2f}")
# raw H2D rate from pinned memory
d = torch.empty(240_000_000, dtype=torch.uint8, device="cuda")
h = torch.empty(240_000_000, dtype=torch.uint8).pin_memory()
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(h, non_blocking=True); torch.cuda.synchronize()
    print(f"raw pinned H2D 240 MB: {1e3*(time.perf_counter()-t0):.2f} ms")
