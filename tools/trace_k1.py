import os, sys
sys.path.insert(0, '/root/repo' if os.path.exists('/root/repo/tools') else '.')
import numpy as np
from min_ratio_cycle_b200 import _native as N, synthetic
import time
for seed in (0, 1, 3):
    n, U, V, C, T = synthetic.random_float(1_000_000, 10_000_000, seed=seed)
    o = np.argsort(V, kind="stable")
    g = N.DeviceGraph(n, U[o].astype(np.int32), V[o].astype(np.int32), C[o], T[o])
    lo, hi = g.bounds()
    for ce in (0, 1, 2):
        for sparse in (0, 1):
            g.set_option("sparse", sparse); g.set_option("check_every", ce)
            g.solve_numeric(lo, hi, K=1)
            os.environ.pop("MRC_B200_TRACE", None)
            t0 = time.perf_counter()
            for _ in range(3): out = g.solve_numeric(lo, hi, K=1)
            dt = (time.perf_counter() - t0) / 3
            print(f"seed {seed} check_every {ce} sparse {sparse}: {dt*1e3:.2f} ms rounds {out['iterations']}", flush=True)
    g.set_option("check_every", 1); g.set_option("sparse", 0)
    sys.stderr.flush()
    print(f"--- trace seed {seed} check_every=1 dense", flush=True)
    # trace is read at solve time via getenv
    os.environ["MRC_B200_TRACE"] = "1"
    g.solve_numeric(lo, hi, K=1)
    os.environ.pop("MRC_B200_TRACE")
    g.close()
