"""One-off soak: dense / sparse / async / K=1 solves of several C3-sized graphs must agree; the oracle certifies each optimum."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from min_ratio_cycle_b200 import _native as N, synthetic
from oracle import oracle
n, m = int(os.environ.get("MRC_N", 1_000_000)), int(os.environ.get("MRC_M", 10_000_000))
bad = 0
for seed in [int(s) for s in os.environ.get("MRC_SEEDS", "3,4,5,6").split(",")]:
    n, U, V, C, T = synthetic.random_float(n, m, seed=seed)
    o = np.argsort(V, kind="stable")
    Us, Vs, Cs, Ts = U[o], V[o], C[o], T[o]
    g = N.DeviceGraph(n, Us, Vs, Cs, Ts)
    lo, hi = g.bounds()
    res = {}
    for name, opts, K in (("dense K=4", {}, 4), ("sparse K=4", {"sparse": 1}, 4), ("async K=4", {"async": 1}, 4),
                          ("dense K=1", {}, 1), ("sparse K=1", {"sparse": 1}, 1), ("sparse+look1 K=4", {"sparse": 1, "sparse_look": 1}, 4)):
        for k in ("sparse", "async", "sparse_look"):
            g.set_option(k, 0)
        for k, v in opts.items():
            g.set_option(k, v)
        t0 = time.perf_counter(); out = g.solve_numeric(lo, hi, K=K); dt = time.perf_counter() - t0
        res[name] = (out["rc"], out["ratio"], len(out["cycle"]) - 1, dt)
    ref = res["dense K=4"][1]
    ok = all(r[0] == 0 and abs(r[1] - ref) <= 1e-9 * max(1.0, abs(ref)) for r in res.values())
    orc = oracle.Oracle(n, U, V, C, T, all_int=False, threads=16)
    orc.set_quirks(False)
    cert, why = orc.certify(out["cycle"], out["ratio"])
    print(f"seed {seed}: ratio {ref!r} len {res['dense K=4'][2]} agree={ok} certified={cert} ({why}) | " +
          " ".join(f"{k}: {1e3*v[3]:.1f} ms" for k, v in res.items()), flush=True)
    bad += (not ok) or (not cert)
    g.close()
print("SOAK", "FAILED" if bad else "OK")
