"""Per-pass relaxation timing for one library build (MRC_B200_LIB selects it)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from min_ratio_cycle_b200 import _native as N, synthetic
n, m = 1_000_000, 10_000_000
cache = "/tmp/c3.npz"
if os.path.exists(cache):
    z = np.load(cache); U, V, C, T = z["U"], z["V"], z["C"], z["T"]
else:
    n, U, V, C, T = synthetic.random_float(n, m, seed=0)
    o = np.argsort(V, kind="stable"); U, V, C, T = U[o], V[o], C[o], T[o]
    np.savez(cache, U=U, V=V, C=C, T=T)
g = N.DeviceGraph(n, U, V, C, T)
r = -14.937873966124124
tag = os.environ.get("MRC_B200_LIB", "default").split("/")[-1]
tma = int(os.environ.get("MRC_TMA", "0"))
g.set_option("tma", tma)
if os.environ.get("MRC_FRONTIER"): g.set_option("frontier", 1); tag += " frontier"
tag += f" tma={tma}"
for K in (1, 2, 4, 8):
    lam = [r - 0.5 - 0.1 * k for k in range(K)]   # all NONNEG: converges in ~25 passes
    g.bench_relax(lam, 3)
    ms = g.bench_relax(lam, 40)
    us = ms * 1e3
    print(f"{tag:28s} K={K} pass1={us[0]:.0f} pass2={us[1]:.0f} pass3={us[2]:.0f} p5={us[4]:.0f} p10={us[9]:.0f} "
          f"p4={us[3]:.0f} p6={us[5]:.0f} p8={us[7]:.0f} p15={us[14]:.0f} p20={us[19]:.0f} p30={us[29]:.0f} p40={us[39]:.0f} mean={us.mean():.1f} late_mean={us[20:].mean():.1f}", flush=True)
# update-heavy launches: candidate sets of a real K=4 solve (first round far from lambda*, a middle round, the
# certifying round next to lambda*); us for passes 1, 2, 3-6 (mean), 7-14 (mean), 15-30 (mean) and the 30-pass total
if not os.environ.get("MRC_SKIP_SETS"):
    sets = {"far": [-60.5, -20.26, 19.99, 60.24], "mid": [-15.16, -10.06, -4.96, 0.137],
            "near": [r - 0.17, r - 0.11, r - 0.056, r - 7.5e-13], "near1": [r - 7.5e-13], "mid1": [-10.06]}
    for name, lam in sets.items():
        g.bench_relax(lam, 3)
        us = g.bench_relax(lam, 30) * 1e3
        print(f"{tag:28s} set={name:6s} K={len(lam)} p1={us[0]:.0f} p2={us[1]:.0f} p3-6={us[2:6].mean():.0f} p7-14={us[6:14].mean():.0f} "
              f"p15-30={us[14:].mean():.0f} total30={us.sum():.0f}", flush=True)
