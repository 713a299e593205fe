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
