"""Experiment: first round of the lambda search at quantiles of the edge-ratio distribution instead of uniformly over
[min c/t - 1, max c/t + 1] (the rest of the search unchanged, warm-started from the first round's best cycle)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from min_ratio_cycle_b200 import _native as N, synthetic

def run(g, lo, hi, K, qs, ratios_sorted):
    cand = np.sort(np.array([ratios_sorted[int(q * (len(ratios_sorted) - 1))] for q in qs]))
    t0 = time.perf_counter()
    res = g.probe_numeric(cand, 1e-15, N.F_MONOTONE_PRUNE | N.F_NO_CYCLES)
    vd = np.array([r["verdict"] for r in res], np.int32)
    cr = np.array([r["sum_cost"] / r["sum_time"] if r["verdict"] in (N.V_NEG, N.V_TIE) else np.nan for r in res])
    lo2, hi2, hc, best = N.reduce_round(cand, vd, cr, lo, hi, False)
    if hc:
        out = g.solve_numeric(lo2, hi2, K=K, hi_is_cycle=True)
        ratio = out["ratio"] if out["cycle"] else hi2
    else:
        out = g.solve_numeric(lo2, hi2, K=K)
        ratio = out["ratio"]
    return time.perf_counter() - t0, ratio, out["iterations"] + 1, [(round(float(c), 2), int(v), int(r["passes"])) for c, v, r in zip(cand, vd, res)]

n, m = int(os.environ.get("N", 1_000_000)), int(os.environ.get("M", 10_000_000))
for seed in [int(x) for x in os.environ.get("SEEDS", "0,1,3").split(",")]:
    _, U, V, C, T = synthetic.random_float(n, m, seed=seed)
    o = np.argsort(V, kind="stable")
    g = N.DeviceGraph(n, U[o].astype(np.int32), V[o].astype(np.int32), C[o], T[o])
    rs = np.sort(C / T)
    lo, hi = g.bounds()
    for sparse in (0, 1):
        g.set_option("sparse", sparse)
        for K, sets in ((4, ([0.003, 0.0175, 0.103, 0.6], [0.005, 0.02, 0.06, 0.2], [0.01, 0.03, 0.08, 0.25])), (1, ([0.05], [0.1], [0.2]))):
            g.solve_numeric(lo, hi, K=K)
            t0 = time.perf_counter()
            for _ in range(3): ref = g.solve_numeric(lo, hi, K=K)
            base = (time.perf_counter() - t0) / 3
            line = f"seed {seed} sparse {sparse} K={K}: plain {base*1e3:.2f} ms ({ref['iterations']} rounds)"
            for qs in sets:
                run(g, lo, hi, K, qs, rs)
                ts = [run(g, lo, hi, K, qs, rs) for _ in range(3)]
                t = min(x[0] for x in ts)
                assert abs(ts[-1][1] - ref["ratio"]) <= 1e-9 * max(1, abs(ref["ratio"])), (ts[-1][1], ref["ratio"])
                line += f" | q={qs}: {t*1e3:.2f} ms ({ts[-1][2]} rounds) first={ts[-1][3]}"
            print(line, flush=True)
    g.close()
