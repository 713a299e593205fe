"""Round-by-round trace of the C3 numeric solve (same planner / reducer as mrc_solve_numeric, driven from Python)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from min_ratio_cycle_b200 import _native as N, synthetic

STOP_R0 = int(os.environ.get("MRC_STOP_R0", "0")); STOP_ALL = int(os.environ.get("MRC_STOP_ALL", "0"))
PATIENCE = int(os.environ.get("MRC_PATIENCE", "0"))
QUIET = int(os.environ.get("MRC_QUIET", "0"))
VN = {0: "NONNEG", 1: "NEG", 2: "NEG_NOCYC", 3: "TIE", 4: "skipNEG", 5: "skipNONNEG", 6: "ABANDON", -1: "?"}


def load(n, m, seed):
    cache = f"/tmp/c3_{n}_{m}_{seed}.npz"
    if os.path.exists(cache):
        z = np.load(cache); return n, z["U"], z["V"], z["C"], z["T"]
    n, U, V, C, T = synthetic.random_float(n, m, seed=seed)
    order = np.argsort(V, kind="stable")
    U, V, C, T = U[order], V[order], C[order], T[order]
    np.savez(cache, U=U, V=V, C=C, T=T)
    return n, U, V, C, T


def trace(g, K, tol=1e-12, flags=1):
    lo, hi = g.bounds()
    have = False
    tot = 0.0
    for rnd in range(60):
        lam = N.plan_candidates(lo, hi, have, tol, K)
        g.reset_stats()
        t0 = time.perf_counter()
        fl = flags | (2 if (STOP_R0 and not have) or STOP_ALL else 0) | (4 if PATIENCE else 0)
        res = g.probe_numeric(lam, flags=fl)
        dt = time.perf_counter() - t0
        tot += dt
        st = g.stats()
        vd = [r["verdict"] for r in res]
        cr = [r["sum_cost"] / r["sum_time"] if r["verdict"] in (1, 3) else float("nan") for r in res]
        lo, hi, have, best = N.reduce_round(lam, vd, cr, lo, hi, have)
        desc = " ".join(f"[{l:.6g} {VN[v]} p{r['passes']}{' len%d r=%.9g' % (r['cycle_len'] - 1, c) if v in (1, 3) else ''}]"
                        for l, v, r, c in zip(lam, vd, res, cr))
        if not QUIET: print(f"  round {rnd}: wall {dt*1e3:.2f} ms relax {st['relax_launches']}L {st['relax_ms']:.2f} ms  "
              f"sparse {st['sparse_passes']}P {st['sparse_ms']:.2f} ms  check {st['check_launches']}L {st['check_ms']:.2f} ms | {desc} -> lo={lo:.9g} hi={hi:.12g}", flush=True)
        if hi - lo <= tol:
            break
    print(f"  total wall {tot*1e3:.2f} ms, rounds {rnd+1}, ratio {hi!r}", flush=True)


def main():
    n = int(os.environ.get("MRC_N", 1_000_000)); m = int(os.environ.get("MRC_M", 10_000_000))
    for seed in [int(s) for s in os.environ.get("MRC_SEEDS", "0").split(",")]:
        n, U, V, C, T = load(n, m, seed)
        g = N.DeviceGraph(n, U, V, C, T)
        for opt in ("burst_init", "frontier", "frontier_from", "stop_on_neg", "narrow", "gate", "patience", "burst_max", "prune_above", "alternate", "sparse", "sparse_div", "sparse_look", "async"):
            if os.environ.get("MRC_" + opt.upper()): g.set_option(opt, int(os.environ["MRC_" + opt.upper()]))
        for K in [int(k) for k in os.environ.get("MRC_KS", "4,1").split(",")]:
            for rep in range(2):
                print(f"seed {seed} K={K} rep {rep}", flush=True)
                trace(g, K)
            g.reset_stats()
            t0 = time.perf_counter()
            lo, hi = g.bounds()
            out = g.solve_numeric(lo, hi, K=K)
            dt = time.perf_counter() - t0
            st = g.stats()
            print(f"seed {seed} K={K} mrc_solve_numeric: {dt*1e3:.2f} ms rounds={out['iterations']} relax {st['relax_launches']}L "
                  f"{st['relax_ms']:.2f} ms sparse {st['sparse_passes']}P {st['sparse_ms']:.2f} ms check {st['check_launches']}L {st['check_ms']:.2f} ms ratio={out['ratio']!r}", flush=True)


main()
