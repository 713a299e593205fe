"""Scratch exploration on a GPU box: time the C3 workload pieces (not the bench contract)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from min_ratio_cycle_b200 import _native as N, synthetic

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    m = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
    t0 = time.time()
    cache = "/tmp/c3.npz"
    if os.path.exists(cache) and n == 1_000_000:
        z = np.load(cache); U, V, C, T = z["U"], z["V"], z["C"], z["T"]
    else:
        n, U, V, C, T = synthetic.random_float(n, m, seed=int(os.environ.get('MRC_SEED', '0')))
        order = np.argsort(V, kind="stable")
        U, V, C, T = U[order], V[order], C[order], T[order]
        if n == 1_000_000: np.savez(cache, U=U, V=V, C=C, T=T)
    print(f"sort {time.time()-t0:.2f}s", flush=True)
    t0 = time.time()
    g = N.DeviceGraph(n, U, V, C, T)
    print(f"upload {time.time()-t0:.3f}s", flush=True)
    lo, hi = g.bounds()
    if os.environ.get("MRC_FRONTIER"): g.set_option("frontier", int(os.environ["MRC_FRONTIER"]))
    if os.environ.get("MRC_FFROM"): g.set_option("frontier_from", int(os.environ["MRC_FFROM"]))
    if os.environ.get("MRC_STOP"): g.set_option("stop_on_neg", int(os.environ["MRC_STOP"]))
    if os.environ.get("MRC_BURST"): g.set_option("burst_init", int(os.environ["MRC_BURST"]))
    print("bounds", lo, hi)
    for K in (1, 2, 4, 4, 4):
        g.reset_stats()
        t0 = time.time()
        out = g.solve_numeric(lo, hi, K=K)
        dt = time.time() - t0
        st = g.stats()
        gbps = st["relax_launches"] * 24 * m / (st["relax_ms"] * 1e-3) / 1e9 if st["relax_ms"] else 0
        print(f"K={K} rc={out['rc']} ratio={out['ratio']!r} len={len(out['cycle'])-1} rounds={out['iterations']} "
              f"wall={dt*1e3:.1f}ms relax_launches={st['relax_launches']} relax_ms={st['relax_ms']:.2f} "
              f"check_launches={st['check_launches']} check_ms={st['check_ms']:.2f} "
              f"us/pass={st['relax_ms']*1e3/max(1,st['relax_launches']):.1f} edge GB/s={gbps:.0f} "
              f"GTEPS={st['relax_edge_visits']/(st['relax_ms']*1e-3)/1e9:.1f} cuts={st['false_cycles_cut']}", flush=True)
    # single probes for pass counts
    r = out["ratio"]
    for lam in (r + 1.0, r + 1e-3, r - 1e-3, r - 1.0):
        g.reset_stats()
        t0 = time.time()
        p = g.probe_numeric([lam])[0]
        st = g.stats()
        print(f"probe lam={lam:.6f} verdict={p['verdict']} passes={p['passes']} wall={1e3*(time.time()-t0):.2f}ms "
              f"relax_ms={st['relax_ms']:.2f} check_ms={st['check_ms']:.2f}", flush=True)

main()
