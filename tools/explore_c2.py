"""C2: exact mode at n=100k, m=1M integer weights: wall time per strategy."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from min_ratio_cycle_b200 import _native as N, synthetic
n, U, V, C, T = synthetic.random_int(100_000, 1_000_000, seed=0)
o = np.argsort(V, kind="stable")
U, V, C, T = U[o], V[o], C[o], T[o]
g = N.DeviceGraph(n, U, V, C, T, C.astype(np.int64), T.astype(np.int64))
for strategy in (1, 0, 1, 0):
    g.reset_stats(); t0 = time.perf_counter(); out = g.solve_exact(strategy=strategy); dt = time.perf_counter() - t0
    st = g.stats()
    print(f"strategy {strategy}: {1e3*dt:.1f} ms  a/b={out['a']}/{out['b']} len={len(out['cycle'])} SB iterations={out['iterations']} "
          f"GPU probes={st['probes']} relax launches={st['relax_launches']} relax {st['relax_ms']:.1f} ms checks {st['check_ms']:.1f} ms")
