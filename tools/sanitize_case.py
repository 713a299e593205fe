"""Small end-to-end exercise of every kernel for compute-sanitizer (no torch import: ctypes + numpy only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from min_ratio_cycle_b200 import _native as N, synthetic

def sorted_graph(gen, **kw):
    n, U, V, C, T = gen(**kw)
    o = np.argsort(V, kind="stable")
    return n, U[o], V[o], C[o], T[o]

n, U, V, C, T = sorted_graph(synthetic.random_float, n=3001, m=17003, seed=3)     # odd sizes: tail paths
g = N.DeviceGraph(n, U, V, C, T)
lo, hi = g.bounds()
ratios = []
for K in (1, 2, 4, 8):
    ratios.append(g.solve_numeric(lo, hi, K=K)["ratio"])
g.set_option("compact_min_edges", 0)          # compact-survivor path (a K >= 2 probe narrowed to one slot)
for opts in ({"compact": 1}, {"sparse": 1, "sparse_min_edges": 0}, {"narrow": 0, "compact": 0}):
    for k, v in opts.items():
        g.set_option(k, v)
    for K in (2, 4, 8):
        ratios.append(g.solve_numeric(lo, hi, K=K)["ratio"])
g.set_option("sparse", 0); g.set_option("narrow", 1); g.set_option("compact", 1)
assert max(ratios) - min(ratios) <= 1e-12, ratios
r = ratios[0]
out = g.probe_numeric_scenarios([r - 1e-12] * 3, [5, -1, 17002], [0.5, 0.0, -3.0], [0.0, 0.0, 0.1])
g.patch_edges([3, 9], [0.25, -0.5], [0.0, 0.5]); g.solve_numeric(lo, hi, K=4); g.restore()
ms = g.bench_relax([r - 1.0, r - 2.0], 3)
g.close()
n, U, V, C, T = sorted_graph(synthetic.random_int, n=500, m=3001, seed=2)
gi = N.DeviceGraph(n, U, V, C, T, C.astype(np.int64), T.astype(np.int64))
a = gi.solve_exact(strategy=1); b = gi.solve_exact(strategy=0)
assert a["cycle"] == b["cycle"] and (a["a"], a["b"]) == (b["a"], b["b"]) and a["rc"] == 0
gi.probe_exact([a["a"] * 7 + 1, a["a"]], [a["b"] * 7, a["b"]])
gi.close()
graphs = [synthetic.currency_graph(17, 50 + i) for i in range(5)] + [synthetic.random_float(9, 30, seed=1)]
from min_ratio_cycle_b200 import batch
res = batch.solve_min_ratio_cycle_batch(graphs)
assert all(x.success for x in res)
so, do, co, to, order, dups = N.sort_edges_by_dst(1000, np.arange(5001) % 1000, (np.arange(5001) * 7) % 1000, np.ones(5001), np.ones(5001))
assert np.all(np.diff(do) >= 0) and dups >= 0
print("sanitize_case OK", ratios[0], a["a"], a["b"], len(res))
