import sys; sys.path.insert(0, "/root/repo")
import numpy as np
from min_ratio_cycle_b200 import synthetic
from min_ratio_cycle_b200.solver import MinRatioCycleSolver, SolverConfig, LogLevel
worst = 0
for name, gen in (("currency64", lambda s: synthetic.currency_graph(64, 1000 + s)), ("dense500", lambda s: synthetic.random_float(500, 8000, seed=s)),
                  ("ring+chords", lambda s: synthetic.ring_plus_random(2000, 4000, seed=s)), ("int ties", lambda s: synthetic.random_int(300, 1500, seed=s)),
                  ("sparse20k", lambda s: synthetic.random_float(20000, 60000, seed=s))):
    its = []
    for s in range(8):
        n, U, V, C, T = gen(s)
        sv = MinRatioCycleSolver(n, SolverConfig(log_level=LogLevel.NONE))
        sv.add_edges_arrays(U, V, C, T)
        r = sv.solve()
        its.append(r.metrics.iterations)
    worst = max(worst, max(its))
    print(name, "iterations (K=1 default):", its)
print("worst", worst)
