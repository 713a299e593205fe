"""C5: sensitivity_analysis with 1,024 single-edge perturbation scenarios on n=100k, m=2M."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from min_ratio_cycle_b200 import synthetic
from min_ratio_cycle_b200.solver import MinRatioCycleSolver, SolverConfig, LogLevel
n, U, V, C, T = synthetic.random_float(100_000, 2_000_000, seed=5)
s = MinRatioCycleSolver(n, SolverConfig(log_level=LogLevel.NONE))
s.add_edges_arrays(U, V, C, T)
t0 = time.perf_counter(); base = s.solve(); t1 = time.perf_counter()
print(f"baseline solve() {1e3*(t1-t0):.0f} ms ratio {base.ratio!r} cycle {base.cycle}")
rng = np.random.default_rng(5)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
idx = rng.integers(0, len(U), S); sign = rng.choice([-1.0, 1.0], S)
scen = [{int(i): (float(sg * 0.1 * C[i]), 0.0)} for i, sg in zip(idx, sign)]
s._dev.reset_stats()
t0 = time.perf_counter(); res = s.sensitivity_analysis(scen); dt = time.perf_counter() - t0
st = s._dev.stats()
moved = sum(r.cycle != base.cycle for r in res)
print(f"{S} scenarios: {dt:.2f} s = {1e3*dt/S:.2f} ms/scenario; optimum moved in {moved}; relax launches {st['relax_launches']}, "
      f"relax {st['relax_ms']:.0f} ms, checks {st['check_ms']:.0f} ms, rounds {st['rounds']}")
