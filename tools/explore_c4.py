"""C4: 4,096 currency graphs (n=64, m=4032) through the batch kernel; timing + relax/s."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from min_ratio_cycle_b200 import batch, synthetic, _native as N
from min_ratio_cycle_b200.solver import SolverConfig, LogLevel
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
t0 = time.time()
graphs = [synthetic.currency_graph(64, 1000 + g) for g in range(B)]
print(f"gen {time.time()-t0:.1f}s")
for K in (1, 2, 4, 8, 4):
    t0 = time.time()
    nv, off, S, D, Cs, Ts = batch._prepare(graphs, True)
    tp = time.time() - t0
    t0 = time.time()
    out = N.batch_solve_numeric(nv, off, S, D, Cs, Ts, K=K)
    dt = time.time() - t0
    st = out["stats"]
    print(f"K={K} prepare={tp*1e3:.0f}ms call={dt*1e3:.1f}ms kernel={st['relax_ms']:.1f}ms relax={st['relax_edge_visits']/1e9:.2f}G "
          f"GTEPS={st['relax_edge_visits']/st['relax_ms']/1e6:.1f} ok={(out['status']==0).sum()} rounds_mean={out['iterations'].mean():.1f} "
          f"ratio[0]={out['ratio'][0]!r}", flush=True)
