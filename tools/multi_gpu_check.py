"""
Launched with torchrun on N GPUs: checks the two sharding modes against single-GPU results.
  (1) lambda-candidate sharding on a 100k/1M float graph == local mrc_solve_numeric (ratio within 1e-12)
  (2) graph-batch sharding of 64 currency graphs == local batch solve
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from min_ratio_cycle_b200 import _native as N, sharded, synthetic, batch
from min_ratio_cycle_b200.solver import SolverConfig, LogLevel

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, U, V, C, T = synthetic.random_float(100_000, 1_000_000, seed=3)
o = np.argsort(V, kind="stable")
g = N.DeviceGraph(n, U[o], V[o], C[o], T[o], device=local)
lo, hi = g.bounds()
ref = g.solve_numeric(lo, hi, K=4)
r = sharded.solve_numeric_sharded(lambda l, s, f: g.probe_numeric(l, s, f), lo, hi, K_local=4)
assert r.rc == N.OK and abs(r.ratio - ref["ratio"]) <= 1e-12, (r.ratio, ref["ratio"])
assert r.cycle == ref["cycle"], (r.cycle, ref["cycle"])
graphs = [synthetic.currency_graph(64, 2000 + i) for i in range(64)]
cfg = SolverConfig(log_level=LogLevel.NONE, gpu_device=local)
allr = batch.solve_min_ratio_cycle_batch_sharded(graphs, cfg)
mine = batch.solve_min_ratio_cycle_batch(graphs, cfg)
assert len(allr) == 64 and all(a.ratio == b.ratio and a.cycle == b.cycle for a, b in zip(allr, mine))
dist.barrier()
if rank == 0:
    print(f"multi_gpu_check OK world={world}: sharded ratio {r.ratio!r} rounds {r.iterations} (single-GPU rounds {ref['iterations']}); "
          f"batch of 64 graphs identical across sharding")
dist.destroy_process_group()
