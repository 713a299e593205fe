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
r = sharded.solve_numeric_sharded(lambda l, s, f: g.probe_numeric(l, s, f), lo, hi, K_local=4, fetch=g.fetch_cycle_sums)
assert r.rc == N.OK and abs(r.ratio - ref["ratio"]) <= 1e-12, (r.ratio, ref["ratio"])
assert r.cycle == ref["cycle"], (r.cycle, ref["cycle"])
graphs = [synthetic.currency_graph(64, 2000 + i) for i in range(64)]
cfg = SolverConfig(log_level=LogLevel.NONE, gpu_device=local)
allr = batch.solve_min_ratio_cycle_batch_sharded(graphs, cfg)
mine = batch.solve_min_ratio_cycle_batch(graphs, cfg)
assert len(allr) == 64 and all(a.ratio == b.ratio and a.cycle == b.cycle for a, b in zip(allr, mine))
# C4 at full size: 4,096 currency graphs split over the ranks (independent units, no data-path collective)
import time
B = 4096
mine_idx = sharded.shard_units(B, rank, world)
big = [synthetic.currency_graph(64, 1000 + i) for i in mine_idx]
nv, off, S, D, Cs, Ts = batch._prepare(big, True)
N.batch_solve_numeric(nv, off, S, D, Cs, Ts, K=4, device=local)           # warm-up
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
outb = N.batch_solve_numeric(nv, off, S, D, Cs, Ts, K=4, device=local)
torch.cuda.synchronize()
loc = torch.tensor([time.perf_counter() - t0, outb["stats"]["relax_ms"] * 1e-3, float(outb["stats"]["relax_edge_visits"]),
                    float((outb["status"] == 0).sum())], dtype=torch.float64, device="cuda")
allv = [torch.zeros_like(loc) for _ in range(world)]
dist.all_gather(allv, loc)
allv = torch.stack(allv).cpu().numpy()
dist.barrier()
if rank == 0:
    print(f"C4 sharded x{world}: {int(allv[:,3].sum())}/{B} solved; call time max over ranks {1e3*allv[:,0].max():.1f} ms "
          f"(kernel {1e3*allv[:,1].max():.2f} ms); {allv[:,2].sum()/allv[:,1].max()/1e9:.0f} G edge-relax/s aggregate (kernel time)")
if rank == 0:
    print(f"multi_gpu_check OK world={world}: sharded ratio {r.ratio!r} rounds {r.iterations} (single-GPU rounds {ref['iterations']}); "
          f"batch of 64 graphs identical across sharding")
dist.destroy_process_group()
