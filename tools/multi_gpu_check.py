"""
Multi-GPU checks of the in-library sharded search (csrc/mrc_comm.cuh) against single-GPU results.
  torchrun (one process per GPU):  python -m torch.distributed.run --nproc-per-node N tools/multi_gpu_check.py
      Comm.from_torch + DeviceGraph.replicated + solve_numeric_sharded on C3 (or N/M env) == local mrc_solve_numeric
  single process:                  python tools/multi_gpu_check.py group
      Comm.group(all devices) + MultiGraph.solve_numeric == local mrc_solve_numeric
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from min_ratio_cycle_b200 import _native as N, synthetic

n = int(os.environ.get("N", 1_000_000)); m = int(os.environ.get("M", 10_000_000)); K = int(os.environ.get("KLOCAL", 4))
seeds = [int(x) for x in os.environ.get("SEEDS", "0").split(",")]
reps = int(os.environ.get("REPS", 5))


def graph(seed):
    _, U, V, C, T = synthetic.random_float(n, m, seed=seed)
    o = np.argsort(V, kind="stable")
    return U[o].astype(np.int32), V[o].astype(np.int32), C[o], T[o]


def same(a, b):
    return a["rc"] == b["rc"] == N.OK and abs(a["ratio"] - b["ratio"]) <= 1e-12 * max(1.0, abs(b["ratio"])) and a["cycle"] == b["cycle"]


if len(sys.argv) > 1 and sys.argv[1] == "group":
    ndev = N.device_count()
    widths = [w for w in (1, 2, 4, 8) if w <= ndev]
    comms = {W: N.Comm.group(list(range(W))) for W in widths}
    for seed in seeds:
        U, V, C, T = graph(seed)
        g = N.DeviceGraph(n, U, V, C, T)
        lo, hi = g.bounds()
        ref = g.solve_numeric(lo, hi, K=K)
        for W in widths:
            t0 = time.perf_counter()
            mg = N.MultiGraph(comms[W], n, U, V, C, T)
            t_rep = time.perf_counter() - t0
            for kl in (K, 1):
                out = mg.solve_numeric(lo, hi, K_local=kl)
                t0 = time.perf_counter()
                for _ in range(reps):
                    out = mg.solve_numeric(lo, hi, K_local=kl)
                dt = (time.perf_counter() - t0) / reps
                assert same(out, ref), (W, seed, kl, out["ratio"], ref["ratio"], out["cycle"][:8], ref["cycle"][:8], out["rc"])
                if os.environ.get("TRACE_W") == str(W) and os.environ.get("TRACE_K") == str(kl):
                    os.environ["MRC_B200_TRACE"] = "1"; mg.solve_numeric(lo, hi, K_local=kl); os.environ.pop("MRC_B200_TRACE")
                visits = sum(mg.replica(i).stats()["relax_edge_visits"] for i in range(W))
                for i in range(W):
                    mg.replica(i).reset_stats()
                print(f"group x{W} seed {seed} K_local {kl}: {dt*1e3:.2f} ms/solve ticks {out['ticks']} probes(rank0) {out['iterations']} "
                      f"(replication {t_rep*1e3:.1f} ms)", flush=True)
            mg.close()
        g.close()
    for c in comms.values():
        c.close()
    print("multi_gpu_check group OK")
    sys.exit(0)

import torch, torch.distributed as dist
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
comm = N.Comm.from_torch(local)
for seed in seeds:
    U, V, C, T = graph(seed) if rank == 0 else (None, None, None, None)
    t0 = time.perf_counter()
    g = N.DeviceGraph.replicated(comm, n, m, U, V, C, T)
    torch.cuda.synchronize(); dist.barrier()
    t_rep = time.perf_counter() - t0
    lo, hi = g.bounds()
    ref = g.solve_numeric(lo, hi, K=K)
    for kl in (K, 1):
        out = g.solve_numeric_sharded(comm, lo, hi, K_local=kl)
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            out = g.solve_numeric_sharded(comm, lo, hi, K_local=kl)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        assert same(out, ref), (rank, seed, kl, out["ratio"], ref["ratio"], out["cycle"][:8], ref["cycle"][:8], out["rc"])
        if rank == 0:
            print(f"torchrun x{world} seed {seed} K_local {kl}: {dt*1e3:.2f} ms/solve ticks {out['ticks']} probes(rank0) {out['iterations']} "
                  f"(replication {t_rep*1e3:.1f} ms incl. barrier); single GPU K={K}: rounds {ref['iterations']}", flush=True)
    g.close()
comm.close()
if rank == 0:
    print(f"multi_gpu_check torchrun OK world={world}")
dist.destroy_process_group()
