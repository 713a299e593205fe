"""
Multi-GPU lambda search: one process per GPU (torch.distributed), every rank holds the full
device-resident edge arrays and owns K_local of the K_total = world * K_local candidate lambdas of a
multisection round (sharding (1)).  The only communication is one all-gather of
K_total (verdict, cycle ratio, cycle length) records per round -- tens of bytes over NCCL/NVLink --
and one broadcast of the winning cycle at the end.  No graph data ever crosses the link.

Candidate planning and bracket folding are the SAME host functions the single-GPU driver uses
(mrc_plan_candidates / mrc_reduce_round in libmrc_b200.so), so every rank derives the same bracket.

`probe` is injectable so the collective logic can be exercised on CPU with the gloo backend
(tests/test_sharded_gloo.py); the product path always passes DeviceGraph.probe_numeric.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import _native as N


@dataclass
class ShardedResult:
    cycle: list            # closed
    sum_cost: float
    sum_time: float
    ratio: float
    iterations: int
    lo: float
    hi: float
    rc: int                # N.OK / N.ERR_NO_CYCLE / N.ERR_CONVERGENCE
    probes_local: int = 0


def _dist():
    import torch.distributed as dist
    return dist


def _comm_device(group=None):
    import torch
    dist = _dist()
    backend = dist.get_backend(group)
    return torch.device("cuda", torch.cuda.current_device()) if "nccl" in str(backend) else torch.device("cpu")


def solve_numeric_sharded(probe, lo: float, hi: float, max_iter: int = 60, tol: float = 1e-12,
                          slack: float = 1e-15, K_local: int = 4, group=None) -> ShardedResult:
    """
    probe(lams, slack, flags) -> list of dicts like DeviceGraph.probe_numeric.
    Collective: every rank of `group` must call this with the same lo/hi/max_iter/tol/K_local.
    """
    import torch
    dist = _dist()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    K_total = K_local * world
    dev = _comm_device(group)
    have_cycle = False
    best = None            # (owner rank, dict) -- only the owner keeps the cycle itself
    best_owner = -1
    iterations = 0
    probes_local = 0
    it = -1
    for it in range(max_iter):
        iterations += 1
        lam = N.plan_candidates(lo, hi, have_cycle, tol, K_total)
        mine = lam[rank * K_local:(rank + 1) * K_local]
        res = probe(mine, slack, N.F_MONOTONE_PRUNE)
        probes_local += len(mine)
        rec = torch.zeros(K_local, 2, dtype=torch.float64)
        for k, r in enumerate(res):
            rec[k, 0] = float(r["verdict"])
            rec[k, 1] = (r["sum_cost"] / r["sum_time"]) if r["verdict"] in (N.V_NEG, N.V_TIE) else float("nan")
        rec = rec.to(dev)
        allrec = torch.empty(world * K_local, 2, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allrec, rec, group=group)
        allrec = allrec.cpu().numpy()
        verdict = allrec[:, 0].astype(np.int32)
        ratios = allrec[:, 1].copy()
        lo, hi, have_cycle, best_idx = N.reduce_round(lam, verdict, ratios, lo, hi, have_cycle)
        if best_idx >= 0:
            best_owner = best_idx // K_local
            if best_owner == rank:
                best = res[best_idx - rank * K_local]
        if hi - lo <= tol:
            break
    if best_owner < 0:
        # one more try at the upper bound (solver.py:1079-1085); every rank probes, rank 0 owns it
        r = probe(np.array([hi]), slack, 0)[0]
        probes_local += 1
        if r["verdict"] == N.V_NEG:
            best_owner, best = 0, r
    if best_owner < 0:
        return ShardedResult([], 0.0, 0.0, float("inf"), iterations, lo, hi, N.ERR_NO_CYCLE, probes_local)
    # broadcast the winning cycle from its owner
    head = torch.zeros(3, dtype=torch.float64)
    if rank == best_owner:
        head[0], head[1], head[2] = float(len(best["cycle"])), best["sum_cost"], best["sum_time"]
    head = head.to(dev)
    src = dist.get_global_rank(group, best_owner) if group is not None else best_owner
    dist.broadcast(head, src=src, group=group)
    head = head.cpu()
    L = int(head[0].item())
    cyc = torch.zeros(L, dtype=torch.int64)
    if rank == best_owner:
        cyc[:] = torch.tensor(best["cycle"], dtype=torch.int64)
    cyc = cyc.to(dev)
    dist.broadcast(cyc, src=src, group=group)
    sc, st = float(head[1].item()), float(head[2].item())
    rc = N.ERR_CONVERGENCE if (it == max_iter - 1 and hi - lo > tol) else N.OK
    return ShardedResult(cyc.cpu().tolist(), sc, st, sc / st, iterations, lo, hi, rc, probes_local)


def shard_units(n_units: int, rank: int, world: int) -> range:
    """Independent-unit sharding: contiguous block of graphs / scenarios for this rank."""
    base, extra = divmod(n_units, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))
