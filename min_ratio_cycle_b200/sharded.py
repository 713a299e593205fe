"""
Multi-GPU lambda search: one process per GPU (torch.distributed), every rank holds the full
device-resident edge arrays and owns K_local of the K_total = world * K_local candidate lambdas of a
multisection round (sharding (1)).  The only communication is one all-gather of
K_total (verdict, cycle ratio, cycle length) records per round -- tens of bytes over NCCL/NVLink --
and one broadcast of the winning cycle at the end (vertex ids as f64: exact below 2^53).  No graph data ever crosses the link.

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
                          slack: float = 1e-15, K_local: int = 4, group=None, fetch=None) -> ShardedResult:
    """
    probe(lams, slack, flags) -> list of dicts like DeviceGraph.probe_numeric.
    fetch(k) -> (closed cycle, sum_cost, sum_time) of candidate k of the LAST probe (DeviceGraph.fetch_cycle_sums);
    when given, probes run with F_NO_CYCLES -- a round usually verifies several negative cycles per rank and only
    the best one of all ranks is ever needed on a host.
    Collective: every rank of `group` must call this with the same lo/hi/max_iter/tol/K_local.
    """
    import torch
    dist = _dist()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    K_total = K_local * world
    dev = _comm_device(group)
    have_cycle = False
    best = None            # (cycle, sum_cost, sum_time) -- only the owner keeps the cycle itself
    best_owner, best_len = -1, 0
    iterations = 0
    probes_local = 0
    it = -1
    flags = N.F_MONOTONE_PRUNE | (N.F_NO_CYCLES if fetch is not None else 0)

    def cycle_of(res, k):
        r = res[k]
        if r.get("cycle") is not None:
            return list(r["cycle"]), r["sum_cost"], r["sum_time"]
        return fetch(k)

    for it in range(max_iter):
        iterations += 1
        lam = N.plan_candidates(lo, hi, have_cycle, tol, K_total)
        mine = lam[rank * K_local:(rank + 1) * K_local]
        res = probe(mine, slack, flags)
        probes_local += len(mine)
        rec = torch.zeros(K_local, 3, dtype=torch.float64)
        for k, r in enumerate(res):
            rec[k, 0] = float(r["verdict"])
            neg = r["verdict"] in (N.V_NEG, N.V_TIE)
            rec[k, 1] = (r["sum_cost"] / r["sum_time"]) if neg else float("nan")
            rec[k, 2] = float(r.get("cycle_len") or len(r.get("cycle") or ())) if neg else 0.0
        rec = rec.to(dev)
        allrec = torch.empty(world * K_local, 3, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allrec, rec, group=group)
        allrec = allrec.cpu().numpy()
        verdict = allrec[:, 0].astype(np.int32)
        ratios = allrec[:, 1].copy()
        lo, hi, have_cycle, best_idx = N.reduce_round(lam, verdict, ratios, lo, hi, have_cycle)
        if best_idx >= 0:
            best_owner, best_len = best_idx // K_local, int(allrec[best_idx, 2])
            if best_owner == rank:
                best = cycle_of(res, best_idx - rank * K_local)   # before the next probe overwrites it
        if hi - lo <= tol:
            break
    if best_owner < 0:
        # one more try at the upper bound (solver.py:1079-1085); every rank probes, rank 0 owns it
        res = probe(np.array([hi]), slack, flags & ~N.F_MONOTONE_PRUNE)
        probes_local += 1
        found = torch.zeros(1, dtype=torch.float64)
        if rank == 0 and res[0]["verdict"] == N.V_NEG:
            best = cycle_of(res, 0)
            found[0] = float(len(best[0]))
        found = found.to(dev)                    # ranks may extract different cycles: rank 0's is the one, and its
        dist.broadcast(found, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)   # length
        if int(found.cpu().item()) > 0:
            best_owner, best_len = 0, int(found.cpu().item())
    if best_owner < 0:
        return ShardedResult([], 0.0, 0.0, float("inf"), iterations, lo, hi, N.ERR_NO_CYCLE, probes_local)
    # one broadcast of the winning cycle from its owner: [sum_cost, sum_time, v_0 .. v_L-1]; every rank knows L from
    # the records of the round that produced it
    msg = torch.zeros(2 + best_len, dtype=torch.float64)
    if rank == best_owner:
        msg[0], msg[1] = best[1], best[2]
        msg[2:] = torch.tensor(best[0], dtype=torch.float64)
    msg = msg.to(dev)
    src = dist.get_global_rank(group, best_owner) if group is not None else best_owner
    dist.broadcast(msg, src=src, group=group)
    msg = msg.cpu()
    sc, st = float(msg[0].item()), float(msg[1].item())
    cyc = [int(x) for x in msg[2:].tolist()]
    rc = N.ERR_CONVERGENCE if (it == max_iter - 1 and hi - lo > tol) else N.OK
    return ShardedResult(cyc, sc, st, sc / st, iterations, lo, hi, rc, probes_local)


def shard_units(n_units: int, rank: int, world: int) -> range:
    """Independent-unit sharding: contiguous block of graphs / scenarios for this rank."""
    base, extra = divmod(n_units, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def sensitivity_analysis_sharded(solver, edge_perturbations, group=None):
    """
    BASELINE config 5 over several GPUs, one process per GPU: every rank holds the same solver
    (same graph, already solved once so that the warm-start cycle is the same everywhere), evaluates the contiguous
    block shard_units(len(scenarios), rank, world) with MinRatioCycleSolver.sensitivity_analysis and one
    all_gather_object returns the full result list on every rank.  No collective on the data path.
    """
    dist = _dist()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    mine = shard_units(len(edge_perturbations), rank, world)
    local = solver.sensitivity_analysis([edge_perturbations[i] for i in mine]) if len(mine) else []
    # Almost every scenario keeps the unperturbed optimal cycle: those travel as (sum_cost, sum_time) pairs in one array and
    # are rebuilt on arrival; only the few results with another cycle (or a failure) travel as objects.  Pickling 1,024
    # SolverResult objects with their configs cost more than the re-solves themselves.
    base = solver._last_result
    base_cycle = list(base.cycle) if base is not None and base.success else None
    plain = [r.success and base_cycle is not None and r.cycle == base_cycle and r.metrics is not None and r.metrics.iterations == 1
             for r in local]
    sums = np.array([[r.sum_cost, r.sum_time] if p else [np.nan, np.nan] for r, p in zip(local, plain)], dtype=np.float64).reshape(-1, 2)
    rest = {i: r for i, (r, p) in enumerate(zip(local, plain)) if not p}
    gathered = [None] * world
    dist.all_gather_object(gathered, (sums.tobytes(), rest), group=group)
    from .solver import SolverMetrics, SolverResult
    out = []
    for blob, others in gathered:
        part = np.frombuffer(blob, dtype=np.float64).reshape(-1, 2)
        for i in range(len(part)):
            if i in others:
                out.append(others[i])
            else:
                sc, st = float(part[i, 0]), float(part[i, 1])
                out.append(SolverResult(cycle=list(base_cycle), sum_cost=sc, sum_time=st, ratio=sc / st, success=True,
                                        metrics=SolverMetrics(iterations=1, mode_used="numeric"), config_used=solver.config))
    return out
