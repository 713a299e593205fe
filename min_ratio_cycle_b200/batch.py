"""
Batches of independent small graphs (BASELINE config 4: 4,096 currency-arbitrage graphs, n=64, m=4,032).

The reference has no batch entry point: a caller loops `solve_min_ratio_cycle(edges)` (solver.py:1941-1969),
one Python solve per graph.  Here the whole batch goes to the GPU in one call (mrc_batch_solve_numeric: one
CTA per graph, the complete lambda search on the device), and across GPUs the graphs -- independent units,
SURVEY.md 8e (2) -- are split in contiguous blocks over the ranks with no data-path collective; the results
are gathered once at the end.
"""

from __future__ import annotations

import numpy as np

from . import _native as N
from .sharded import shard_units
from .solver import SolverConfig, SolverMetrics, SolverResult, dominated_mask


def _prepare(graphs, preprocess: bool):
    nv, off, S, D, Cs, Ts = [], [0], [], [], [], []
    for n, U, V, Cw, Tw in graphs:
        U = np.asarray(U, dtype=np.int64); V = np.asarray(V, dtype=np.int64)
        Cw = np.asarray(Cw, dtype=np.float64); Tw = np.asarray(Tw, dtype=np.float64)
        if len(U) == 0:
            raise ValueError("Graph has no edges")
        if U.min() < 0 or V.min() < 0 or U.max() >= n or V.max() >= n:
            raise ValueError("u and v must be valid vertex indices")
        if not np.all(Tw > 0):
            raise ValueError("Edge transit time must be strictly positive")
        if preprocess:
            drop = dominated_mask(int(n), U, V, Cw, Tw)          # solver.py:531-575
            if drop is not None:
                k = ~drop
                U, V, Cw, Tw = U[k], V[k], Cw[k], Tw[k]
        o = np.argsort(V, kind="stable")                         # solver.py:476-478
        nv.append(int(n)); off.append(off[-1] + len(U))
        S.append(U[o].astype(np.int32)); D.append(V[o].astype(np.int32)); Cs.append(Cw[o]); Ts.append(Tw[o])
    return (np.array(nv, np.int32), np.array(off, np.int64), np.concatenate(S), np.concatenate(D),
            np.concatenate(Cs), np.concatenate(Ts))


def _cycle_weights_float(closed_cycle, src, dst, cost, time):
    sc = st = 0.0
    for u, v in zip(closed_cycle[:-1], closed_cycle[1:]):
        a, b = np.searchsorted(dst, v, "left"), np.searchsorted(dst, v, "right")
        hit = a + np.flatnonzero(src[a:b] == u)
        k = hit[int(np.argmin(cost[hit] / time[hit]))] if hit.size > 1 else hit[0]
        sc += float(cost[k]); st += float(time[k])
    return sc, st


def solve_min_ratio_cycle_batch(graphs, config: SolverConfig | None = None, return_stats: bool = False):
    """
    graphs: sequence of (n_vertices, U, V, cost, time) arrays in insertion order.
    Returns one SolverResult per graph (success=False with the reference's message when a graph has no cycle).
    """
    cfg = config or SolverConfig()
    nv, off, S, D, Cs, Ts = _prepare(graphs, cfg.enable_preprocessing)
    out = N.batch_solve_numeric(nv, off, S, D, Cs, Ts, max_iter=cfg.numeric_max_iter, tol=cfg.numeric_tolerance,
                                slack=cfg.numeric_cycle_slack, K=cfg.gpu_candidates, device=cfg.gpu_device)
    results = []
    for g in range(len(nv)):
        st = int(out["status"][g])
        if st != N.ERR_NO_CYCLE:
            # reported sums follow the reference's hop rule (_compute_cycle_weights_float, solver.py:1296-1327):
            # among parallel (u,v) edges the first with minimal cost/time, accumulated in cycle order
            lo, hi = int(off[g]), int(off[g + 1])
            sc, stt = _cycle_weights_float(out["cycles"][g], S[lo:hi], D[lo:hi], Cs[lo:hi], Ts[lo:hi])
            out["sum_cost"][g], out["sum_time"][g], out["ratio"][g] = sc, stt, sc / stt
        if st == N.ERR_NO_CYCLE:
            results.append(SolverResult([], 0, 0, float("inf"), success=False, error_message="No negative cycle found"))
            continue
        r = SolverResult(cycle=[int(x) for x in out["cycles"][g]], sum_cost=float(out["sum_cost"][g]),
                         sum_time=float(out["sum_time"][g]), ratio=float(out["ratio"][g]), success=True,
                         metrics=SolverMetrics(iterations=int(out["iterations"][g]), mode_used="numeric"),
                         config_used=cfg)
        if st == N.ERR_CONVERGENCE:
            r.success, r.error_message = False, "Numeric solver failed to converge"
        results.append(r)
    return (results, out["stats"]) if return_stats else results


def solve_min_ratio_cycle_batch_sharded(graphs, config: SolverConfig | None = None, group=None):
    """
    One process per GPU: rank r solves graphs shard_units(B, r, world); a single all_gather_object at the end
    returns the full result list on every rank.  No collective on the data path.
    """
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    mine = shard_units(len(graphs), rank, world)
    local = solve_min_ratio_cycle_batch([graphs[i] for i in mine], config) if len(mine) else []
    gathered = [None] * world
    dist.all_gather_object(gathered, local, group=group)
    return [r for part in gathered for r in part]
