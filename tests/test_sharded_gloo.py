"""
World-size-2 test of the multi-GPU lambda-sharding driver on CPU with the gloo backend.
The collective logic (candidate split, all-gather of verdicts, bracket folding, cycle broadcast) is the
product code in min_ratio_cycle_b200/sharded.py; the probe is injected and answered by the CPU oracle
(test infrastructure), because there is no GPU here.
"""

from __future__ import annotations

import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_probe_factory(o, N):
    def probe(lams, slack, flags):
        out = []
        for lam in lams:
            r = o.detect_negative_cycle_numeric(float(lam), slack, True)
            if r.has_neg and r.cycle is not None:
                cyc = list(r.cycle)
                k = cyc.index(min(cyc))
                cyc = cyc[k:] + cyc[:k]
                sc, st = o.cycle_weights_float(cyc)
                out.append(dict(lam=float(lam), verdict=N.V_NEG, cycle=cyc + [cyc[0]], sum_cost=sc, sum_time=st, passes=0))
            elif r.has_neg:
                out.append(dict(lam=float(lam), verdict=N.V_NEG_NOCYCLE, cycle=None, sum_cost=0.0, sum_time=0.0, passes=0))
            else:
                out.append(dict(lam=float(lam), verdict=N.V_NONNEG, cycle=None, sum_cost=0.0, sum_time=0.0, passes=0))
        return out
    return probe


def _tick_search(o, N, rank, world, K, lo, hi, tol=1e-12, max_probes=40):
    """
    The tick protocol of mrc_solve_numeric_sharded (csrc/mrc_comm.cuh: shard_solve_rank) enacted on CPU: every rank
    probes K candidates of its own slice of the plan, exchanges (lambda, state, ratio, length) records once per tick,
    folds ALL ranks' records with the library's mrc_shard_fold, abandons what the bracket implies and re-plans without
    waiting for the other rank.  A candidate's verdict (the oracle's) only becomes visible after 1-3 ticks, so the two
    ranks really run out of phase.
    """
    import torch
    import torch.distributed as dist
    have_cycle, probes, ticks = False, 0, 0
    mine, due, best = None, None, None
    rec = np.zeros((K, 4))
    rec[:, 1] = N.V_ABANDONED
    while True:
        if mine is None and probes < max_probes:
            plan = N.plan_candidates(lo, hi, have_cycle, tol, world * K)
            mine = plan[rank * K:(rank + 1) * K]
            due = [1 + (int(abs(x) * 1e6) + rank) % 3 for x in mine]
            rec[:, 0], rec[:, 1], rec[:, 2], rec[:, 3] = mine, -1.0, np.nan, 0.0
            cycles = [None] * K
            probes += 1
        if mine is not None:
            for k in range(K):
                if rec[k, 1] < 0:
                    due[k] -= 1
                    if due[k] <= 0:
                        r = o.detect_negative_cycle_numeric(float(mine[k]), 1e-15, True)
                        if r.has_neg and r.cycle is not None:
                            cyc = list(r.cycle)
                            sc, st = o.cycle_weights_float(cyc)
                            rec[k, 1:] = (N.V_NEG, sc / st, len(cyc))
                            cycles[k] = (cyc + [cyc[0]], sc, st)
                        else:
                            rec[k, 1] = N.V_NEG_NOCYCLE if r.has_neg else N.V_NONNEG
        allrec = [torch.zeros(K, 4, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(allrec, torch.from_numpy(rec.copy()))
        allrec = torch.cat(allrec).numpy()
        ticks += 1
        lo, hi, have_cycle, bi, implied = N.shard_fold(allrec, lo, hi, have_cycle)
        if bi >= 0:
            best = cycles[bi - rank * K] if bi // K == rank else ("other", bi // K)
        if mine is not None:
            for k in range(K):
                if implied[rank * K + k]:
                    rec[k, 1] = N.V_SKIPPED_NONNEG if rec[k, 0] <= lo else N.V_SKIPPED_NEG
            if not (rec[:, 1] < 0).any():
                mine = None
        if hi - lo <= tol or ticks > 400:
            break
    owner = torch.tensor([rank if isinstance(best, tuple) and best[0] != "other" else -1])
    owners = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(owners, owner)
    return lo, hi, ticks, probes, best, [int(x) for x in owners]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from min_ratio_cycle_b200 import _native as N
        from min_ratio_cycle_b200 import sharded, synthetic
        from oracle import oracle
        res = {}
        for seed in (1, 2):
            n, U, V, C, T = synthetic.ring_plus_random(40, 120, seed=seed)
            o = oracle.Oracle(n, U, V, C, T, all_int=False)
            o.set_quirks(False)
            lo, hi = o.bounds()
            r = sharded.solve_numeric_sharded(_oracle_probe_factory(o, N), lo, hi, K_local=3)
            ok, why = o.certify(r.cycle, r.ratio)
            res[seed] = (r.rc, r.cycle, r.ratio, r.iterations, r.lo, r.hi, ok, why)
            # the product path: probes keep their cycles (F_NO_CYCLES), the owner of the best one fetches it afterwards
            full = _oracle_probe_factory(o, N)
            stash = {}

            def lean(lams, slack, flags, _full=full, _stash=stash):
                assert flags & N.F_NO_CYCLES
                out = _full(lams, slack, flags)
                _stash.clear()
                for k, d in enumerate(out):
                    if d["cycle"] is not None:
                        _stash[k] = (d["cycle"], d["sum_cost"], d["sum_time"])
                        d["cycle_len"] = len(d["cycle"])
                        d["cycle"] = None
                return out

            r2 = sharded.solve_numeric_sharded(lean, lo, hi, K_local=3, fetch=lambda k, _stash=stash: _stash[k])
            res[("lean", seed)] = (r2.rc, r2.cycle, r2.ratio, r2.iterations) == (r.rc, r.cycle, r.ratio, r.iterations)
            # the tick protocol of the in-library driver (ranks out of phase, records folded by mrc_shard_fold)
            tlo, thi, ticks, probes, best, owners = _tick_search(o, N, rank, world, 2, lo, hi)
            mine_ok = True
            if isinstance(best, tuple) and best[0] != "other":
                okc, whyc = o.certify(best[0], best[1] / best[2])
                mine_ok = okc and abs(best[1] / best[2] - thi) <= 1e-12
            res[("tick", seed)] = (tlo, thi, ticks, sorted(x for x in owners if x >= 0), abs(thi - r.ratio) <= 1e-9, mine_ok)
        # acyclic graph: every rank must agree on NO_CYCLE
        U = np.array([0, 1, 2]); V = np.array([1, 2, 3])
        o = oracle.Oracle(4, U, V, np.array([1.5, -2.0, 0.5]), np.array([1.0, 1.0, 2.0]), all_int=False)
        o.set_quirks(False)
        lo, hi = o.bounds()
        r = sharded.solve_numeric_sharded(_oracle_probe_factory(o, N), lo, hi, K_local=2)
        res["acyclic"] = r.rc
        res["units"] = list(sharded.shard_units(11, rank, world))
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_lambda_sharding_world2():
    from min_ratio_cycle_b200 import build
    build.build()
    from oracle import oracle
    oracle.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from min_ratio_cycle_b200 import _native as N
    for seed in (1, 2):
        a, b = got[0][seed], got[1][seed]
        assert a == b, "ranks disagree"
        rc, cycle, ratio, iters, lo, hi, ok, why = a
        assert rc == N.OK and ok, why
        assert hi - lo <= 1e-12 and cycle[0] == cycle[-1]
        assert iters <= 12
        assert got[0][("lean", seed)] and got[1][("lean", seed)], "fetch-after-reduce path differs from the plain one"
        ta, tb = got[0][("tick", seed)], got[1][("tick", seed)]
        assert ta[:4] == tb[:4], "ranks disagree on the bracket / tick count / owner of the best cycle"
        assert ta[1] - ta[0] <= 1e-12 and len(ta[3]) == 1 and ta[4] and tb[4] and ta[5] and tb[5], (ta, tb)
    assert got[0]["acyclic"] == got[1]["acyclic"] == N.ERR_NO_CYCLE
    assert got[0]["units"] + got[1]["units"] == list(range(11))
