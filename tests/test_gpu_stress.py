"""
Randomised differential tests on the GPU (run with -m gpu): many seeded graphs of different shapes, every answer
checked by the CPU oracle.

  numeric : certificate C(r) of SURVEY 8c from the quirk-free oracle detector (cycle exists, recomputed ratio equals
            the reported one, no negative cycle at r - 1e-7*max(1,|r|)), for K in {1,2,4,8}
  exact   : bit-exact against the oracle's restatement of _solve_exact (cycle list, integer sums, reduced fraction,
            Stern-Brocot probe sequence, iteration count), incl. graphs with many tied ratios (small integer weights)
            where only the reference's DFS order decides which zero cycle is reported
"""

from __future__ import annotations

from fractions import Fraction

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    import min_ratio_cycle_b200.solver as mod
    from min_ratio_cycle_b200 import _native as N
    assert N.device_count() > 0
    return mod


def _graph(kind, rng, integer):
    from min_ratio_cycle_b200 import synthetic
    if kind == "ring_chords":            # one long cycle + a few chords: long optimal cycles, deep propagation
        n = int(rng.integers(20, 400))
        n_, U, V, C, T = synthetic.ring_plus_random(n, int(rng.integers(0, n // 2 + 1)), seed=int(rng.integers(1 << 30)), integer=integer)
    elif kind == "sparse":
        n = int(rng.integers(10, 600)); m = int(n * rng.uniform(1.5, 6))
        n_, U, V, C, T = (synthetic.random_int if integer else synthetic.random_float)(n, max(m, n), seed=int(rng.integers(1 << 30)))
    elif kind == "dense":
        n = int(rng.integers(5, 40)); m = int(n * (n - 1) * rng.uniform(0.4, 0.95))
        n_, U, V, C, T = (synthetic.random_int if integer else synthetic.random_float)(n, max(m, n), seed=int(rng.integers(1 << 30)))
    else:                                # "ties": tiny integer weights -> many cycles share the optimal ratio
        n = int(rng.integers(6, 120)); m = int(n * rng.uniform(2, 5))
        n_, U, V, _, _ = synthetic.random_int(n, max(m, n), seed=int(rng.integers(1 << 30)))
        C = rng.integers(-3, 4, len(U)).astype(np.float64); T = rng.integers(1, 4, len(U)).astype(np.float64)
    return n_, U, V, C, T


@pytest.fixture(params=["default", "sparse_forced", "persist_all", "persist_all_sparse", "burst_only"])
def relax_mode(request, monkeypatch):
    """default: small graphs sweep densely; single-candidate probes run as ONE persistent probe kernel, wider ones through
    the burst driver.  sparse_forced: worklist passes from the first burst boundary on, on every graph (environment
    switches read at graph creation), so the sparse kernel sees ties, parallel edges, self-loops.  persist_all: every probe
    of up to 4 candidates in the persistent kernel (narrowing, compact survivor, pruning and check schedule on the device),
    with and without forced worklist passes.  burst_only: no persistent kernel at all."""
    if request.param in ("sparse_forced", "persist_all_sparse"):
        monkeypatch.setenv("MRC_B200_SPARSE_MIN_EDGES", "0")
        monkeypatch.setenv("MRC_B200_SPARSE_DIV", "1")
    if request.param.startswith("persist_all"):
        monkeypatch.setenv("MRC_B200_PERSIST", "1")
    if request.param == "burst_only":
        monkeypatch.setenv("MRC_B200_PERSIST", "0")
    monkeypatch.setenv("MRC_B200_COMPACT_MIN_EDGES", "0")   # compact-survivor path on every K >= 2 probe that narrows to one slot
    return request.param


def test_numeric_random_graphs_certified(S, oracle_mod, relax_mode):
    rng = np.random.default_rng(20261020)
    kinds = ["ring_chords", "sparse", "dense", "ties"]
    done = 0
    for i in range(160):
        kind = kinds[i % 4]
        integer = kind == "ties" or (i % 8) >= 4
        n, U, V, C, T = _graph(kind, rng, integer)
        K = int(rng.choice([1, 2, 4, 8]))
        s = S.MinRatioCycleSolver(n, S.SolverConfig(log_level=S.LogLevel.NONE, gpu_candidates=K))
        s.add_edges_arrays(U, V, C, T)
        r = s.solve()
        o = oracle_mod.Oracle(n, U, V, C, T, all_int=integer)
        o.set_quirks(False)
        ok, why = o.certify(r.cycle, r.ratio)
        assert ok, (i, kind, n, len(U), K, why)
        assert r.cycle[0] == r.cycle[-1] and abs(r.ratio - r.sum_cost / r.sum_time) < 1e-10
        done += 1
    assert done == 160


def test_exact_random_graphs_bit_exact(S, oracle_mod, relax_mode):
    rng = np.random.default_rng(7)
    kinds = ["ties", "sparse", "ring_chords", "dense"]
    checked = 0
    for i in range(120):
        kind = kinds[i % 4]
        n, U, V, C, T = _graph(kind, rng, True)
        o = oracle_mod.Oracle(n, U, V, C, T, all_int=True)
        try:
            x = o.solve_exact()
        except oracle_mod.OracleError as e:
            if e.kind == "reduceat":     # reference crash, not a result
                continue
            x = e
        for strategy in (1, 0):
            s = S.MinRatioCycleSolver(n, S.SolverConfig(log_level=S.LogLevel.NONE, gpu_exact_strategy=strategy))
            s.add_edges_arrays(U, V, C, T)
            s._preprocess_graph()
            s._build_arrays_if_needed()
            r = s._solve_exact()
            if isinstance(x, Exception):
                assert not r.success, (i, kind, strategy)
                continue
            assert r.success, (i, kind, strategy, r.error_message)
            assert r.cycle == x.cycle, (i, kind, strategy)
            assert (int(r.sum_cost), int(r.sum_time)) == (x.sum_cost, x.sum_time)
            assert Fraction(int(r.sum_cost), int(r.sum_time)) == x.fraction
            assert s._last_exact["probes"] == x.probes and (s._last_exact["a"], s._last_exact["b"]) == x.ab
            if x.iterations >= 0:
                assert s._metrics.iterations == x.iterations
        checked += 1
    assert checked >= 100
