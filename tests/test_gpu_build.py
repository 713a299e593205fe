"""
Device-side array build (mrc_graph_build), host mirrors on demand and cycle weights on the device (-m gpu):
SURVEY 8f rows 1-2 -- _preprocess_graph / _remove_dominated_edges / _build_arrays_if_needed (solver.py:450-575) and
_compute_cycle_weights_float / _verify_cycle_exists (:1296-1327, :1563-1582).  Bit-exact against the NumPy path
(which is checked against the oracle and the reference fixtures in test_abi_and_host.py / test_gpu_parity.py).
"""

from __future__ import annotations

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    import min_ratio_cycle_b200.solver as S
    from min_ratio_cycle_b200 import _native as N
    assert N.device_count() > 0
    return S, N


def _graph_with_parallel_edges(n, m, seed, ndup):
    """random simple digraph + ndup copies of existing pairs, some dominated, some not, some exact duplicates"""
    from min_ratio_cycle_b200 import synthetic
    _, U, V, C, T = synthetic.random_float(n, m, seed=seed)
    rng = np.random.default_rng(seed + 100)
    pick = rng.integers(0, m, ndup)
    kind = rng.integers(0, 4, ndup)
    dc = np.where(kind == 0, 0.5, np.where(kind == 1, -0.5, np.where(kind == 2, 0.5, 0.0)))    # worse / better / mixed / equal
    dt = np.where(kind == 0, 0.25, np.where(kind == 1, -0.01, np.where(kind == 2, -0.01, 0.0)))
    U2, V2 = np.concatenate([U, U[pick]]), np.concatenate([V, V[pick]])
    C2, T2 = np.concatenate([C, C[pick] + dc]), np.concatenate([T, np.maximum(T[pick] + dt, 0.05)])
    p = rng.permutation(len(U2))
    return U2[p], V2[p], C2[p], T2[p]


@pytest.mark.parametrize("n,m,ndup", [(50_000, 400_000, 0), (30_000, 300_000, 5000), (1000, 250_000, 20_000)])
def test_device_build_equals_numpy_path(mods, n, m, ndup):
    S, N = mods
    U, V, C, T = _graph_with_parallel_edges(n, m, 3, ndup)
    drop = S.dominated_mask(n, U, V, C, T)                      # host restatement of :531-575 (checked against the oracle)
    keep = np.ones(len(U), bool) if drop is None else ~drop
    Uk, Vk, Ck, Tk = U[keep], V[keep], C[keep], T[keep]
    order = np.argsort(Vk, kind="stable")                      # :476-478
    g, dups, removed = N.DeviceGraph.build(n, U, V, C, T, remove_dominated=True)
    assert removed == int((~keep).sum()) and g.m == int(keep.sum())
    assert (dups > 0) == (ndup > 0)
    so, do, co, to, oo, kk = g.fetch_arrays(order=True, keep=True)
    assert np.array_equal(kk.astype(bool), keep)
    assert np.array_equal(oo, order) and np.array_equal(so, Uk[order]) and np.array_equal(do, Vk[order])
    assert np.array_equal(co, Ck[order]) and np.array_equal(to, Tk[order])
    # the same without preprocessing, int32 ids
    g2, _, rem2 = N.DeviceGraph.build(n, U.astype(np.int32), V.astype(np.int32), C, T, remove_dominated=False)
    o2 = np.argsort(V, kind="stable")
    s2, d2, c2, t2, oo2 = g2.fetch_arrays(order=True)
    assert rem2 == 0 and g2.m == len(U) and np.array_equal(oo2, o2) and np.array_equal(s2, U[o2]) and np.array_equal(c2, C[o2])
    # both graphs solve to the same optimum as a graph created from the host-sorted arrays
    ref = N.DeviceGraph(n, Uk[order], Vk[order], Ck[order], Tk[order])
    lo, hi = ref.bounds()
    a, b = ref.solve_numeric(lo, hi, K=1), g.solve_numeric(lo, hi, K=1)
    assert a["rc"] == b["rc"] and abs(a["ratio"] - b["ratio"]) <= 1e-12 * max(1.0, abs(a["ratio"]))
    for x in (g, g2, ref):
        x.close()


def test_device_build_rejects_bad_input(mods):
    S, N = mods
    n = 1000
    U = np.arange(300_000) % n
    V = (np.arange(300_000) * 7 + 1) % n
    C = np.ones(300_000)
    T = np.ones(300_000)
    bad = V.copy(); bad[12345] = n
    with pytest.raises(N.NativeError):
        N.DeviceGraph.build(n, U, bad, C, T)
    Tb = T.copy(); Tb[777] = 0.0
    with pytest.raises(N.NativeError):
        N.DeviceGraph.build(n, U, V, C, Tb)
    # ... and through the facade these are the reference's ValueErrors (raised when the arrays are built)
    s = S.MinRatioCycleSolver(n, S.SolverConfig(log_level=S.LogLevel.NONE))
    s.add_edges_arrays(U, V, C + 0.5, Tb)
    with pytest.raises(ValueError, match="time must be strictly positive"):
        s.solve()
    s = S.MinRatioCycleSolver(n, S.SolverConfig(log_level=S.LogLevel.NONE))
    s.add_edges_arrays(U, bad, C + 0.5, T)
    with pytest.raises(ValueError, match="valid vertex indices"):
        s.solve()


def test_facade_fast_path_matches_host_path(mods, oracle_mod, monkeypatch):
    """add_edges_arrays + solve() through mrc_graph_build (mirrors fetched lazily) == the NumPy build path."""
    S, N = mods
    n = 40_000
    U, V, C, T = _graph_with_parallel_edges(n, 400_000, 11, 3000)
    cfg = S.SolverConfig(log_level=S.LogLevel.NONE)
    fast = S.MinRatioCycleSolver(n, cfg)
    fast.add_edges_arrays(U, V, C, T)
    rf = fast.solve()
    assert fast._mirrors_pending, "the fast path must not have fetched the host mirrors for a plain solve()"
    monkeypatch.setattr(S, "_DEVICE_BUILD_MIN_EDGES", 10**12)
    slow = S.MinRatioCycleSolver(n, cfg)
    slow.add_edges_arrays(U, V, C, T)
    rs = slow.solve()
    assert rf.cycle == rs.cycle and rf.sum_cost == rs.sum_cost and rf.sum_time == rs.sum_time and rf.ratio == rs.ratio
    assert rf.metrics.graph_properties["edges"] == rs.metrics.graph_properties["edges"] == fast._num_edges() == slow._num_edges()
    for name in ("_U", "_V", "_C", "_T", "_starts", "_counts", "_order"):
        assert np.array_equal(getattr(fast, name), getattr(slow, name)), name
    assert not fast._mirrors_pending
    o = oracle_mod.Oracle(n, U, V, C, T, all_int=False, threads=8)
    o.set_quirks(False)
    ok, why = o.certify(rf.cycle, rf.ratio)
    assert ok, why
    # sensitivity analysis and edge removal keep working on a device-built graph
    res = fast.sensitivity_analysis([{0: (0.5, 0.0)}, {7: (-0.5, 0.0)}])
    assert len(res) == 2 and all(r.success for r in res)
    u0, v0 = int(U[0]), int(V[0])
    before = fast._num_edges()
    assert fast.remove_edge(u0, v0) and fast._num_edges() == before - 1
    assert fast.solve().success


def test_cycle_weights_on_device(mods):
    S, N = mods
    from min_ratio_cycle_b200 import synthetic
    n, U, V, C, T = synthetic.random_float(5000, 60_000, seed=2)
    # parallel edges on the hops of a known cycle: the hop rule picks the first edge of minimal cost/time
    s = S.MinRatioCycleSolver(n, S.SolverConfig(log_level=S.LogLevel.NONE, enable_preprocessing=False))
    s.add_edges_arrays(U, V, C, T)
    r = s.solve()
    hops = list(zip(r.cycle[:-1], r.cycle[1:]))
    extra = [(u, v, 100.0, 1.0) for u, v in hops] + [(u, v, -1e-3, 50.0) for u, v in hops[:2]]
    for u, v, c, t in extra:
        s.add_edge(int(u), int(v), c, t)
    s._build_arrays_if_needed()
    host = s._compute_cycle_weights_float(r.cycle[:-1])
    sc, st, missing, he = s._dev.cycle_weights(r.cycle)
    assert missing == -1 and (sc, st) == host
    assert all(int(s._U[e]) == u and int(s._V[e]) == v for e, (u, v) in zip(he, hops))
    bogus = list(r.cycle)
    bogus[1] = (bogus[1] + 1) % n
    assert s._dev.cycle_weights(bogus)[2] in (0, 1)
