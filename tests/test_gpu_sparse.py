"""
GPU tests of the sparse (worklist) relaxation passes (option "sparse", relax_sparse_kernel): verdicts, ratios,
exact potentials and exact fractions must be the same as with dense sweeps only; certified by the CPU oracle.
The edge-count floor below which the library keeps sweeping densely is lifted (option "sparse_min_edges").
"""

from __future__ import annotations

from fractions import Fraction

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _sorted(n, U, V, C, T):
    o = np.argsort(V, kind="stable")
    return n, U[o], V[o], C[o], T[o]


def _graph(N, args, sparse, integer=False, **opts):
    n, U, V, C, T = args
    kw = dict(Ci=C.astype(np.int64), Ti=T.astype(np.int64)) if integer else {}
    g = N.DeviceGraph(n, U, V, C, T, **kw)
    g.set_option("sparse", 1 if sparse else 0)
    g.set_option("sparse_min_edges", 0)          # the default floor (1.5 M edges) is a performance choice, not a limit
    for k, v in opts.items():
        g.set_option(k, v)
    return g


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("K", [1, 4, 8])
def test_numeric_solve_same_as_dense(seed, K, oracle_mod):
    from min_ratio_cycle_b200 import _native as N, synthetic
    args = _sorted(*synthetic.random_float(20_000, 200_000, seed=seed))
    out = {}
    for sparse in (0, 1):
        g = _graph(N, args, sparse)
        lo, hi = g.bounds()
        g.reset_stats()
        out[sparse] = g.solve_numeric(lo, hi, K=K)
        st = g.stats()
        assert out[sparse]["rc"] == 0
        assert (st["sparse_passes"] > 0) == bool(sparse), st
        g.close()
    assert abs(out[0]["ratio"] - out[1]["ratio"]) <= 1e-9 * max(1.0, abs(out[0]["ratio"]))
    o = oracle_mod.Oracle(*args, all_int=False)
    o.set_quirks(False)
    ok, why = o.certify(out[1]["cycle"], out[1]["ratio"])
    assert ok, why


@pytest.mark.parametrize("div", [1, 4, 32])
def test_probe_verdicts_and_thresholds(div):
    """Every switch threshold (div=1: sparse from the first burst boundary on) gives the dense verdicts."""
    from min_ratio_cycle_b200 import _native as N, synthetic
    args = _sorted(*synthetic.random_float(30_000, 300_000, seed=7))
    g = _graph(N, args, 0)
    lo, hi = g.bounds()
    star = g.solve_numeric(lo, hi, K=4)["ratio"]
    lams = [star - 1.0, star - 1e-3, star - 1e-9, star + 1e-6, star + 0.5, star + 5.0]
    dense = [g.probe_numeric([x])[0] for x in lams]
    g.close()
    g = _graph(N, args, 1, sparse_div=div)
    for x, d in zip(lams, dense):
        p = g.probe_numeric([x])[0]
        assert p["verdict"] == d["verdict"], (x, p["verdict"], d["verdict"])
        if p["verdict"] == N.V_NEG:
            assert p["sum_cost"] / p["sum_time"] < x
    both = g.probe_numeric(lams[:4], flags=1)
    assert all(b["verdict"] in (N.V_NONNEG, N.V_SKIPPED_NONNEG) for b in both[:3]), [b["verdict"] for b in both]
    assert both[3]["verdict"] == N.V_NEG
    g.close()


@pytest.mark.parametrize("seed", [11, 12])
def test_exact_bit_identical_to_dense(seed):
    from min_ratio_cycle_b200 import _native as N, synthetic
    n, U, V, C, T = synthetic.random_int(20_000, 200_000, seed=seed)
    args = _sorted(n, U, V, C.astype(np.float64), T.astype(np.float64))
    res = {}
    for sparse in (0, 1):
        g = _graph(N, args, sparse, integer=True)
        ex = g.solve_exact(strategy=1)
        assert ex["rc"] == 0, ex["message"]
        ok, pot = g.potentials_exact(ex["a"], ex["b"])
        res[sparse] = (ex["a"], ex["b"], ex["cycle"], ex["sum_cost"], ex["sum_time"], ex["iterations"], ex["probes"], ok, pot.copy())
        g.close()
    for i in range(8):
        assert res[0][i] == res[1][i], i
    assert np.array_equal(res[0][8], res[1][8])          # potentials: the unique least fixed point
    assert Fraction(res[1][3], res[1][4]) == Fraction(res[1][0], res[1][1])


def test_scenario_probes_same_as_dense():
    from min_ratio_cycle_b200 import _native as N, synthetic
    args = _sorted(*synthetic.random_float(20_000, 200_000, seed=21))
    n, U, V, C, T = args
    rng = np.random.default_rng(3)
    pe = rng.integers(0, len(U), 4)
    dc = 0.5 * C[pe]; dt = np.zeros(4)
    g = _graph(N, args, 0)
    lo, hi = g.bounds()
    star = g.solve_numeric(lo, hi, K=4)["ratio"]
    lams = [star - 1e-6, star + 1e-3, star - 0.1, star + 0.2]
    dense = g.probe_numeric_scenarios(lams, pe, dc, dt)
    g.close()
    g = _graph(N, args, 1)
    sp = g.probe_numeric_scenarios(lams, pe, dc, dt)
    assert [a["verdict"] for a in sp] == [b["verdict"] for b in dense]
    g.close()


def test_long_tail_and_ring_through_sparse_passes():
    """Deep propagation (one hop per worklist pass): ring of 70,000 vertices plus random chords."""
    from min_ratio_cycle_b200 import _native as N, synthetic
    args = _sorted(*synthetic.ring_plus_random(70_000, 30_000, seed=5))
    r = {}
    for sparse in (0, 1):
        g = _graph(N, args, sparse)
        lo, hi = g.bounds()
        r[sparse] = g.solve_numeric(lo, hi, K=4)
        assert r[sparse]["rc"] == 0
        g.close()
    assert abs(r[0]["ratio"] - r[1]["ratio"]) <= 1e-9 * max(1.0, abs(r[0]["ratio"]))


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_back_to_back_probes_leave_no_stale_worklist_state(seed, oracle_mod):
    """
    A probe that ends on a verified negative cycle leaves a non-empty worklist (and its membership bits) behind; the
    next probe must not inherit them.  Newton iteration (K=1) with the earliest possible switch to worklist passes
    runs 6-12 such probes in a row; the last, certifying one must still reach the true fixed point.
    """
    from min_ratio_cycle_b200 import _native as N, synthetic
    args = _sorted(*synthetic.random_float(50_000, 500_000, seed=seed))
    g = _graph(N, args, 0)
    lo, hi = g.bounds()
    want = g.solve_numeric(lo, hi, K=1)["ratio"]
    g.close()
    for look in (1, 2):
        g = _graph(N, args, 1, sparse_look=look, sparse_div=4)
        for K in (1, 4, 1):
            out = g.solve_numeric(lo, hi, K=K)
            assert out["rc"] == 0
            assert abs(out["ratio"] - want) <= 1e-9 * max(1.0, abs(want)), (look, K, out["ratio"], want)
        # verdicts around the optimum, straight after a negative probe
        for _ in range(3):
            assert g.probe_numeric([want + 1e-3])[0]["verdict"] == N.V_NEG
            assert g.probe_numeric([want - 1e-6])[0]["verdict"] == N.V_NONNEG
        g.close()
    o = oracle_mod.Oracle(*args, all_int=False)
    o.set_quirks(False)
    pr = o.detect_negative_cycle_numeric(want - 1e-7 * max(1.0, abs(want)), 1e-15, True)
    assert pr.has_neg == 0


@pytest.mark.parametrize("K", [1, 4])
def test_membership_bitmaps_are_reset_before_use(K):
    """Garbage in the worklist membership bitmaps (test hook "sparse_poison") must not survive into a probe."""
    from min_ratio_cycle_b200 import _native as N, synthetic
    args = _sorted(*synthetic.random_float(50_000, 500_000, seed=9))
    g = _graph(N, args, 0)
    lo, hi = g.bounds()
    want = g.solve_numeric(lo, hi, K=4)["ratio"]
    g.close()
    g = _graph(N, args, 1, sparse_look=1, sparse_div=2)
    for lam, verdict in ((want - 1e-6, N.V_NONNEG), (want + 1e-4, N.V_NEG), (want - 1e-6, N.V_NONNEG)):
        g.set_option("sparse_poison", 1)
        g.reset_stats()
        p = g.probe_numeric([lam] if K == 1 else [lam - 3.0, lam - 2.0, lam - 1.0, lam])[-1]
        assert g.stats()["sparse_passes"] > 0
        assert p["verdict"] == verdict, (lam, p["verdict"])
    g.set_option("sparse_poison", 1)
    out = g.solve_numeric(lo, hi, K=K)
    assert out["rc"] == 0 and abs(out["ratio"] - want) <= 1e-9 * max(1.0, abs(want))
    g.close()
