"""
GPU tests for the cycle-extraction corner cases of cycle_check_kernel: cycles longer than the capped laps
(2,048 / 4,096 steps), and a representative that sits thousands of predecessor steps away from a short cycle
(the doubled pointers cannot land on it, the capped laps end on the tail, Brent's fallback finishes the job).
Checked against closed-form answers and the CPU oracle's detector (same inputs, through the C ABI).
"""

from __future__ import annotations

from fractions import Fraction

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _sorted_by_dst(U, V, C, T):
    o = np.argsort(V, kind="stable")
    return U[o].astype(np.int32), V[o].astype(np.int32), C[o], T[o]


def ring(n, seed, integer):
    rng = np.random.default_rng(seed)
    U = np.arange(n, dtype=np.int64); V = (U + 1) % n
    if integer:
        C = -rng.integers(1, 50, n).astype(np.float64); T = rng.integers(1, 9, n).astype(np.float64)
    else:
        C = -rng.uniform(0.5, 10, n); T = rng.uniform(0.1, 5, n)
    return U, V, C, T


def lollipop(tail, integer):
    """3-cycle on the three largest ids, and a path leaving it whose far end is vertex 0 (tail -> ... -> 1 -> 0)."""
    L = tail
    U = [L, L + 1, L + 2] + list(range(L, 0, -1))
    V = [L + 1, L + 2, L] + list(range(L - 1, -1, -1))
    C = [-3.0, -2.0, -4.0] + [-1.0] * L
    T = [1.0, 2.0, 1.0] + [1.0] * L
    return (np.array(U, np.int64), np.array(V, np.int64), np.array(C), np.array(T)), L + 3


@pytest.mark.parametrize("n", [3000, 6000])
@pytest.mark.parametrize("K", [1, 4])
def test_ring_longer_than_the_capped_laps(n, K, oracle_mod):
    from min_ratio_cycle_b200 import _native as N
    U, V, C, T = ring(n, seed=n, integer=False)
    s, d, c, t = _sorted_by_dst(U, V, C, T)
    g = N.DeviceGraph(n, s, d, c, t)
    expect = float(np.sum(C)) / float(np.sum(T))
    # every edge is negative at lambda = 0: the predecessor graph after one pass IS the ring
    pr = g.probe_numeric([0.0] * 1 if K == 1 else [-1e6, 0.0, 1.0, 2.0][:K], flags=1)
    neg = [p for p in pr if p["verdict"] == N.V_NEG]
    assert neg, [p["verdict"] for p in pr]
    p0 = neg[0]
    assert p0["cycle_len"] == n + 1 and p0["cycle"][0] == 0 and p0["cycle"][-1] == 0
    assert p0["cycle"][:-1] == list(range(n))
    assert abs(p0["sum_cost"] / p0["sum_time"] - expect) <= 1e-12 * abs(expect)
    lo, hi = g.bounds()
    out = g.solve_numeric(lo, hi, K=K)
    assert out["rc"] == 0 and out["cycle"][:-1] == list(range(n))
    assert abs(out["ratio"] - expect) <= 1e-9 * max(1.0, abs(expect))
    o = oracle_mod.Oracle(n, U, V, C, T, all_int=False)
    o.set_quirks(False)
    ok, why = o.certify(out["cycle"], out["ratio"])
    assert ok, why
    g.close()


@pytest.mark.parametrize("n", [3000, 6000])
def test_ring_exact(n):
    from min_ratio_cycle_b200 import _native as N
    U, V, C, T = ring(n, seed=n + 1, integer=True)
    s, d, c, t = _sorted_by_dst(U, V, C, T)
    g = N.DeviceGraph(n, s, d, c, t, Ci=c.astype(np.int64), Ti=t.astype(np.int64))
    want = Fraction(int(C.sum()), int(T.sum()))
    pr = g.probe_exact([0], [1])[0]          # lambda = 0/1: every edge negative
    assert pr["verdict"] == N.V_NEG and pr["cycle_len"] == n + 1
    assert Fraction(pr["sum_cost"], pr["sum_time"]) == want
    ex = g.solve_exact(strategy=1)
    assert ex["rc"] == 0 and Fraction(ex["a"], ex["b"]) == want, (ex["rc"], ex["message"], ex["a"], ex["b"])
    assert set(ex["cycle"]) == set(range(n))
    assert Fraction(ex["sum_cost"], ex["sum_time"]) == want
    g.close()


@pytest.mark.parametrize("tail", [700, 5000, 9000])
@pytest.mark.parametrize("integer", [False, True])
def test_representative_far_from_a_short_cycle(tail, integer):
    from min_ratio_cycle_b200 import _native as N
    (U, V, C, T), n = lollipop(tail, integer)
    s, d, c, t = _sorted_by_dst(U, V, C, T)
    kw = dict(Ci=c.astype(np.int64), Ti=t.astype(np.int64)) if integer else {}
    g = N.DeviceGraph(n, s, d, c, t, **kw)
    L = tail
    want_cycle = [L, L + 1, L + 2, L]
    if integer:
        pr = g.probe_exact([0], [1])[0]
        assert pr["verdict"] == N.V_NEG and pr["cycle"] == want_cycle, pr
        assert (pr["sum_cost"], pr["sum_time"]) == (-9, 4)
        ex = g.solve_exact(strategy=1)
        assert ex["rc"] == 0 and Fraction(ex["a"], ex["b"]) == Fraction(-9, 4) and set(ex["cycle"]) == set(want_cycle)
    else:
        for K in (1, 4):
            pr = g.probe_numeric([0.0] if K == 1 else [-100.0, -50.0, -10.0, 0.0], flags=1)
            neg = [p for p in pr if p["verdict"] == N.V_NEG]
            assert neg and neg[0]["cycle"] == want_cycle, pr
            assert neg[0]["sum_cost"] == -9.0 and neg[0]["sum_time"] == 4.0
        lo, hi = g.bounds()
        out = g.solve_numeric(lo, hi, K=4)
        assert out["rc"] == 0 and out["cycle"] == want_cycle and out["ratio"] == -2.25
    g.close()
