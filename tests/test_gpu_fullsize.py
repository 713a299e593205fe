"""
Full-size parity at the BASELINE configs (run with -m gpu).  The Python reference cannot finish these sizes
(n passes per negative probe), so parity is established through size-independent properties
checked by the CPU oracle (which restates the reference and is pinned on the small fixtures):

  C2  exact, n=100k m=1M ints : at the returned lambda* = p/q the ORACLE's _has_negative_cycle_exact is False,
                                and the ORACLE's _find_zero_cycle_exact(p, q) -- potentials + tight-subgraph DFS,
                                cheap because relaxation converges in tens of passes at lambda* -- returns the
                                SAME cycle and the SAME integer sums, bit for bit; Fraction equality follows.
                                Both search strategies (probe-every-step / Newton+replay) agree on everything.
  C3  numeric, n=1M m=10M f64 : the returned cycle exists, its recomputed ratio equals the reported one, and the
                                oracle's detector finds NO negative cycle at ratio - 1e-7*max(1,|ratio|)
                                (certificate C(r) of SURVEY 8c); K = 1, 4, 8, with and without worklist passes,
                                agree to 1e-12.
  C5  sensitivity, n=100k m=2M: 16 perturbation re-solves each certified the same way against the perturbed graph.
"""

from __future__ import annotations

from fractions import Fraction

import numpy as np
import pytest

from parity import cycle_is_valid

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    import min_ratio_cycle_b200.solver as mod
    from min_ratio_cycle_b200 import _native as N
    assert N.device_count() > 0
    return mod


def test_c2_exact_full_size(S, oracle_mod):
    from min_ratio_cycle_b200 import synthetic
    n, U, V, C, T = synthetic.random_int(100_000, 1_000_000, seed=0)
    outs = []
    for strategy in (1, 0):
        s = S.MinRatioCycleSolver(n, S.SolverConfig(log_level=S.LogLevel.NONE, gpu_exact_strategy=strategy))
        s.add_edges_arrays(U, V, C, T)
        cyc, frac = s.solve_exact_fraction()
        outs.append((cyc, frac, s._metrics.iterations, s._last_exact["probes"], s._last_exact["a"], s._last_exact["b"]))
    assert outs[0] == outs[1], "Newton+replay and probe-every-step disagree"
    cyc, frac, iters, probes, a, b = outs[0]
    assert Fraction(a, b) == frac and cycle_is_valid(cyc, U, V)
    o = oracle_mod.Oracle(n, U, V, C, T, all_int=True, threads=8)
    assert o.has_negative_cycle_exact(a, b) is False
    z = o.find_zero_cycle_exact(a, b)
    assert z is not None
    zc, sc, st = z
    assert zc + [zc[0]] == cyc, "cycle differs from the reference DFS order"
    assert Fraction(sc, st) == frac and (sc, st) == (int(s._last_exact["sum_cost"]), int(s._last_exact["sum_time"]))
    # one Stern-Brocot step to the right of lambda* is negative (cheap to see on the GPU, n passes on the CPU)
    assert s._has_negative_cycle_exact(a * 1000 + 1, b * 1000) is True
    print(f"C2: lambda* = {frac}, cycle length {len(cyc) - 1}, Stern-Brocot iterations {iters}, probes {len(probes)}")


def test_c3_numeric_full_size(S, oracle_mod):
    from min_ratio_cycle_b200 import synthetic
    n, U, V, C, T = synthetic.random_float(1_000_000, 10_000_000, seed=0)
    ratios = []
    for K, sparse in ((4, True), (1, True), (8, True), (4, False), (1, False)):   # worklist passes on (facade default) / off
        s = S.MinRatioCycleSolver(n, S.SolverConfig(log_level=S.LogLevel.NONE, gpu_candidates=K, gpu_sparse=sparse,
                                                      enable_preprocessing=False))
        s.add_edges_arrays(U, V, C, T)
        r = s.solve(mode="numeric")
        ratios.append(r.ratio)
        assert r.cycle[0] == r.cycle[-1] and cycle_is_valid(r.cycle, U, V)
        assert abs(r.ratio - r.sum_cost / r.sum_time) < 1e-10 and r.sum_time > 0
        if K == 4 and sparse:
            keep = r
        s._dev.close()
    assert max(ratios) - min(ratios) <= 1e-12
    o = oracle_mod.Oracle(n, U, V, C, T, all_int=False, preprocess=False, threads=8)
    o.set_quirks(False)
    ok, why = o.certify(keep.cycle, keep.ratio)
    assert ok, why
    print(f"C3: ratio {keep.ratio!r}, cycle {keep.cycle}, rounds {keep.metrics.iterations}")


def test_c5_sensitivity_full_size(S, oracle_mod):
    from min_ratio_cycle_b200 import synthetic
    n, U, V, C, T = synthetic.random_float(100_000, 2_000_000, seed=5)
    s = S.MinRatioCycleSolver(n, S.SolverConfig(log_level=S.LogLevel.NONE, enable_preprocessing=False))
    s.add_edges_arrays(U, V, C, T)
    base = s.solve()
    rng = np.random.default_rng(5)
    idx = rng.integers(0, len(U), 16)
    on = int(np.flatnonzero((U == base.cycle[0]) & (V == base.cycle[1]))[0])
    idx[0] = on                                        # one scenario hits the optimal cycle itself
    sign = rng.choice([-1.0, 1.0], 16)
    sign[0] = 1.0
    scen = [{int(i): (float(sg * 0.1 * C[i]) if i != on else 1000.0, 0.0)} for i, sg in zip(idx, sign)]
    res = s.sensitivity_analysis(scen)
    changed = 0
    for sc, r in zip(scen, res):
        C2 = C.copy()
        for i, (dc, dt) in sc.items():
            C2[i] += dc
        o = oracle_mod.Oracle(n, U, V, C2, T, all_int=False, preprocess=False, threads=8)
        o.set_quirks(False)
        ok, why = o.certify(r.cycle, r.ratio)
        assert ok, (sc, why)
        changed += r.cycle != base.cycle
    assert changed >= 1                                # the scenario on the optimal cycle must move the optimum
    again = s.solve()
    assert again.ratio == base.ratio


def test_c4_all_4096_currency_graphs(S, oracle_mod):
    """BASELINE config 4 at its stated size: 4,096 currency graphs (n=64, m=4,032) in ONE batch call; every result is
    certified by the oracle detector, and the graphs with a reference solve in the fixtures match it."""
    from conftest import load_golden
    from min_ratio_cycle_b200 import batch, synthetic
    seeded = load_golden("seeded.json")
    seeds = list(range(1000, 1000 + 4096))
    graphs = [synthetic.currency_graph(64, s) for s in seeds]
    res, stats = batch.solve_min_ratio_cycle_batch(graphs, S.SolverConfig(log_level=S.LogLevel.NONE), return_stats=True)
    assert len(res) == 4096 and stats["relax_launches"] == 1
    matched = 0
    for s, (n, U, V, C, T), r in zip(seeds, graphs, res):
        assert r.success and r.cycle[0] == r.cycle[-1]
        o = oracle_mod.Oracle(n, U, V, C, T, all_int=False, preprocess=False)
        o.set_quirks(False)
        ok, why = o.certify(r.cycle, r.ratio)
        assert ok, (s, why)
        gold = seeded.get(f"currency_{s}")
        if gold is not None:
            assert abs(r.ratio - gold["solve"]["ratio"]) <= 1e-9 * max(1.0, abs(gold["solve"]["ratio"])), s
            matched += 1
    assert matched >= 1
    print(f"C4: 4096 graphs, kernel {stats['relax_ms']:.2f} ms, {stats['relax_edge_visits'] / stats['relax_ms'] / 1e6:.0f} G relax/s")


def test_c5_all_1024_scenarios(S, oracle_mod):
    """BASELINE config 5 at its stated size: 1,024 single-edge perturbation re-solves of n=100k, m=2M.  Certified:
    64 scenarios spread evenly over the list plus every scenario whose optimal cycle moved; all others must report
    the unperturbed cycle with its sums re-costed under the scenario's weights."""
    from min_ratio_cycle_b200 import synthetic
    n, U, V, C, T = synthetic.random_float(100_000, 2_000_000, seed=5)
    m = len(U)
    s = S.MinRatioCycleSolver(n, S.SolverConfig(log_level=S.LogLevel.NONE, enable_preprocessing=False))
    s.add_edges_arrays(U, V, C, T)
    base = s.solve()
    rng = np.random.default_rng(5)
    idx = rng.integers(0, m, 1024)
    sign = rng.choice([-1.0, 1.0], 1024)
    hops = list(zip(base.cycle[:-1], base.cycle[1:]))
    for j, (a, b) in enumerate(hops[:8]):               # a few scenarios hit the optimal cycle itself, both ways
        idx[j] = int(np.flatnonzero((U == a) & (V == b))[0])
        sign[j] = 1.0 if j % 2 == 0 else -1.0
    scen = [{int(i): (float(sg * 0.1 * C[i]), 0.0)} for i, sg in zip(idx, sign)]
    scen[0] = {int(idx[0]): (1000.0, 0.0)}              # ... and one makes that cycle useless
    res = s.sensitivity_analysis(scen)
    assert len(res) == 1024 and all(r.success for r in res)
    moved = [k for k, r in enumerate(res) if r.cycle != base.cycle]
    assert 0 in moved
    check = sorted(set(range(0, 1024, 16)) | set(moved))
    for k in check:
        C2 = C.copy()
        for i, (dc, dt) in scen[k].items():
            C2[i] += dc
        o = oracle_mod.Oracle(n, U, V, C2, T, all_int=False, preprocess=False, threads=8)
        o.set_quirks(False)
        ok, why = o.certify(res[k].cycle, res[k].ratio)
        assert ok, (k, scen[k], why)
    for k, r in enumerate(res):
        if k in moved:
            continue
        (i, (dc, dt)), = scen[k].items()
        exp_c = base.sum_cost + (dc if (int(U[i]), int(V[i])) in set(hops) else 0.0)
        assert abs(r.sum_cost - exp_c) <= 1e-9 * max(1.0, abs(exp_c)), k
    assert s.solve().ratio == base.ratio                # device arrays restored
    print(f"C5: 1024 scenarios, {len(moved)} moved the optimum, {len(check)} certified")
