"""
CPU-only checks: the C-ABI library loads without a GPU, exports every symbol include/mrc_b200.h declares,
fails loudly (no fallback) when there is no device, and the pure-host search helpers behave.
"""

from __future__ import annotations

import os
import re

import numpy as np
import pytest

from conftest import ROOT, has_cuda


@pytest.fixture(scope="module")
def native():
    from min_ratio_cycle_b200 import build
    build.build()
    from min_ratio_cycle_b200 import _native as N
    return N


def header_symbols():
    text = open(os.path.join(ROOT, "include", "mrc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mrc_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_exported(native):
    L = native.lib()
    syms = header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/mrc_b200.h but not exported"
    assert sorted(native.EXPORTS) == syms
    assert L.mrc_abi_version() == 4


def test_status_codes_match_header(native):
    text = open(os.path.join(ROOT, "include", "mrc_b200.h")).read()
    defs = dict(re.findall(r"#define (MRC_[A-Z_]+) (\d+)u?\b", text))
    assert int(defs["MRC_ERR_NO_CYCLE"]) == native.ERR_NO_CYCLE
    assert int(defs["MRC_ERR_CONVERGENCE"]) == native.ERR_CONVERGENCE
    assert int(defs["MRC_ERR_TIMEOUT"]) == native.ERR_TIMEOUT
    assert int(defs["MRC_ERR_SB_STEPS"]) == native.ERR_SB_STEPS
    assert int(defs["MRC_V_TIE"]) == native.V_TIE
    assert int(defs["MRC_V_SKIPPED_NONNEG"]) == native.V_SKIPPED_NONNEG
    assert int(defs["MRC_V_ABANDONED"]) == native.V_ABANDONED and int(defs["MRC_F_STOP_ON_NEG"]) == native.F_STOP_ON_NEG
    assert int(defs["MRC_KMAX"]) == native.KMAX


@pytest.mark.skipif(has_cuda(), reason="checks the no-device failure mode")
def test_no_device_fails_loudly(native):
    """The product path has no CPU fallback: without a device graph creation raises."""
    with pytest.raises(native.NativeError) as ei:
        native.DeviceGraph(3, [0, 1, 2], [1, 2, 0], [1.0, 1.0, 1.0], [1.0, 1.0, 1.0])
    assert ei.value.code == native.ERR_NO_DEVICE
    from min_ratio_cycle_b200.solver import LogLevel, MinRatioCycleSolver, SolverConfig
    s = MinRatioCycleSolver(3, SolverConfig(log_level=LogLevel.NONE))
    s.add_edges([(0, 1, 1.5, 1.0), (1, 2, 1.0, 1.0), (2, 0, 1.0, 1.0)])
    with pytest.raises(native.NativeError):
        s.solve()


def test_plan_candidates(native):
    lam = native.plan_candidates(-3.0, 5.0, False, 1e-12, 1)
    assert lam[0] == (-3.0 + 5.0) / 2.0                      # the reference's midpoint (solver.py:1053)
    lam = native.plan_candidates(0.0, 10.0, False, 1e-12, 4)
    assert np.allclose(lam, [2, 4, 6, 8]) and np.all(np.diff(lam) > 0)
    lam = native.plan_candidates(0.0, 10.0, True, 1e-12, 4)  # hi is a known cycle ratio: certifier last
    assert np.all(np.diff(lam) > 0) and 0 < lam[0] and lam[-1] < 10.0 and 10.0 - lam[-1] <= 0.75e-12 * 1.01 + 2e-15
    lam = native.plan_candidates(1.0, 1.0 + 4e-13, True, 1e-12, 8)
    assert np.all(lam > 1.0) and np.all(lam < 1.0 + 4e-13) and np.all(np.diff(lam) >= 0)


def test_reduce_round(native):
    N = native
    nan = float("nan")
    lo, hi, hc, best = N.reduce_round([2, 4, 6, 8], [N.V_NONNEG, N.V_NONNEG, N.V_NEG, N.V_SKIPPED_NEG],
                                      [nan, nan, 5.5, nan], 0.0, 10.0, False)
    assert (lo, hi, hc, best) == (4.0, 5.5, True, 2)
    # a NEG without extracted cycle only moves hi to lambda (solver.py:1057-1060)
    lo, hi, hc, best = N.reduce_round([2, 4], [N.V_NONNEG, N.V_NEG_NOCYCLE], [nan, nan], 0.0, 10.0, False)
    assert (lo, hi, hc, best) == (2.0, 4.0, False, -1)
    # cycle worse than the known one is ignored
    lo, hi, hc, best = N.reduce_round([2, 4], [N.V_NONNEG, N.V_NEG], [nan, 3.9], 0.0, 3.5, True)
    assert (lo, hi, hc, best) == (2.0, 3.5, True, -1)
    # TIE: lambda ~ ratio of the extracted cycle: both ends move
    lo, hi, hc, best = N.reduce_round([3.0], [N.V_TIE], [3.0 + 1e-15], 0.0, 10.0, False)
    assert lo == 3.0 and hi == 3.0 + 1e-15 and hc and best == 0


def test_multisection_converges_on_model(native):
    """Drive plan/reduce with an analytic verdict model: cycles with ratios R, lambda* = min R."""
    N = native
    R = np.array([3.7, 1.25, 9.0, 1.2500001])
    for K in (1, 2, 4, 8, 32):
        lo, hi, hc, it = -5.0, 12.0, False, 0
        while hi - lo > 1e-12 and it < 60:
            it += 1
            lam = N.plan_candidates(lo, hi, hc, 1e-12, K)
            vd, cr = [], []
            for x in lam:
                neg = R[R < x]
                if neg.size:
                    vd.append(N.V_NEG); cr.append(float(neg.max()))      # adversarial: the worst negative cycle
                else:
                    vd.append(N.V_NONNEG); cr.append(float("nan"))
            lo, hi, hc, _ = N.reduce_round(lam, vd, cr, lo, hi, hc)
        assert hi == 1.25 and hi - lo <= 1e-12, (K, lo, hi, it)
        assert it <= (60 if K == 1 else 30)


def test_facade_api_surface():
    import min_ratio_cycle_b200 as pkg
    assert set(pkg.__all__) == {"MinRatioCycleSolver", "Edge"}          # tests/test_regression.py:50
    from min_ratio_cycle_b200.solver import (Edge, LogLevel, MinRatioCycleSolver, SolverConfig, SolverMetrics,
                                             SolverMode, SolverResult, solve_min_ratio_cycle)
    cfg = SolverConfig()
    assert (cfg.numeric_max_iter, cfg.numeric_tolerance, cfg.numeric_cycle_slack) == (60, 1e-12, 1e-15)
    assert cfg.validate_cycles and cfg.repair_cycles and cfg.use_kahan_summation and cfg.max_solve_time is None
    assert [m.value for m in SolverMode] == ["auto", "exact", "numeric", "approximate"]
    s = MinRatioCycleSolver(4, SolverConfig(log_level=LogLevel.NONE))
    for name in ("add_edge", "add_edges", "remove_edge", "clear_edges", "solve", "sensitivity_analysis",
                 "stability_region", "get_graph_properties", "detect_issues", "get_debug_info", "health_check",
                 "_build_arrays_if_needed", "_build_numpy_arrays_once", "_solve_numeric", "_solve_exact",
                 "_stern_brocot_search", "_has_negative_cycle_exact", "_compute_potentials_exact",
                 "_find_zero_cycle_exact", "_detect_negative_cycle_numeric", "_compute_cycle_weights_float",
                 "_compute_cycle_weights_exact", "_preprocess_graph", "_remove_dominated_edges", "_validate_result"):
        assert callable(getattr(s, name)), name
    for attr in ("n", "config", "_edges", "_all_int", "_arrays_built", "_U", "_V", "_C", "_T", "_starts",
                 "_counts", "_last_result"):
        assert hasattr(s, attr), attr
    with pytest.raises(ValueError):
        MinRatioCycleSolver(0)
    with pytest.raises(ValueError, match="valid vertex indices"):
        s.add_edge(0, 9, 1, 1)
    with pytest.raises(ValueError, match="strictly positive"):
        s.add_edge(0, 1, 1, 0)
    with pytest.raises(ValueError):
        Edge(0.5, 1, 1, 1)
    s.add_edge(0, 1, 2, 1)
    assert s._all_int
    s.add_edge(1, 0, 2.5, 1)
    assert not s._all_int and len(s._edges) == 2 and not s._arrays_built
    assert s.remove_edge(1, 0) and not s.remove_edge(3, 3)
    r = SolverResult([0, 1, 0], 1.0, 2.0, 0.5)
    assert tuple(r) == ([0, 1, 0], 1.0, 2.0, 0.5)
    assert SolverMetrics().mode_used == "" and callable(solve_min_ratio_cycle)
    with pytest.raises(ValueError):
        s.solve(mode="bogus")


def test_dominated_edge_removal_matches_oracle(oracle_mod):
    """Vectorised _remove_dominated_edges vs the oracle's restatement of solver.py:531-575."""
    from min_ratio_cycle_b200.solver import LogLevel, MinRatioCycleSolver, SolverConfig
    rng = np.random.default_rng(5)
    for trial in range(30):
        n = int(rng.integers(2, 7))
        m = int(rng.integers(1, 40))
        U = rng.integers(0, n, m); V = rng.integers(0, n, m)
        C = rng.integers(-3, 4, m).astype(float); T = rng.integers(1, 4, m).astype(float)
        s = MinRatioCycleSolver(n, SolverConfig(log_level=LogLevel.NONE))
        half = m // 2
        for i in range(half):
            s.add_edge(int(U[i]), int(V[i]), float(C[i]), float(T[i]))
        s.add_edges_arrays(U[half:], V[half:], C[half:], T[half:])
        s._remove_dominated_edges()
        got = s._insertion_arrays()
        o = oracle_mod.Oracle(n, U, V, C, T)
        exp = o.ins
        for a, b in zip(got, exp):
            assert np.array_equal(a, b), trial


def test_dimacs_loader_and_regression_helpers(tmp_path):
    """Host-only parts of benchmarks.py (reference benchmarks.py:19-66, :143-234)."""
    from min_ratio_cycle_b200 import benchmarks as B
    lines = ["c unit triangle", "p sp 3 3", "a 1 2 1 1", "a 2 3 1 1", "", "a 3 1 1 1"]
    s = B.load_dimacs_graph(lines)
    assert s.n == 3 and [(e.u, e.v, e.cost, e.time) for e in s._edges] == [(0, 1, 1.0, 1.0), (1, 2, 1.0, 1.0), (2, 0, 1.0, 1.0)]
    assert B.load_dimacs_graph(["p 5 0"]).n == 5
    p = tmp_path / "sub" / "base.json"
    B.save_baseline(p, {"ratio": 1.0, "times": [0.01, 0.012, 0.011]})
    base = B.load_baseline(p)
    assert B.regression_check([0.011, 0.0105, 0.012], 1.0, base) == {"performance": False, "quality": False}
    assert B.regression_check([0.5, 0.6, 0.55], 1.0 + 1e-6, base) == {"performance": True, "quality": True}
    assert B.regression_check([0.5], 1.0, base)["performance"] is True          # < 2 samples: 1.5x mean rule


def test_first_round_levels_are_a_low_quantile_ladder():
    """plan_first_round (csrc/mrc_b200.cu): the first round of a search without a known cycle looks at low quantiles of the
    edge-ratio sample -- increasing, inside (0, 1), denser towards the low tail; one candidate sits at 5 %."""
    from min_ratio_cycle_b200 import _native as N
    assert N.first_round_levels(1)[0] == pytest.approx(0.05)
    for K in (2, 3, 4, 8, 16, 32, 64):
        q = N.first_round_levels(K)
        assert len(q) == K and np.all(np.diff(q) > 0) and q[0] > 0 and q[-1] <= 0.5 + 1e-12
        assert np.allclose(q[1:] / q[:-1], q[1] / q[0])          # geometric
    assert N.first_round_levels(4)[0] == pytest.approx(0.01) and N.first_round_levels(4)[-1] == pytest.approx(0.25)
    with pytest.raises(N.NativeError):
        N.first_round_levels(0)
