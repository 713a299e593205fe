"""Re-statement of /root/reference/tests/test_integration.py (line references are to that file)."""

from __future__ import annotations

import math
import time

import psutil
import pytest

from min_ratio_cycle.benchmarks import benchmark_solver, load_dimacs_graph
from min_ratio_cycle.exceptions import SolverError
from min_ratio_cycle.solver import MinRatioCycleSolver

pytestmark = pytest.mark.gpu


def solver_of(n, edges):
    s = MinRatioCycleSolver(n)
    s.add_edges(edges)
    return s


def solved(s, ga):
    cycle, sc, stt, r = s.solve()
    ga.assert_valid_cycle(cycle, s.n)
    ga.assert_positive_time(stt)
    ga.assert_ratio_consistency(sc, stt, r)
    return cycle, sc, stt, r


# ---- TestIntegration (:14-169)
def test_known_optimum(simple_triangle, graph_assertions):                  # :19-35
    s, best = simple_triangle
    assert solved(s, graph_assertions)[3] <= best + 1e-6


def test_negative_ratio(negative_cycle, graph_assertions):                  # :37-51
    s, best = negative_cycle
    r = solved(s, graph_assertions)[3]
    assert r < 0 and abs(r - best) < 1e-6


def test_integer_graph_returns_integral_floats(integer_weights_only):       # :53-71
    s = integer_weights_only
    assert s._all_int is True
    _, sc, stt, r = s.solve()
    assert isinstance(sc, float) and isinstance(stt, float) and isinstance(r, float)
    assert sc == int(sc) and stt == int(stt)


def test_float_graph(float_weights):                                        # :73-87
    s = float_weights
    assert s._all_int is False
    cycle, _, stt, r = s.solve()
    assert len(cycle) >= 3 and stt > 0 and math.isfinite(r)


def test_disconnected(disconnected_graph, graph_assertions):                # :89-102
    s, best = disconnected_graph
    assert abs(solved(s, graph_assertions)[3] - best) < 1e-6


def test_complete_graphs(complete_graph, graph_assertions):                 # :104-118
    assert math.isfinite(solved(complete_graph, graph_assertions)[3])


def test_parallel_edges(parallel_edges_graph, graph_assertions):            # :120-134
    assert solved(parallel_edges_graph, graph_assertions)[3] < 10


def test_large_graph_under_five_seconds(large_graph, graph_assertions):     # :136-155
    t0 = time.time()
    solved(large_graph, graph_assertions)
    assert time.time() - t0 < 5.0


def test_pathological(pathological_graph, graph_assertions):                # :157-169
    assert solved(pathological_graph, graph_assertions)[3] < 0


# ---- TestRealWorldScenarios (:172-300): shapes only -- len(cycle) >= 3, time > 0, sign of the ratio where stated
SCENARIOS = {
    "arbitrage": (4, [(0, 1, -math.log(0.85), 1), (1, 2, -math.log(1.15), 1), (2, 3, -math.log(150.0), 1),
                      (3, 0, -math.log(0.0075), 1), (1, 0, -math.log(1.18), 1), (2, 1, -math.log(0.87), 1)], None),
    "scheduling": (5, [(0, 1, 10, 2), (1, 2, 15, 3), (2, 3, 8, 1), (3, 4, 12, 4), (4, 1, 5, 2), (0, 3, 20, 5),
                       (2, 0, 6, 1)], "positive"),
    "routing": (6, [(0, 1, 5, 10), (1, 2, 3, 5), (2, 3, 7, 15), (3, 4, 2, 8), (4, 5, 4, 12), (5, 0, 6, 20),
                    (0, 2, 12, 18), (1, 4, 8, 25), (3, 0, 15, 30)], None),
    "manufacturing": (4, [(0, 1, 100, 30), (1, 2, 150, 45), (2, 3, 80, 20), (3, 0, 50, 15), (0, 2, 200, 60),
                          (1, 3, 120, 35)], "positive"),
}


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_real_world_shapes(name):                                           # :177-300
    n, edges, sign = SCENARIOS[name]
    cycle, _, stt, r = solver_of(n, edges).solve()
    assert len(cycle) >= 3 and stt > 0
    if sign == "positive":
        assert r > 0


# ---- TestEdgeConfiguration (:303-414)
def test_only_self_loops():                                                 # :306-322
    cycle, _, _, r = solver_of(3, [(0, 0, 5, 2), (1, 1, 3, 2), (2, 2, 7, 4)]).solve()
    assert len(cycle) == 2 and cycle[0] == cycle[1] and abs(r - 1.5) < 1e-10


def test_star_cycle_goes_through_the_centre():                              # :324-347
    n, e = 6, []
    for i in range(1, n):
        e += [(0, i, i * 2, i), (i, 0, i + 1, 1)]
    e += [(i, i + 1, 1, 2) for i in range(1, n - 1)]
    cycle, _, stt, _ = solver_of(n, e).solve()
    assert len(cycle) >= 3 and stt > 0 and 0 in cycle[:-1]


def test_bipartite_cycles_are_even():                                       # :349-370
    e = []
    for u in (0, 1, 2):
        for v in (3, 4, 5):
            c, t = abs(u - v) + 1, (u + v) % 3 + 1
            e += [(u, v, c, t), (v, u, c + 1, t)]
    cycle, _, stt, _ = solver_of(6, e).solve()
    assert len(cycle) >= 3 and stt > 0 and (len(cycle) - 1) % 2 == 0


def test_layered_cycle_spans_layers():                                      # :372-414
    layers, width = 4, 3
    nid = lambda layer, pos: layer * width + pos  # noqa: E731
    e = [(nid(a, p), nid(a + 1, q), abs(p - q) + 1, a + 1) for a in range(layers - 1) for p in range(width) for q in range(width)]
    e += [(nid(layers - 1, p), nid(0, q), p + q + 1, layers) for p in range(width) for q in range(width)]
    cycle, _, stt, _ = solver_of(layers * width, e).solve()
    assert len(cycle) >= 3 and stt > 0
    assert len({v // width for v in cycle[:-1]}) >= 2


# ---- TestNumericalStability (:417-485)
def test_differences_of_1e_12():                                            # :422-441
    eps = 1e-12
    cycle, _, stt, r = solver_of(3, [(0, 1, 1.0, 1.0), (1, 2, 1.0 + eps, 1.0), (2, 0, 1.0 - eps, 1.0)]).solve()
    assert len(cycle) >= 3 and stt > 0 and abs(r - 1.0) < 1e-10


def test_mixed_scales():                                                    # :443-463
    cycle, _, stt, r = solver_of(3, [(0, 1, 1e6, 1.0), (1, 2, 1e-6, 1e6), (2, 0, 1.0, 1.0)]).solve()
    assert len(cycle) >= 3 and stt > 0 and math.isfinite(r)


def test_near_zero_ratio():                                                 # :465-485
    cycle, _, stt, r = solver_of(4, [(0, 1, 1000, 1), (1, 2, -999.99, 1), (2, 3, -0.005, 1), (3, 0, -0.005, 1)]).solve()
    assert len(cycle) >= 3 and stt > 0 and abs(r) < 0.1


# ---- TestEndToEndWorkflows (:488-545)
UNIT_TRIANGLE = ["c simple triangle", "p 3 3", "a 1 2 1 1", "a 2 3 1 1", "a 3 1 1 1"]


def test_dimacs_end_to_end():                                               # :493-515
    s = load_dimacs_graph(UNIT_TRIANGLE)
    proc = psutil.Process()
    before = proc.memory_info().rss
    stats = benchmark_solver(s)
    grown = proc.memory_info().rss - before
    assert stats["diff"] is None or stats["diff"] < 1e-9
    assert 0 < stats["time"] < 1.0
    assert grown < 50 * 1024 * 1024


def test_two_solvers_same_ratio_exactly():                                  # :517-526
    e = [(0, 1, 1.0, 1.0), (1, 2, 1.0, 1.0), (2, 0, 1.0, 1.0)]
    assert solver_of(3, e).solve().ratio == solver_of(3, e).solve().ratio


def test_acyclic_dimacs_raises_solver_error_with_hints():                   # :528-545
    s = load_dimacs_graph(["p 4 2", "a 1 2 1 1", "a 3 4 1 1"])
    with pytest.raises(SolverError) as info:
        benchmark_solver(s)
    assert "suggested_fix" in info.value.details and "recovery_hint" in info.value.details
