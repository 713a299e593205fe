"""
Re-statement of /root/reference/tests/test_regression.py, test_analytics.py (solver-facing parts),
test_performance.py (without the pytest-benchmark fixture) and test_exceptions.py.
"""

from __future__ import annotations

import json
import math
import statistics
import time

import psutil
import pytest

import min_ratio_cycle as mrc
from min_ratio_cycle.benchmarks import benchmark_solver, load_dimacs_graph
from min_ratio_cycle.exceptions import GraphStructureError, NumericalInstabilityError, ResourceExhaustionError
from min_ratio_cycle.solver import MinRatioCycleSolver

pytestmark = pytest.mark.gpu

# tests/baselines/triangle.json of the reference (api, ratio, times of the unit triangle)
BASELINE = json.loads('{"api": ["MinRatioCycleSolver", "Edge"], "ratio": 1.0, "times": [0.0173, 0.0145, 0.0159]}')
UNIT_TRIANGLE = ["c simple triangle", "p 3 3", "a 1 2 1 1", "a 2 3 1 1", "a 3 1 1 1"]


def welch_t(x, y):   # analytics.compare_solutions (analytics.py:99-128) -- statistics helper, out of scope of the package
    d = math.sqrt(statistics.variance(x) / len(x) + statistics.variance(y) / len(y))
    return math.inf if d == 0 else (statistics.mean(x) - statistics.mean(y)) / d


def test_regression_baseline():                             # test_regression.py:20-50
    times, ratios = [], []
    for _ in range(3):
        stats = benchmark_solver(load_dimacs_graph(UNIT_TRIANGLE), compare=False)
        times.append(stats["time"])
        ratios.append(stats["ratio"])
    base_mean = statistics.mean(BASELINE["times"])
    assert welch_t(times, BASELINE["times"]) < 5.0
    assert statistics.mean(times) <= 1.5 * base_mean
    assert all(abs(r - BASELINE["ratio"]) < 1e-9 for r in ratios)
    assert max(ratios) - min(ratios) < 1e-12
    assert set(mrc.__all__) == set(BASELINE["api"])


def unit_triangle():
    s = MinRatioCycleSolver(3)
    for u, v in ((0, 1), (1, 2), (2, 0)):
        s.add_edge(u, v, 1, 1)
    return s


def test_sensitivity_and_stability_sizes():                 # test_analytics.py:23-29
    s = unit_triangle()
    s.solve()
    assert len(s.sensitivity_analysis([{0: (0.5, 0.0)}])) == 1
    assert len(s.stability_region(0.05)) == 3


def test_dimacs_five_vertices_matches_brute_force():        # test_analytics.py:50-64
    g = ["p 5 7", "a 1 2 3 1", "a 2 3 4 1", "a 3 1 2 1", "a 2 4 1 1", "a 4 5 3 2", "a 5 2 0 1", "a 3 5 2 2"]
    stats = benchmark_solver(load_dimacs_graph(g))
    assert stats["diff"] == pytest.approx(0.0, abs=1e-9) and stats["time"] >= 0.0


def test_benchmark_without_compare():                       # test_analytics.py:67-78
    g = ["c sample graph", "p sp 3 3", "a 1 2 1 1", "a 2 3 1 1", "a 3 1 1 1"]
    stats = benchmark_solver(load_dimacs_graph(g), compare=False)
    assert stats["diff"] is None and stats["ratio"] == pytest.approx(1.0, abs=1e-9)


def two_way_ring(n):                                        # test_performance.py:8-13
    s = MinRatioCycleSolver(n)
    for i in range(n):
        j = (i + 1) % n
        s.add_edge(i, j, 1, 1)
        s.add_edge(j, i, 1, 1)
    return s


def test_fifty_solves_do_not_leak():                        # test_performance.py:24-35
    two_way_ring(5).solve()                                 # first use of the library (context, stream pool) is not a leak
    proc = psutil.Process()
    before = proc.memory_info().rss
    for _ in range(50):
        two_way_ring(5).solve()
    assert proc.memory_info().rss - before < 5 * 1024 * 1024


def test_doubling_the_ring_is_less_than_5x():               # test_performance.py:38-50
    def timed(n):
        s = two_way_ring(n)
        t0 = time.perf_counter()
        s.solve()
        return time.perf_counter() - t0
    timed(20)
    assert timed(40) < 5 * timed(20)


def test_large_graph_mean_below_600_ms(large_graph):        # test_performance.py:16-21 without pytest-benchmark
    large_graph.solve()
    t0 = time.perf_counter()
    for _ in range(5):
        large_graph.solve()
    assert (time.perf_counter() - t0) / 5 < 0.6


def test_exception_details():                               # test_exceptions.py:8-21
    err = NumericalInstabilityError("unstable", suggested_fix="increase precision", recovery_hint="use exact mode")
    assert "increase precision" in str(err) and err.details["recovery_hint"] == "use exact mode"
    assert GraphStructureError("bad graph", suggested_fix="connect components").details["suggested_fix"] == "connect components"
    assert ResourceExhaustionError("oom", resource="memory", limit=1, usage=2).details["usage"] == 2
