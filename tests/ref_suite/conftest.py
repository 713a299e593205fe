"""
The reference's own test-suite, re-stated against the drop-in ("missing" item 4).

What is checked is what /root/reference/tests checks for the solve path -- tests/test_solver.py:26-285,
:336-406, :484-575, tests/test_integration.py:19-545, tests/test_regression.py:20-50,
tests/test_analytics.py:23-29,50-78, tests/test_performance.py:24-50, tests/test_exceptions.py -- written
table-driven here (the graphs live in GRAPHS below, the assertions in the test modules), with
`min_ratio_cycle_b200` installed under the name `min_ratio_cycle`, so the test modules import exactly what the
reference's tests import (`from min_ratio_cycle.solver import MinRatioCycleSolver, SolverConfig`, ...).

Not re-stated, because the code they exercise is outside SURVEY 8 (plotting, statistics helpers, utils/):
test_analytics.py::test_visualization_tools, ::test_statistical_functions, tests/test_validation.py,
test_performance.py::test_solver_benchmark (needs the pytest-benchmark plugin, absent here and in the reference run).

Self-check of the re-statement: `MRC_REF_SUITE_TARGET=reference python -m pytest tests/ref_suite -m gpu -q`
(only in the build container, where /root/reference exists) runs the same files against the UNMODIFIED reference.
"""

from __future__ import annotations

import os
import sys
import types

import numpy as np
import pytest

TARGET = os.environ.get("MRC_REF_SUITE_TARGET", "b200")


def _install_target():
    if TARGET == "reference":
        ref = os.environ.get("MRC_REFERENCE", "/root/reference")
        if "matplotlib" not in sys.modules:
            try:
                import matplotlib.pyplot  # noqa: F401
            except Exception:
                mpl = types.ModuleType("matplotlib")
                mpl.use = lambda *a, **k: None
                plt = types.ModuleType("matplotlib.pyplot")
                plt.Figure = type("Figure", (), {})
                plt.Axes = type("Axes", (), {})
                mpl.pyplot = plt
                sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt
        sys.path.insert(0, ref)
        import min_ratio_cycle  # noqa: F401
        return
    import min_ratio_cycle_b200 as pkg
    import min_ratio_cycle_b200.benchmarks as bm
    import min_ratio_cycle_b200.exceptions as ex
    import min_ratio_cycle_b200.solver as sv
    sys.modules.setdefault("min_ratio_cycle", pkg)
    sys.modules.setdefault("min_ratio_cycle.solver", sv)
    sys.modules.setdefault("min_ratio_cycle.exceptions", ex)
    sys.modules.setdefault("min_ratio_cycle.benchmarks", bm)


_install_target()

from min_ratio_cycle.solver import MinRatioCycleSolver  # noqa: E402


def _ring_plus_random(n, extra, seed):
    """conftest.py:63-89: a ring through all vertices plus random chords, legacy NumPy generator."""
    rs = np.random.RandomState(seed)
    edges = []
    for i in range(n):
        edges.append((i, (i + 1) % n, int(rs.randint(-5, 6)), int(rs.randint(1, 4))))
    for _ in range(extra):
        u, v = int(rs.randint(0, n)), int(rs.randint(0, n))
        edges.append((u, v, int(rs.randint(-10, 11)), int(rs.randint(1, 6))))
    return edges


def _complete(n, seed):
    """conftest.py:114-130"""
    rs = np.random.RandomState(seed)
    return [(u, v, int(rs.randint(-3, 4)), int(rs.randint(1, 3))) for u in range(n) for v in range(n) if u != v]


def _pathological(n=20):
    """conftest.py:133-150: expensive path, cheap negative triangle at its end"""
    e = [(i, i + 1, 100, 1) for i in range(n - 1)]
    return e + [(n - 3, n - 2, 1, 10), (n - 2, n - 1, 1, 10), (n - 1, n - 3, -10, 1)]


# name -> (n, edges, best known ratio or None)
GRAPHS = {
    "simple_triangle": (3, [(0, 1, 2, 1), (1, 2, 3, 2), (2, 0, 1, 1)], 1.5),
    "negative_cycle": (3, [(0, 1, -1, 1), (1, 2, -2, 1), (2, 0, 1, 1)], -2 / 3),
    "integer_weights_only": (4, [(0, 1, 5, 2), (1, 2, 3, 1), (2, 3, 2, 2), (3, 0, 1, 1)], None),
    "float_weights": (4, [(0, 1, 5.5, 2.0), (1, 2, 3.2, 1.1), (2, 3, 2.7, 2.3), (3, 0, 1.1, 1.0)], None),
    "large_graph": (50, _ring_plus_random(50, 100, 42), None),
    "disconnected_graph": (8, [(0, 1, 3, 2), (1, 2, 2, 1), (2, 0, 1, 1), (3, 4, 1, 2), (4, 5, 1, 2), (5, 3, 0, 1)], 0.4),
    "pathological_graph": (20, _pathological(), None),
    "parallel_edges_graph": (3, [(0, 1, 10, 2), (0, 1, 6, 3), (0, 1, 8, 1), (1, 2, 2, 1), (2, 0, 1, 1)], None),
}


def build(name):
    n, edges, _ = GRAPHS[name]
    s = MinRatioCycleSolver(n)
    for u, v, c, t in edges:
        s.add_edge(u, v, c, t)
    return s


def _fixture_for(name, with_ratio):
    @pytest.fixture(name=name)
    def fx():
        s = build(name)
        return (s, GRAPHS[name][2]) if with_ratio else s
    return fx


simple_triangle = _fixture_for("simple_triangle", True)
negative_cycle = _fixture_for("negative_cycle", True)
disconnected_graph = _fixture_for("disconnected_graph", True)
integer_weights_only = _fixture_for("integer_weights_only", False)
float_weights = _fixture_for("float_weights", False)
large_graph = _fixture_for("large_graph", False)
pathological_graph = _fixture_for("pathological_graph", False)
parallel_edges_graph = _fixture_for("parallel_edges_graph", False)


@pytest.fixture(params=[3, 5, 7, 10])
def complete_graph(request):
    n = request.param
    s = MinRatioCycleSolver(n)
    for u, v, c, t in _complete(n, 42):
        s.add_edge(u, v, c, t)
    return s


class GraphAssertions:
    """conftest.py:172-211"""

    @staticmethod
    def assert_valid_cycle(cycle, n_vertices):
        assert isinstance(cycle, list) and len(cycle) >= 3 and cycle[0] == cycle[-1]
        assert all(0 <= v < n_vertices for v in cycle)
        assert all(a != b for a, b in zip(cycle[:-2], cycle[1:-1])), "consecutive duplicate"

    @staticmethod
    def assert_positive_time(sum_time):
        assert sum_time > 0

    @staticmethod
    def assert_ratio_consistency(sum_cost, sum_time, ratio):
        assert abs(ratio - sum_cost / sum_time) < 1e-10


@pytest.fixture
def graph_assertions():
    return GraphAssertions()
