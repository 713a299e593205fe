"""
Re-statement of /root/reference/tests/test_solver.py against the drop-in (see conftest.py of this package).
Line references are to that file.
"""

from __future__ import annotations

import itertools
import os

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from min_ratio_cycle.exceptions import NumericalInstabilityError, ResourceExhaustionError
from min_ratio_cycle.solver import MinRatioCycleSolver, SolverConfig

pytestmark = pytest.mark.gpu


def solver_of(n, edges):
    s = MinRatioCycleSolver(n)
    for u, v, c, t in edges:
        s.add_edge(u, v, c, t)
    return s


# ---- TestEdgeCases (:26-153)
@pytest.mark.parametrize("n", [3, 1])
def test_graph_without_edges_is_a_value_error(n):          # :31-45
    with pytest.raises(ValueError, match="Graph has no edges"):
        MinRatioCycleSolver(n).solve()


def test_zero_vertices_rejected():                          # :47-52
    with pytest.raises(ValueError, match="n_vertices must be a positive integer"):
        MinRatioCycleSolver(0)


def test_tree_has_no_cycle():                               # :54-67
    s = solver_of(4, [(0, 1, 5, 2), (1, 2, 3, 1), (1, 3, 4, 2)])
    with pytest.raises(NumericalInstabilityError):
        s.solve()


def test_single_self_loop():                                # :69-80
    cycle, sc, stt, r = solver_of(1, [(0, 0, 3, 2)]).solve()
    assert cycle == [0, 0]
    assert sc == 3 and stt == 2
    assert abs(r - 1.5) < 1e-10


def test_parallel_edges_pick_the_better_one():              # :82-96
    cycle, sc, stt, r = solver_of(2, [(0, 1, 10, 2), (0, 1, 6, 3), (1, 0, 1, 1)]).solve()
    assert len(cycle) == 3 and cycle[0] == cycle[-1]
    assert abs(r - 1.75) < 1e-10


def test_best_component_wins():                             # :98-114
    e = [(0, 1, 2, 1), (1, 2, 3, 2), (2, 0, 1, 1), (3, 4, 1, 2), (4, 3, 1, 2)]
    assert abs(solver_of(6, e).solve().ratio - 0.5) < 1e-10


@pytest.mark.parametrize("t", [0, -1])
def test_non_positive_time_rejected(t):                     # :116-127
    with pytest.raises(ValueError, match="time must be strictly positive"):
        MinRatioCycleSolver(2).add_edge(0, 1, 1, t)


@pytest.mark.parametrize("u,v", [(-1, 0), (0, 3)])
def test_vertex_ids_checked(u, v):                          # :129-138
    with pytest.raises(ValueError, match="valid vertex indices"):
        MinRatioCycleSolver(3).add_edge(u, v, 1, 1)


def test_weights_of_1e15_end_in_instability_error():        # :140-153
    big = 10**15
    s = solver_of(3, [(0, 1, big, 1), (1, 2, 1, big), (2, 0, 1, 1)])
    with pytest.raises(NumericalInstabilityError):
        s.solve()


# ---- TestCorrectnessValidation (:156-249): walk the returned cycle through the solver's own arrays
def walk_through_arrays(s, cycle):
    s._build_numpy_arrays_once()
    first = {}
    for i in range(len(s._U)):
        first.setdefault((int(s._U[i]), int(s._V[i])), (float(s._C[i]), float(s._T[i])))
    c = t = 0.0
    for hop in zip(cycle[:-1], cycle[1:]):
        assert hop in first, f"edge {hop} of the cycle is not in the graph"
        c += first[hop][0]
        t += first[hop][1]
    return c, t


@pytest.mark.parametrize("edges", [
    [(0, 1, 2, 1), (1, 2, 3, 2), (2, 0, 1, 1)],                 # :228-238
    [(0, 1, 2.5, 1.0), (1, 2, 3, 2.0), (2, 0, 1.5, 1.0)],       # :240-249
])
def test_cycle_recomputed_from_private_arrays(edges):
    s = solver_of(3, edges)
    cycle, sc, stt, r = s.solve()
    assert cycle[0] == cycle[-1] and len(cycle) >= 3
    c, t = walk_through_arrays(s, cycle)
    assert abs(c - sc) < 1e-10 and abs(t - stt) < 1e-10 and abs(c / t - r) < 1e-10


# ---- TestResultPackaging (:252-285)
def test_result_unpacks_like_a_tuple(simple_triangle):      # :257-264
    s, _ = simple_triangle
    res = s.solve()
    cycle, cost, time, ratio = res
    assert (cycle, cost, time, ratio) == (res.cycle, res.sum_cost, res.sum_time, res.ratio)


def test_arrays_are_built_once():                           # :266-285
    s = solver_of(2, [(0, 1, 1, 1), (1, 0, 1, 1)])
    builds, inner = [], s._build_arrays_if_needed

    def spy():
        if not s._arrays_built:
            builds.append(1)
        inner()

    s._build_arrays_if_needed = spy
    s._build_numpy_arrays_once()
    s._build_numpy_arrays_once()
    assert len(builds) == 1


# ---- TestPropertyBased (:288-406)
@st.composite
def small_graphs(draw, integer):                            # :288-328 (n <= 6, m <= 15; loops and parallel edges allowed)
    n = draw(st.integers(2, 6))
    m = draw(st.integers(n, 15))
    vid = st.integers(0, n - 1)
    if integer:
        cw, tw = st.integers(-100, 100), st.integers(1, 20)
    else:
        cw = st.floats(-100.0, 100.0, allow_nan=False, allow_infinity=False)
        tw = st.floats(0.1, 20.0, allow_nan=False, allow_infinity=False)
    return n, [(draw(vid), draw(vid), draw(cw), draw(tw)) for _ in range(m)]


def _check_properties(n, edges):                            # :347-364
    try:
        cycle, sc, stt, r = solver_of(n, edges).solve()
    except (RuntimeError, ValueError, NumericalInstabilityError):
        return                                              # no cycle etc.: tolerated by the reference's test too
    assert cycle[0] == cycle[-1]
    assert abs(r - sc / stt) < 1e-10
    assert stt > 0
    assert 2 <= len(cycle) <= len(set(edges)) + 1


@given(small_graphs(True))
@settings(max_examples=50, deadline=None, suppress_health_check=list(HealthCheck))
def test_property_integer_graphs(g):                        # :336-364
    _check_properties(*g)


@given(small_graphs(False))
@settings(max_examples=50, deadline=None, suppress_health_check=list(HealthCheck))
def test_property_float_graphs(g):                          # :366-382
    _check_properties(*g)


@pytest.mark.parametrize("n", range(2, 9))
def test_complete_graph_always_has_a_cycle(n):              # :384-406
    rs = np.random.RandomState(n)
    e = [(u, v, int(rs.randint(-10, 11)), int(rs.randint(1, 6))) for u in range(n) for v in range(n) if u != v]
    cycle, _, stt, _ = solver_of(n, e).solve()
    assert len(cycle) >= 3 and cycle[0] == cycle[-1] and stt > 0


# ---- TestBenchmarks (:408-566): brute force over simple cycles
def brute_force_ratio(n, edges):                            # :408-459 restated with itertools
    best = None
    w = {}
    for u, v, c, t in edges:
        w.setdefault((u, v), []).append((c, t))
    for L in range(1, n + 1):
        for verts in itertools.permutations(range(n), L):
            if verts[0] != min(verts):
                continue
            hops = list(zip(verts, verts[1:] + verts[:1]))
            if L == 1:
                continue                                    # the reference's DFS needs len(path) >= 2 to close
            if not all(h in w for h in hops):
                continue
            for choice in itertools.product(*(w[h] for h in hops)):
                r = sum(c for c, _ in choice) / sum(t for _, t in choice)
                best = r if best is None or r < best else best
    return best


@pytest.mark.parametrize("n", [3, 4, 5])
def test_against_brute_force(n):                            # :484-532 (the reference swallows failures; here they count)
    rs = np.random.RandomState(42)
    edges = [(int(rs.randint(0, n)), int(rs.randint(0, n)), int(rs.randint(-5, 6)), int(rs.randint(1, 4)))
             for _ in range(2 * n)]
    try:
        r = solver_of(n, edges).solve().ratio
    except (RuntimeError, NumericalInstabilityError):
        pytest.skip("graph has no cycle")
    loops = [c / t for u, v, c, t in edges if u == v]
    bf = brute_force_ratio(n, edges)
    cands = ([bf] if bf is not None else []) + loops        # a self-loop may beat every cycle the reference's DFS sees
    assert cands and r >= min(cands) - 1e-6                 # never better than the true optimum
    if os.environ.get("MRC_REF_SUITE_TARGET", "b200") != "reference":
        # the reference itself is sub-optimal on the n=5 draw (-2/3 against -8/7: SURVEY 0.6) and hides it in a bare
        # try/except (:520-532); the drop-in has to find the optimum
        assert r <= min(cands) + 1e-6


@pytest.mark.parametrize("n,density", [(10, 0.1), (20, 0.1), (50, 0.1), (100, 0.1), (20, 0.05), (20, 0.2), (20, 0.5)])
def test_random_density_sweep_runs(n, density):             # :534-566 (prints timings; must not crash)
    rs = np.random.RandomState(n + int(100 * density))
    s = MinRatioCycleSolver(n)
    for _ in range(int(n * n * density)):
        s.add_edge(int(rs.randint(0, n)), int(rs.randint(0, n)), int(rs.randint(-10, 11)), int(rs.randint(1, 6)))
    try:
        cycle, _, stt, r = s.solve()
    except (RuntimeError, ValueError, NumericalInstabilityError):
        return
    assert cycle[0] == cycle[-1] and stt > 0 and np.isfinite(r)


def test_memory_limit_enforced():                           # :569-575
    s = MinRatioCycleSolver(2, SolverConfig(max_memory_mb=0))
    s.add_edge(0, 1, 1, 1)
    s.add_edge(1, 0, 1, 1)
    with pytest.raises(ResourceExhaustionError):
        s.solve()
