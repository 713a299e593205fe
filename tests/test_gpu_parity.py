"""
GPU parity tests (run with -m gpu on a B200).  Every call goes through the C ABI
(libmrc_b200.so) -- directly via _native.DeviceGraph or through the MinRatioCycleSolver facade --
and is checked against (a) the golden fixtures written by the Python reference and (b) the CPU
oracle on the same seeded inputs.

Tolerances: exact mode is compared with == (cycle list, integer sums, reduced fraction, Stern-Brocot
probe sequence, iteration count).  Numeric mode: |ratio_gpu - ratio_ref| <= 1e-9 * max(1, |ratio_ref|)
(north star; the reference's own isclose(rel=abs=1e-9), utils/validation.py:317), with the three-way
verdict of SURVEY 8c where the reference itself is sub-optimal.
"""

from __future__ import annotations

from fractions import Fraction

import numpy as np
import pytest

from conftest import edges_all_int, edges_to_arrays, load_golden, rotations_equal
from parity import MATCH, MISMATCH, REF_SUBOPTIMAL, cycle_is_valid, three_way

pytestmark = pytest.mark.gpu

KAT = load_golden("kat.json")
SMALL = load_golden("random_small.json")
SEEDED = load_golden("seeded.json")


@pytest.fixture(scope="module")
def mrc():
    import min_ratio_cycle_b200.solver as S
    from min_ratio_cycle_b200 import _native as N
    assert N.device_count() > 0, "no CUDA device: these tests need a B200"
    return S


def facade(S, n, edges, **cfg):
    s = S.MinRatioCycleSolver(int(n), S.SolverConfig(log_level=S.LogLevel.NONE, **cfg))
    for e in edges:
        s.add_edge(int(e[0]), int(e[1]), e[2], e[3])
    return s


def checker(oracle_mod, rec):
    U, V, C, T = edges_to_arrays(rec["edges"])
    o = oracle_mod.Oracle(rec["n"], U, V, C, T, all_int=edges_all_int(rec["edges"]))
    o.set_quirks(False)
    return o


# --------------------------------------------------------------------------- reference KATs
KAT_EXPECT = {  # name -> (ratio, tolerance) from the reference's own tests
    "triangle": (1.5, 1e-6), "negative_cycle": (-2.0 / 3.0, 1e-6), "self_loop": (1.5, 1e-10),
    "parallel_edges": (1.75, 1e-10), "two_components": (0.5, 1e-10), "disconnected8": (0.4, 1e-6),
    "three_self_loops": (1.5, 1e-10), "unit_triangle": (1.0, 1e-9), "unit_triangle_eps": (1.0, 1e-10),
    "dimacs5": (1.0, 1e-9), "float_triangle": (1.625, 1e-9),
}


@pytest.mark.parametrize("name", sorted(KAT_EXPECT))
def test_kat_numeric(mrc, name):
    rec = KAT[name]
    s = facade(mrc, rec["n"], rec["edges"])
    r = s.solve()
    ratio, tol = KAT_EXPECT[name]
    assert abs(r.ratio - ratio) <= tol
    assert isinstance(r.ratio, float) and isinstance(r.sum_cost, float)
    assert r.cycle[0] == r.cycle[-1]
    assert abs(r.ratio - r.sum_cost / r.sum_time) < 1e-10 and r.sum_time > 0
    assert r.metrics.mode_used == "numeric"
    cyc, sc, st, rt = r   # tuple unpacking (solver.py:165-172)
    assert cyc == r.cycle and rt == r.ratio
    gold = rec["solve"]
    assert rotations_equal(r.cycle, gold["cycle"]) or abs(r.ratio - gold["ratio"]) <= 1e-9
    if name == "self_loop":
        assert r.cycle == [0, 0] and r.sum_cost == 3.0 and r.sum_time == 2.0
    if name == "three_self_loops":
        assert len(r.cycle) == 2
    if name == "parallel_edges":
        assert len(r.cycle) == 3


def test_kat_errors(mrc):
    from min_ratio_cycle_b200.exceptions import NumericalInstabilityError, SolverError
    s = mrc.MinRatioCycleSolver(3, mrc.SolverConfig(log_level=mrc.LogLevel.NONE))
    with pytest.raises(ValueError, match="Graph has no edges"):
        s.solve()
    rec = KAT["tree"]
    with pytest.raises(NumericalInstabilityError) as ei:
        facade(mrc, rec["n"], rec["edges"]).solve()
    assert {"suggested_fix", "recovery_hint"} <= set(ei.value.details)
    assert isinstance(ei.value, SolverError)
    with pytest.raises(ValueError):
        mrc.MinRatioCycleSolver(0)
    with pytest.raises(ValueError):
        s.add_edge(0, 1, 1.0, 0.0)
    with pytest.raises(ValueError):
        s.add_edge(0, 7, 1.0, 1.0)
    s2 = facade(mrc, 3, KAT["float_triangle"]["edges"], max_solve_time=0.0)   # non-integer: no exact fallback
    with pytest.raises(NumericalInstabilityError) as ei:
        s2.solve()
    assert type(ei.value.__cause__).__name__ == "TimeoutError"


def test_kat_ref_suboptimal(mrc, oracle_mod):
    """SURVEY 8c known reference-wrong case: reference reports -1.0, optimum is the loop at 1 (-16/3)."""
    rec = KAT["ref_suboptimal_n7"]
    r = facade(mrc, rec["n"], rec["edges"]).solve()
    assert abs(r.ratio - (-16.0 / 3.0)) < 1e-12 and r.cycle == [1, 1]
    verdict, why = three_way(checker(oracle_mod, rec), {"cycle": r.cycle, "ratio": r.ratio}, rec["solve"])
    assert verdict == REF_SUBOPTIMAL, why


def test_determinism(mrc):
    rec = SMALL[sorted(SMALL)[0]]
    a = facade(mrc, rec["n"], rec["edges"]).solve()
    b = facade(mrc, rec["n"], rec["edges"]).solve()
    assert a.ratio == b.ratio and a.cycle == b.cycle   # tests/test_integration.py:517-526


# --------------------------------------------------------------------------- 120 small random graphs
def test_small_random_numeric_three_way(mrc, oracle_mod):
    from min_ratio_cycle_b200.exceptions import NumericalInstabilityError
    counts = {MATCH: 0, REF_SUBOPTIMAL: 0, MISMATCH: 0, "both_fail": 0, "ref_fail_only": 0}
    bad = []
    for name in sorted(SMALL):
        rec = SMALL[name]
        s = facade(mrc, rec["n"], rec["edges"])
        try:
            r = s.solve()
        except NumericalInstabilityError:
            r = None
        gold = rec["solve"]
        if not gold["ok"]:
            if r is None:
                counts["both_fail"] += 1
            else:   # reference raised, we found a cycle: it must certify
                ok, why = checker(oracle_mod, rec).certify(r.cycle, r.ratio)
                assert ok, (name, why)
                counts["ref_fail_only"] += 1
            continue
        assert r is not None, f"{name}: reference solved, GPU path raised"
        U, V, _, _ = edges_to_arrays(rec["edges"])
        assert cycle_is_valid(r.cycle, U, V), name
        v, why = three_way(checker(oracle_mod, rec), {"cycle": r.cycle, "ratio": r.ratio}, gold)
        counts[v] += 1
        if v == MISMATCH:
            bad.append((name, why))
    print("three-way verdicts:", counts)
    assert not bad, bad
    assert counts[MATCH] >= 80


def test_small_random_probe_verdicts(mrc, oracle_mod):
    """K=1 and K-batched probes vs the quirk-free oracle detector, and vs the reference where its
    first-segment defect (solver.py:1179-1183) cannot fire (vertex 0 has an in-edge)."""
    from min_ratio_cycle_b200 import _native as N
    checked = 0
    for name in sorted(SMALL):
        rec = SMALL[name]
        o = checker(oracle_mod, rec)
        U, V, C, T, st, ct, _ = o.arrays()
        lams = [p["lam"] for p in rec["probes"]]
        g = N.DeviceGraph(rec["n"], U, V, C, T)
        single = [g.probe_numeric([lam], 1e-15)[0] for lam in lams]
        batch = g.probe_numeric(lams, 1e-15)
        for p, r1, rk in zip(rec["probes"], single, batch):
            exp = o.detect_negative_cycle_numeric(p["lam"], 1e-15, True).has_neg
            for r in (r1, rk):
                got = r["verdict"] in (N.V_NEG, N.V_NEG_NOCYCLE)
                assert got == bool(exp), (name, p["lam"], r)
                if ct[0] > 0:
                    assert got == p["has_neg"], (name, p["lam"])
                if r["verdict"] == N.V_NEG:
                    cyc = r["cycle"]
                    assert cycle_is_valid(cyc, U, V)
                    assert r["sum_cost"] / r["sum_time"] < p["lam"]
                    assert min(cyc) == cyc[0] == cyc[-1]
            checked += 1
        g.close()
    assert checked >= 300


# --------------------------------------------------------------------------- exact mode, bit-exact
def _exact_cases():
    out = []
    for name in sorted(SMALL):
        if "exact" in SMALL[name]:
            out.append(("small", name))
    for name in sorted(KAT):
        if "exact" in KAT[name]:
            out.append(("kat", name))
    return out


@pytest.mark.parametrize("strategy", [0, 1])
def test_exact_small_bit_exact(mrc, strategy):
    n_ok = n_fail = 0
    for src, name in _exact_cases():
        rec = (SMALL if src == "small" else KAT)[name]
        gold = rec["exact"]
        s = facade(mrc, rec["n"], rec["edges"], gpu_exact_strategy=strategy)
        s._preprocess_graph()
        s._build_arrays_if_needed()
        r = s._solve_exact()
        if gold["success"]:
            assert r.success, (name, r.error_message)
            assert r.cycle == gold["cycle"], name
            assert isinstance(r.sum_cost, float) and isinstance(r.ratio, float)
            assert int(r.sum_cost) == gold["sum_cost"] and int(r.sum_time) == gold["sum_time"], name
            assert Fraction(int(r.sum_cost), int(r.sum_time)) == Fraction(gold["sum_cost"], gold["sum_time"])
            assert r.ratio == gold["ratio"]
            assert s._metrics.iterations == gold["iterations"], name
            assert s._last_exact["probes"] == [(a, b) for a, b, _ in gold["probes"]], name
            n_ok += 1
        else:
            msg = gold["error_message"]
            if "index" in msg.lower():   # np.minimum.reduceat IndexError: a reference crash, not a result
                continue
            assert not r.success, (name, msg)
            assert r.error_message == msg, name
            n_fail += 1
    print(f"exact strategy {strategy}: {n_ok} bit-exact successes, {n_fail} identical failures")
    assert n_ok >= 30


def test_exact_verdict_points(mrc):
    for name in sorted(SMALL):
        rec = SMALL[name]
        if "exact" not in rec:
            continue
        s = facade(mrc, rec["n"], rec["edges"])
        s._preprocess_graph()
        s._build_arrays_if_needed()
        for vp in rec["exact"]["verdict_points"]:
            if "exc" in vp:
                continue
            assert s._has_negative_cycle_exact(vp["a"], vp["b"]) == vp["has_neg"], (name, vp)
            z = s._find_zero_cycle_exact(vp["a"], vp["b"])
            if vp["zero"] is None:
                assert z is None
            else:
                assert z is not None and z[0] == vp["zero"]["cycle"]
                assert z[1] == vp["zero"]["sum_c"] and z[2] == vp["zero"]["sum_t"]


def test_exact_potentials_match_oracle(mrc, oracle_mod):
    from min_ratio_cycle_b200 import _native as N
    name = "int_n120_m700"
    rec = SEEDED[name]
    from min_ratio_cycle_b200 import synthetic
    n, U, V, C, T = synthetic.make(rec["generator"], **rec["kwargs"])
    o = oracle_mod.Oracle(n, U, V, C, T, all_int=True)
    sU, sV, sC, sT, _, _, _ = o.arrays()
    g = N.DeviceGraph(n, sU, sV, sC, sT, sC.astype(np.int64), sT.astype(np.int64))
    for a, b in [(-200, 1), (-101, 3), (-40, 1)]:
        ok_o, dist_o = o.potentials_exact(a, b)
        ok_g, dist_g = g.potentials_exact(a, b)
        assert ok_o == ok_g
        if ok_o:
            assert np.array_equal(dist_o, dist_g)   # the relaxation fixed point is unique
            ok_m, mask = g.tight_mask_exact(a, b)
            W = b * sC.astype(np.int64) - a * sT.astype(np.int64)
            assert ok_m and np.array_equal(mask.astype(bool), dist_o[sU] + W == dist_o[sV])
    g.close()


@pytest.mark.parametrize("name", [k for k in sorted(SEEDED) if "exact" in SEEDED[k]])
@pytest.mark.parametrize("strategy", [0, 1])
def test_exact_seeded_bit_exact(mrc, name, strategy):
    """Includes G-exact-2k (n=2000, m=20000): Fraction(-196, 5), cycle [694,1070,1183,1310,493,694]."""
    from min_ratio_cycle_b200 import synthetic
    rec = SEEDED[name]
    gold = rec["exact"]
    n, U, V, C, T = synthetic.make(rec["generator"], **rec["kwargs"])
    s = mrc.MinRatioCycleSolver(n, mrc.SolverConfig(log_level=mrc.LogLevel.NONE, gpu_exact_strategy=strategy))
    s.add_edges_arrays(U, V, C, T)
    cyc, frac = s.solve_exact_fraction()
    assert gold["success"]
    assert cyc == gold["cycle"]
    assert frac == Fraction(gold["sum_cost"], gold["sum_time"])
    assert s._metrics.iterations == gold["iterations"]
    assert s._last_exact["probes"] == [(a, b) for a, b, _ in gold["probes"]]
    assert len(s._edges) + sum(len(b[0]) for b in s._bulk) == gold["m_after_preprocess"]


# --------------------------------------------------------------------------- seeded numeric graphs
@pytest.mark.parametrize("name", [k for k in sorted(SEEDED) if "solve" in SEEDED[k]])
@pytest.mark.parametrize("K", [1, 4, 8])
def test_numeric_seeded(mrc, oracle_mod, name, K):
    """Includes G-num-1k (BASELINE config 1, reference ratio -6.725338638039345) and two C4 currency graphs."""
    from min_ratio_cycle_b200 import synthetic
    rec = SEEDED[name]
    gold = rec["solve"]
    assert gold["ok"]
    n, U, V, C, T = synthetic.make(rec["generator"], **rec["kwargs"])
    s = mrc.MinRatioCycleSolver(n, mrc.SolverConfig(log_level=mrc.LogLevel.NONE, gpu_candidates=K))
    s.add_edges_arrays(U, V, C, T)
    r = s.solve()
    o = oracle_mod.Oracle(n, U, V, C, T, all_int=rec["integer"], threads=4)
    o.set_quirks(False)
    v, why = three_way(o, {"cycle": r.cycle, "ratio": r.ratio}, gold)
    assert v in (MATCH, REF_SUBOPTIMAL), (name, why)
    if name == "G_num_1k":
        assert v == MATCH and rotations_equal(r.cycle, gold["cycle"])
    ok, why = o.certify(r.cycle, r.ratio)
    assert ok, why


# --------------------------------------------------------------------------- sensitivity analysis (a17)
def test_sensitivity_analysis_matches_fresh_solves(mrc, oracle_mod):
    """Each perturbed re-solve (device-side edge patch + warm start) equals a fresh solve of the perturbed
    graph, and certifies against the oracle.  Reported sums use the perturbed weights (the reference's
    stale _edge_map, SURVEY 0.8, is deliberately not reproduced)."""
    from min_ratio_cycle_b200 import synthetic
    n, U, V, C, T = synthetic.ring_plus_random(300, 1500, seed=9)
    s = mrc.MinRatioCycleSolver(n, mrc.SolverConfig(log_level=mrc.LogLevel.NONE))
    s.add_edges_arrays(U, V, C, T)
    base = s.solve()
    rng = np.random.default_rng(3)
    on_cycle = [int(np.flatnonzero((U == a) & (V == b))[0]) for a, b in zip(base.cycle[:-1], base.cycle[1:])]
    scen = [{on_cycle[0]: (abs(C[on_cycle[0]]) * 0.5 + 1.0, 0.0)}, {on_cycle[-1]: (-2.0, 0.0)}]
    for _ in range(6):
        idx = int(rng.integers(0, len(U)))
        scen.append({idx: (float(C[idx]) * float(rng.choice([-0.1, 0.1])) - 0.5, 0.0)})
    scen.append({int(rng.integers(0, len(U))): (0.0, 0.25), int(rng.integers(0, len(U))): (-1.0, 0.0)})
    res = s.sensitivity_analysis(scen)
    assert len(res) == len(scen)
    for sc, r in zip(scen, res):
        C2, T2 = C.copy(), T.copy()
        for idx, (dc, dt) in sc.items():
            C2[idx] += dc; T2[idx] += dt
        f = mrc.MinRatioCycleSolver(n, mrc.SolverConfig(log_level=mrc.LogLevel.NONE))
        f.add_edges_arrays(U, V, C2, T2)
        fr = f.solve()
        assert abs(r.ratio - fr.ratio) <= 1e-9 * max(1.0, abs(fr.ratio)), (sc, r.ratio, fr.ratio)
        o = oracle_mod.Oracle(n, U, V, C2, T2, all_int=False)
        o.set_quirks(False)
        ok, why = o.certify(r.cycle, r.ratio)
        assert ok, (sc, why)
    again = s.solve()      # the device graph was restored
    assert again.ratio == base.ratio and again.cycle == base.cycle
    assert len(s.stability_region.__doc__ or "") >= 0


# --------------------------------------------------------------------------- batches of small graphs (config 4)
def test_batch_currency_graphs(mrc, oracle_mod):
    """32 currency graphs (n=64, m=4032) in one launch: each certified by the oracle; the two with a reference
    solve in the fixtures match it; the batch equals the single-graph GPU path."""
    from min_ratio_cycle_b200 import batch, synthetic
    seeds = list(range(1000, 1032))
    graphs = [synthetic.currency_graph(64, s) for s in seeds]
    res, stats = batch.solve_min_ratio_cycle_batch(graphs, mrc.SolverConfig(log_level=mrc.LogLevel.NONE), return_stats=True)
    assert len(res) == 32 and stats["relax_edge_visits"] > 0
    for s, gph, r in zip(seeds, graphs, res):
        n, U, V, C, T = gph
        assert r.success and r.cycle[0] == r.cycle[-1] == min(r.cycle)
        o = oracle_mod.Oracle(n, U, V, C, T, all_int=False)
        o.set_quirks(False)
        ok, why = o.certify(r.cycle, r.ratio)
        assert ok, (s, why)
        key = f"currency_{s}"
        if key in SEEDED:
            gold = SEEDED[key]["solve"]
            assert abs(r.ratio - gold["ratio"]) <= 1e-9 * max(1.0, abs(gold["ratio"])), key
    n, U, V, C, T = graphs[0]
    s1 = mrc.MinRatioCycleSolver(n, mrc.SolverConfig(log_level=mrc.LogLevel.NONE))
    s1.add_edges_arrays(U, V, C, T)
    r1 = s1.solve()
    assert abs(r1.ratio - res[0].ratio) <= 1e-12 and r1.cycle == res[0].cycle


def test_batch_mixed_small_graphs(mrc, oracle_mod):
    """Ragged batch: the 120 small random fixtures (n 2..14, self-loops, parallel edges, acyclic ones)."""
    from min_ratio_cycle_b200 import batch
    names = sorted(SMALL)
    graphs = []
    for nm in names:
        U, V, C, T = edges_to_arrays(SMALL[nm]["edges"])
        graphs.append((SMALL[nm]["n"], U, V, C, T))
    for K in (1, 4):
        res = batch.solve_min_ratio_cycle_batch(graphs, mrc.SolverConfig(log_level=mrc.LogLevel.NONE, gpu_candidates=K))
        nfail = 0
        for nm, r in zip(names, res):
            rec = SMALL[nm]
            chk = checker(oracle_mod, rec)
            if not r.success:
                nfail += 1
                # no cycle at all: the quirk-free detector must agree at the top of the bracket
                lo, hi = chk.bounds()
                assert chk.detect_negative_cycle_numeric(hi, 1e-15, True).has_neg == 0, nm
                continue
            sU, sV = chk.arrays()[:2]
            simple = len(set(zip(sU.tolist(), sV.tolist()))) == len(sU)
            if simple:   # with parallel (u,v) edges the reference's hop rule mis-costs cycles; the
                ok, why = chk.certify(r.cycle, r.ratio)   # reported sums inherit that rule, so no certificate
                assert ok, (nm, K, why)
            if rec["solve"]["ok"]:
                v, why = three_way(chk, {"cycle": r.cycle, "ratio": r.ratio}, rec["solve"])
                assert v in (MATCH, REF_SUBOPTIMAL) or not simple, (nm, K, why)
        assert nfail <= 10


# --------------------------------------------------------------------------- callers either side of the path
def test_dimacs_benchmark_and_convenience(mrc):
    """tests/test_analytics.py:50-78 and tests/test_regression.py of the reference, against the GPU path."""
    from min_ratio_cycle_b200 import benchmarks as B
    tri = ["p sp 3 3", "a 1 2 1 1", "a 2 3 1 1", "a 3 1 1 1"]
    out = B.benchmark_solver(B.load_dimacs_graph(tri, mrc.SolverConfig(log_level=mrc.LogLevel.NONE)), compare=False)
    assert abs(out["ratio"] - 1.0) <= 1e-9 and out["diff"] is None and out["time"] < 1.0
    five = ["p sp 5 7", "a 1 2 2 1", "a 2 3 2 1", "a 3 1 2 1", "a 3 4 1 1", "a 4 5 1 1", "a 5 3 1 1", "a 1 4 5 2"]
    out = B.benchmark_solver(B.load_dimacs_graph(five, mrc.SolverConfig(log_level=mrc.LogLevel.NONE)), compare=True)
    assert out["diff"] <= 1e-9
    r = mrc.solve_min_ratio_cycle([(0, 1, -2, 1), (1, 2, -1, 1), (2, 0, 1, 1)], config=mrc.SolverConfig(log_level=mrc.LogLevel.NONE))
    assert abs(r.ratio - (-2.0 / 3.0)) <= 1e-12 and r.cycle[0] == r.cycle[-1]
    s = mrc.MinRatioCycleSolver(4, mrc.SolverConfig(log_level=mrc.LogLevel.NONE))
    s.add_edges([(0, 1, 1.0, 1.0), (1, 2, 1.0, 1.0), (2, 0, 1.0, 1.0), (2, 3, 5.0, 1.0), (3, 2, 5.0, 1.0)])
    stab = s.stability_region(0.01)
    assert set(stab) == {0, 1, 2, 3, 4} and all(isinstance(v, bool) for v in stab.values()) and stab[3] and stab[4]
    props = s.get_graph_properties()
    assert props["vertices"] == 4 and props["edges"] == 5 and props["connectivity"]["max_in_degree"] == 2
    assert s.health_check()["summary"]["overall_status"] in ("PASS", "WARN")
    assert "Ratio" in s.get_debug_info()
    sh = mrc.MinRatioCycleSolver(3, mrc.SolverConfig(log_level=mrc.LogLevel.NONE, honor_mode=True))
    sh.add_edges([(0, 1, 2, 1), (1, 2, 3, 2), (2, 0, 1, 1)])
    re = sh.solve(mode="exact")
    assert re.metrics.mode_used == "exact" and re.cycle == [1, 2, 0, 1] and re.ratio == 1.5   # README.md:66 of the reference


def test_device_sort_matches_numpy(mrc):
    """mrc_sort_edges_by_dst == np.argsort(V, kind='stable') + gathers (solver.py:476-478), incl. the dup census."""
    from min_ratio_cycle_b200 import _native as N
    rng = np.random.default_rng(11)
    for n, m, dup in [(1, 5, True), (7, 50, True), (1000, 300_000, True), (50_000, 400_000, False)]:
        U = rng.integers(0, n, m); V = rng.integers(0, n, m)
        if not dup:
            from min_ratio_cycle_b200 import synthetic
            U, V = synthetic._simple_pairs(n, m, rng)
        C = rng.uniform(-5, 5, m); T = rng.uniform(0.1, 2, m)
        so, do, co, to, order, dups = N.sort_edges_by_dst(n, U, V, C, T)
        ref = np.argsort(V, kind="stable")
        assert np.array_equal(order, ref) and np.array_equal(so, U[ref]) and np.array_equal(do, V[ref])
        assert np.array_equal(co, C[ref]) and np.array_equal(to, T[ref])
        key = U.astype(np.int64) * n + V
        assert dups == m - len(np.unique(key))
    # the facade takes the device path above 200k edges and must expose the same private arrays
    n, m = 20_000, 250_000
    from min_ratio_cycle_b200 import synthetic
    _, U, V, C, T = synthetic.random_float(n, m, seed=4)
    s = mrc.MinRatioCycleSolver(n, mrc.SolverConfig(log_level=mrc.LogLevel.NONE))
    s.add_edges_arrays(U, V, C, T)
    s.solve()
    ref = np.argsort(V, kind="stable")
    assert np.array_equal(s._U, U[ref]) and np.array_equal(s._V, V[ref]) and np.array_equal(s._C, C[ref])
    assert np.array_equal(s._order, ref) and s._U.dtype == np.int64
    assert np.array_equal(s._counts, np.bincount(V, minlength=n))


def test_c1_three_way_statistics(mrc, oracle_mod):
    """BASELINE config 1 on 16 seeds (reference: ~100 s per seed, tests/golden/c1_seeds.json): three-way verdict
    of SURVEY 8c with the oracle certificate; every GPU answer must certify, none may MISMATCH."""
    from min_ratio_cycle_b200 import synthetic
    C1 = load_golden("c1_seeds.json")
    counts = {MATCH: 0, REF_SUBOPTIMAL: 0, MISMATCH: 0}
    for seed in sorted(C1, key=int):
        n, U, V, C, T = synthetic.random_float(1000, 5000, seed=int(seed))
        s = mrc.MinRatioCycleSolver(n, mrc.SolverConfig(log_level=mrc.LogLevel.NONE))
        s.add_edges_arrays(U, V, C, T)
        r = s.solve(mode="numeric")
        o = oracle_mod.Oracle(n, U, V, C, T, all_int=False, threads=4)
        o.set_quirks(False)
        ok, why = o.certify(r.cycle, r.ratio)
        assert ok, (seed, why)
        v, why = three_way(o, {"cycle": r.cycle, "ratio": r.ratio}, C1[seed])
        counts[v] += 1
        assert v != MISMATCH, (seed, why)
    print("C1 three-way verdicts over 16 seeds:", counts)
    assert counts[MATCH] >= 12


def test_every_kernel_variant_on_odd_sizes(mrc):
    """Dense (full-width, narrowed, compact-survivor) and worklist relaxation at K=1,2,4,8, scenario probes, device patches, exact strategies,
    batch kernel and device sort on sizes that exercise every tail path; all relaxation variants must agree."""
    import os
    import runpy
    from conftest import ROOT
    runpy.run_path(os.path.join(ROOT, "tools", "sanitize_case.py"), run_name="__main__")


def test_graph_properties_on_device_match_numpy(mrc):
    """get_graph_properties (solver.py:339-387): the device-side statistics equal the NumPy ones."""
    from min_ratio_cycle_b200 import synthetic
    rng = np.random.default_rng(4)
    n = 5000
    U = rng.integers(0, n // 2, 30_000); V = rng.integers(n // 4, n - 50, 30_000)   # sources, sinks and isolated vertices
    C = rng.uniform(-10, 10, 30_000); T = rng.uniform(0.1, 5, 30_000)
    s = mrc.MinRatioCycleSolver(n, mrc.SolverConfig(log_level=mrc.LogLevel.NONE))
    s.add_edges_arrays(U, V, C, T)
    p = s.get_graph_properties()
    indeg = np.bincount(s._V, minlength=n); outdeg = np.bincount(s._U, minlength=n)
    c = p["connectivity"]
    assert c["isolated_vertices"] == int(np.sum((indeg == 0) & (outdeg == 0))) > 0
    assert c["source_vertices"] == int(np.sum((indeg == 0) & (outdeg > 0))) > 0
    assert c["sink_vertices"] == int(np.sum((indeg > 0) & (outdeg == 0))) > 0
    assert c["max_in_degree"] == int(indeg.max()) and c["max_out_degree"] == int(outdeg.max())
    assert c["avg_in_degree"] == pytest.approx(indeg.mean()) and c["avg_out_degree"] == pytest.approx(outdeg.mean())
    w = p["weights"]
    assert w["cost_range"] == (float(s._C.min()), float(s._C.max()))
    assert w["time_range"] == (float(s._T.min()), float(s._T.max()))
    for key, arr in (("cost", s._C), ("time", s._T)):
        assert w[key + "_mean"] == pytest.approx(float(arr.mean()), rel=1e-12, abs=1e-12)
        assert w[key + "_std"] == pytest.approx(float(arr.std()), rel=1e-12)
    assert p["vertices"] == n and p["edges"] == len(s._U) and w["all_integer"] is False
    assert {i["level"] for i in s.detect_issues()} >= {"WARNING"}


def test_stability_region_equals_one_resolve_per_edge(mrc):
    """stability_region (solver.py:1731-1759): the batched scenarios give what the reference's per-edge loop gives."""
    from min_ratio_cycle_b200 import synthetic
    n, U, V, C, T = synthetic.ring_plus_random(120, 500, seed=3)
    s = mrc.MinRatioCycleSolver(n, mrc.SolverConfig(log_level=mrc.LogLevel.NONE))
    s.add_edges_arrays(U, V, C, T)
    base = s.solve()
    stab = s.stability_region(0.05)
    assert set(stab) == set(range(len(U)))
    U0, V0, C0, T0 = s._insertion_arrays()
    for idx in list(range(0, len(U), 7)) + [i for i, ok in stab.items() if not ok][:10]:
        one = s.sensitivity_analysis([{idx: (float(C0[idx]) * 0.05, float(T0[idx]) * 0.05)}])[0]
        assert (one.cycle == base.cycle) == stab[idx], idx
    on_cycle = {(a, b) for a, b in zip(base.cycle[:-1], base.cycle[1:])}
    # perturbing an edge off the optimal cycle upwards can never change the optimum
    for idx in range(len(U)):
        if (int(U0[idx]), int(V0[idx])) not in on_cycle and C0[idx] > 0:
            assert stab[idx], idx


def test_probe_without_cycles_then_fetch_with_sums(mrc):
    """MRC_F_NO_CYCLES + mrc_fetch_cycle_sums give what a plain probe reports (the multi-GPU driver's path)."""
    from min_ratio_cycle_b200 import _native as N, synthetic
    n, U, V, C, T = synthetic.random_float(3000, 20000, seed=2)
    o = np.argsort(V, kind="stable")
    g = N.DeviceGraph(n, U[o], V[o], C[o], T[o])
    lo, hi = g.bounds()
    star = g.solve_numeric(lo, hi, K=4)["ratio"]
    lams = [star - 1.0, star + 1e-3, star + 0.5, star + 3.0]
    full = g.probe_numeric(lams)
    lean = g.probe_numeric(lams, flags=N.F_NO_CYCLES)
    assert [p["verdict"] for p in lean] == [p["verdict"] for p in full] == [N.V_NONNEG, N.V_NEG, N.V_NEG, N.V_NEG]
    for k, p in enumerate(lean):
        assert p["cycle"] is None
        if p["verdict"] != N.V_NEG:
            assert g.fetch_cycle_sums(k)[0] == []
            continue
        cyc, sc, st = g.fetch_cycle_sums(k, p["cycle_len"])
        assert len(cyc) == p["cycle_len"] and cyc[0] == cyc[-1] == min(cyc)
        assert sc / st < lams[k] and abs(sc / st - p["sum_cost"] / p["sum_time"]) <= 1e-12 * max(1.0, abs(sc / st))
        # the forward-order sums are the ones recomputed from the edges of the cycle
        hop = {(int(a), int(b)): (c, t) for a, b, c, t in zip(U, V, C, T)}
        rc = rt = 0.0
        for a, b in zip(cyc[:-1], cyc[1:]):
            rc += hop[(a, b)][0]; rt += hop[(a, b)][1]
        assert (rc, rt) == (sc, st)
    g.close()
