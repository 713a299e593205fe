"""
GPU tests (run with -m gpu) of the two round-2 additions to the probe engine, through the C ABI:

  persistent probe kernel (csrc/mrc_probe.cuh, option "persist"): one cooperative launch per probe with the pass loop, the
      check schedule, narrowing, the compact survivor, the worklist switch and pruning by monotonicity on the device.  It must
      give the verdicts, cycles and ratios of the burst driver (host round trip per burst) on the same probes.
  quantile first round (option "qstart"): the first round of a search without a known cycle places its candidates at low
      quantiles of the edge-ratio sample; any candidate set is legitimate, so the optimum must not move.

Reference being replaced in both cases: the pass loop of _detect_negative_cycle_numeric (min_ratio_cycle/solver.py:1124-1294)
and the bisection of _solve_numeric (:1017-1117).
"""

from __future__ import annotations

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _sorted(n, U, V, C, T):
    o = np.argsort(V, kind="stable")
    return n, U[o].astype(np.int32), V[o].astype(np.int32), C[o], T[o]


def _graph(N, args, **opts):
    g = N.DeviceGraph(*args)
    for k, v in opts.items():
        g.set_option(k, v)
    return g


@pytest.mark.parametrize("K", [1, 2, 3, 4])
@pytest.mark.parametrize("sparse", [0, 1])
def test_persistent_probe_matches_burst_driver(K, sparse, oracle_mod):
    from min_ratio_cycle_b200 import _native as N, synthetic
    for seed, (n, m) in enumerate([(3000, 20000), (20000, 200000), (500, 6000)]):
        raw = synthetic.random_float(n, m, seed=40 + seed)
        args = _sorted(*raw)
        common = dict(sparse=sparse, sparse_min_edges=0, sparse_div=2, compact_min_edges=0, qstart_min_edges=0)
        gb = _graph(N, args, persist=0, **common)
        gp = _graph(N, args, persist=1, **common)
        lo, hi = gb.bounds()
        a = gb.solve_numeric(lo, hi, K=K)
        gp.reset_stats()
        b = gp.solve_numeric(lo, hi, K=K)
        st = gp.stats()
        assert a["rc"] == b["rc"] == N.OK
        assert abs(a["ratio"] - b["ratio"]) <= 1e-12 * max(1.0, abs(a["ratio"])), (seed, K, a["ratio"], b["ratio"])
        o = oracle_mod.Oracle(*raw, all_int=False)
        o.set_quirks(False)
        ok, why = o.certify([int(x) for x in b["cycle"]], b["ratio"])
        assert ok, (seed, K, why)
        # one launch per probe, whatever the number of sweeps, checks and worklist passes inside it
        assert st["probe_launches"] >= b["iterations"] and st["relax_launches"] > st["probe_launches"]
        assert st["probe_ms"] > 0 and st["relax_ms"] > 0 and st["check_launches"] > 0
        if sparse and m >= 20000:
            assert st["sparse_passes"] > 0
        # verdicts of single probes around the optimum, both drivers
        r = a["ratio"]
        lams = [r - 1.0, r - 1e-6, r + 1e-6, r + 1.0][:K] if K > 1 else [r - 1e-6]
        va = [p["verdict"] for p in gb.probe_numeric(lams)]
        vb = [p["verdict"] for p in gp.probe_numeric(lams)]
        assert va == vb, (seed, K, va, vb)
        gb.close(); gp.close()


def test_persistent_probe_exact_bit_exact(oracle_mod):
    """int64 probes in the persistent kernel: potentials and Stern-Brocot results identical to the oracle's restatement."""
    import min_ratio_cycle_b200.solver as S
    from min_ratio_cycle_b200 import _native as N, synthetic
    for seed in (3, 4):
        n, U, V, C, T = synthetic.random_int(1500, 9000, seed=seed)
        x = oracle_mod.Oracle(n, U, V, C, T, all_int=True).solve_exact()
        for strategy in (1, 0):
            s = S.MinRatioCycleSolver(n, S.SolverConfig(log_level=S.LogLevel.NONE, gpu_exact_strategy=strategy))
            s.add_edges_arrays(U, V, C, T)
            cyc, frac = s.solve_exact_fraction()
            assert cyc == x.cycle and frac == x.fraction
            assert s._dev.stats()["probe_launches"] > 0      # exact probes are single-candidate: persistent by default
        o = oracle_mod.Oracle(n, U, V, C, T, all_int=True)
        sU, sV, sC, sT, _, _, _ = o.arrays()
        g = N.DeviceGraph(n, sU, sV, sC, sT, sC.astype(np.int64), sT.astype(np.int64))
        a, b = x.ab
        for aa, bb in ((a, b), (a - 1, b), (2 * a - 3, 2 * b)):
            ok_o, dist_o = o.potentials_exact(aa, bb)
            ok_g, dist_g = g.potentials_exact(aa, bb)
            assert ok_o == ok_g and (not ok_o or np.array_equal(dist_o, dist_g))   # the relaxation fixed point is unique
        assert g.stats()["probe_launches"] > 0
        g.close()


def test_quantile_first_round_keeps_the_optimum(oracle_mod):
    from min_ratio_cycle_b200 import _native as N, synthetic
    raw = synthetic.random_float(30000, 300000, seed=77)
    args = _sorted(*raw)
    o = oracle_mod.Oracle(*raw, all_int=False)
    o.set_quirks(False)
    for K in (1, 2, 4, 8):
        res = {}
        for q in (0, 1):
            g = _graph(N, args, qstart=q)
            lo, hi = g.bounds()
            res[q] = g.solve_numeric(lo, hi, K=K)
            g.close()
            assert res[q]["rc"] == N.OK
        assert abs(res[0]["ratio"] - res[1]["ratio"]) <= 1e-9 * max(1.0, abs(res[0]["ratio"])), (K, res[0]["ratio"], res[1]["ratio"])
        ok, why = o.certify([int(x) for x in res[1]["cycle"]], res[1]["ratio"])
        assert ok, (K, why)
        assert res[1]["iterations"] <= res[0]["iterations"] + 1, (K, res[0]["iterations"], res[1]["iterations"])


def test_quantile_first_round_on_degenerate_ratio_distributions():
    """Ties in the edge-ratio sample (unit times, few distinct costs) and optima far from the low tail: the search falls back
    to the uniform plan or simply loses a round -- never the answer."""
    from min_ratio_cycle_b200 import _native as N, synthetic
    rng = np.random.default_rng(5)
    n, m = 5000, 40000
    _, U, V, _, _ = synthetic.random_float(n, m, seed=9)
    cases = {
        "three_costs_unit_time": (rng.choice([-1.0, 0.0, 2.0], len(U)), np.ones(len(U))),
        "all_positive": (rng.uniform(1.0, 2.0, len(U)), rng.uniform(0.5, 1.0, len(U))),
        "constant_ratio": (np.full(len(U), 3.0), np.full(len(U), 2.0)),
    }
    for name, (C, T) in cases.items():
        args = _sorted(n, U, V, C, T)
        out = {}
        for q in (0, 1):
            g = _graph(N, args, qstart=q, qstart_min_edges=0)
            lo, hi = g.bounds()
            out[q] = g.solve_numeric(lo, hi, K=4)
            g.close()
        assert out[0]["rc"] == out[1]["rc"], (name, out[0]["rc"], out[1]["rc"])
        if out[0]["rc"] == N.OK:
            assert abs(out[0]["ratio"] - out[1]["ratio"]) <= 1e-9 * max(1.0, abs(out[0]["ratio"])), (name, out[0]["ratio"], out[1]["ratio"])
