"""
Generates the golden fixtures in this directory by running the UNMODIFIED Python
reference (DiogoRibeiro7/min_ratio_cycle, mounted read-only at /root/reference).

    python tests/golden/make_golden.py [--skip-slow]

The reference cannot travel to the GPU box, so its outputs are committed here as
small JSON files; tests compare the CPU oracle (oracle/) and the CUDA path against
them.  `import min_ratio_cycle` needs matplotlib (absent here, only used for
plotting, solver.py:24) -- a two-class stub is injected into sys.modules first.

Python floats are written with repr() precision by json, so they round-trip bit-exactly.
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("MRC_REFERENCE", "/root/reference")


def import_reference():
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
            import matplotlib.pyplot  # noqa: F401
        except Exception:
            mpl = types.ModuleType("matplotlib")
            mpl.use = lambda *a, **k: None
            plt = types.ModuleType("matplotlib.pyplot")
            plt.Figure = type("Figure", (), {})
            plt.Axes = type("Axes", (), {})
            mpl.pyplot = plt
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pyplot"] = plt
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import min_ratio_cycle.solver as S
    return S


sys.path.insert(0, ROOT)
from min_ratio_cycle_b200 import synthetic  # noqa: E402


def py_edges(U, V, C, T, integer):
    out = []
    for u, v, c, t in zip(U, V, C, T):
        if integer:
            out.append((int(u), int(v), int(c), int(t)))
        else:
            out.append((int(u), int(v), float(c), float(t)))
    return out


def new_solver(S, n, edges, **cfg):
    s = S.MinRatioCycleSolver(int(n), S.SolverConfig(log_level=S.LogLevel.NONE, **cfg))
    for e in edges:
        s.add_edge(*e)
    return s


def run_solve(S, n, edges, **kw):
    s = new_solver(S, n, edges)
    t0 = time.time()
    try:
        r = s.solve(**kw)
        return {
            "ok": True, "cycle": [int(x) for x in r.cycle], "sum_cost": float(r.sum_cost),
            "sum_time": float(r.sum_time), "ratio": float(r.ratio),
            "iterations": int(r.metrics.iterations), "mode_used": r.metrics.mode_used,
            "m_after_preprocess": len(s._edges), "seconds": round(time.time() - t0, 3),
        }
    except Exception as e:  # noqa: BLE001
        d = {"ok": False, "exc": type(e).__name__, "msg": str(e)[:200]}
        if getattr(e, "__cause__", None) is not None:
            d["cause"] = type(e.__cause__).__name__
        if hasattr(e, "details"):
            d["detail_keys"] = sorted(e.details.keys())
        return d


def run_probes(S, n, edges, lams, slack=1e-15):
    s = new_solver(S, n, edges)
    s._preprocess_graph()
    s._build_arrays_if_needed()
    out = []
    for lam in lams:
        has_neg, data = s._detect_negative_cycle_numeric(float(lam), slack)
        rec = {"lam": float(lam), "has_neg": bool(has_neg), "data": None}
        if data is not None:
            cyc, sc, st, r = data
            rec["data"] = {"cycle": [int(x) for x in cyc], "sum_cost": float(sc), "sum_time": float(st),
                           "ratio": float(r)}
        out.append(rec)
    return out


def run_exact(S, n, edges, verdict_points=()):
    s = new_solver(S, n, edges)
    s._preprocess_graph()
    s._build_arrays_if_needed()
    probes = []
    orig = s._has_negative_cycle_exact

    def rec(a, b):
        v = orig(a, b)
        probes.append([int(a), int(b), bool(v)])
        return v

    s._has_negative_cycle_exact = rec
    r = s._solve_exact()
    d = {"success": bool(r.success), "probes": probes, "m_after_preprocess": len(s._edges)}
    if r.success:
        d.update(cycle=[int(x) for x in r.cycle], sum_cost=int(r.sum_cost), sum_time=int(r.sum_time),
                 ratio=float(r.ratio), iterations=int(s._metrics.iterations))
    else:
        d["error_message"] = r.error_message[:200]
    s._has_negative_cycle_exact = orig
    vp = []
    for a, b in verdict_points:
        try:
            hn = bool(s._has_negative_cycle_exact(a, b))
            z = s._find_zero_cycle_exact(a, b)
            vp.append({"a": a, "b": b, "has_neg": hn,
                       "zero": None if z is None else {"cycle": [int(x) for x in z[0]], "sum_c": int(z[1]),
                                                       "sum_t": int(z[2])}})
        except Exception as e:  # noqa: BLE001
            vp.append({"a": a, "b": b, "exc": type(e).__name__})
    d["verdict_points"] = vp
    return d


def sorted_arrays(S, n, edges):
    s = new_solver(S, n, edges)
    s._preprocess_graph()
    s._build_arrays_if_needed()
    return {"U": s._U.tolist(), "V": s._V.tolist(), "C": s._C.tolist(), "T": s._T.tolist(),
            "starts": s._starts.tolist(), "counts": s._counts.tolist(), "all_int": bool(s._all_int)}


# ----------------------------------------------------------------------------- KATs
def kat_cases():
    """Known-answer inputs held by the reference's own tests."""
    dis = [(0, 1, 1, 1), (1, 2, 1, 1), (2, 0, 1, 1), (3, 4, 2, 5), (4, 3, 2, 5), (5, 6, 1, 1), (6, 7, 1, 1)]
    return {
        "triangle": (3, [(0, 1, 2, 1), (1, 2, 3, 2), (2, 0, 1, 1)]),                 # tests/conftest.py:11-21
        "negative_cycle": (3, [(0, 1, -1, 1), (1, 2, -2, 1), (2, 0, 1, 1)]),          # tests/conftest.py:24-34
        "self_loop": (1, [(0, 0, 3, 2)]),                                            # tests/test_solver.py:69-80
        "parallel_edges": (2, [(0, 1, 10, 2), (0, 1, 6, 3), (1, 0, 1, 1)]),           # tests/test_solver.py:82-96
        "two_components": (5, [(0, 1, 1, 1), (1, 2, 1, 1), (2, 0, 1, 1), (3, 4, 1, 2), (4, 3, 1, 2)]),
        "disconnected8": (8, dis),                                                   # tests/conftest.py:92-111
        "three_self_loops": (3, [(0, 0, 5, 2), (1, 1, 3, 2), (2, 2, 7, 4)]),          # tests/test_integration.py:306-322
        "unit_triangle": (3, [(0, 1, 1.0, 1.0), (1, 2, 1.0, 1.0), (2, 0, 1.0, 1.0)]),  # tests/baselines/triangle.json
        "unit_triangle_eps": (3, [(0, 1, 1.0 + 1e-12, 1.0), (1, 2, 1.0 - 1e-12, 1.0), (2, 0, 1.0, 1.0)]),
        "float_triangle": (3, [(0, 1, 2.5, 1.0), (1, 2, 3.0, 2.0), (2, 0, 1.0, 1.0)]),
        "tree": (4, [(0, 1, 1, 1), (0, 2, 1, 1), (1, 3, 1, 1)]),                      # tests/test_solver.py:54-67
        "huge_weights": (3, [(0, 1, 10**15, 1), (1, 2, 10**15, 1), (2, 0, -2 * 10**15, 1)]),
        "ref_suboptimal_n7": (7, [(5, 4, 43, 3), (0, 0, -14, 14), (0, 6, 88, 15), (0, 6, 51, 3), (6, 0, 41, 10),
                                  (4, 1, 12, 6), (1, 1, -32, 6), (6, 5, -64, 9), (2, 3, -46, 4), (3, 4, -81, 14)]),
        "dimacs5": (5, [(0, 1, 2.0, 1.0), (1, 2, 2.0, 1.0), (2, 0, 2.0, 1.0), (2, 3, 1.0, 1.0), (3, 4, 1.0, 1.0),
                        (4, 2, 1.0, 1.0), (0, 3, 5.0, 2.0)]),
    }


def gen_kat(S):
    out = {}
    for name, (n, edges) in kat_cases().items():
        out[name] = {"n": n, "edges": [list(e) for e in edges], "solve": run_solve(S, n, edges)}
        integer = all(isinstance(e[2], int) and isinstance(e[3], int) for e in edges)
        if integer and name not in ("tree",):
            out[name]["exact"] = run_exact(S, n, edges)
        out[name]["arrays"] = sorted_arrays(S, n, edges)
    return out


# ------------------------------------------------------------------- small random graphs
def small_random_cases():
    rng = np.random.default_rng(20261019)
    cases = []
    for i in range(120):
        n = int(rng.integers(2, 15))
        kind = ["simple_float", "simple_int", "raw_float", "raw_int"][i % 4]
        if kind.startswith("simple"):
            extra = int(rng.integers(0, min(2 * n, n * (n - 1) - n) + 1)) if n > 2 else 0
            _, U, V, C, T = synthetic.ring_plus_random(n, extra, seed=1000 + i, integer=kind.endswith("int"))
        else:
            m = int(rng.integers(n, 3 * n + 2))
            U = rng.integers(0, n, m)
            V = rng.integers(0, n, m)
            if kind.endswith("int"):
                C = rng.integers(-100, 101, m).astype(float)
                T = rng.integers(1, 21, m).astype(float)
            else:
                C = rng.uniform(-10, 10, m)
                T = rng.uniform(0.1, 5, m)
        cases.append((f"{kind}_{i}", n, py_edges(U, V, C, T, kind.endswith("int")), kind))
    return cases


def gen_small(S):
    out = {}
    for name, n, edges, kind in small_random_cases():
        rec = {"n": n, "edges": [list(e) for e in edges], "kind": kind, "solve": run_solve(S, n, edges)}
        lams = []
        if rec["solve"]["ok"]:
            r = rec["solve"]["ratio"]
            lams = [r - 0.5, r - 1e-6, r + 1e-6, r + 0.5, r + 3.0]
        else:
            lams = [-1.0, 0.0, 1.0]
        rec["probes"] = run_probes(S, n, edges, lams)
        if kind.endswith("int"):
            rec["exact"] = run_exact(S, n, edges, verdict_points=[(-5, 1), (0, 1), (3, 2), (7, 3)])
        out[name] = rec
    return out


# ------------------------------------------------------------------- seeded (generator) graphs
def seeded_cases(skip_slow):
    cases = [
        ("float_n60_m300", "random_float", dict(n=60, m=300, seed=1), False, "numeric"),
        ("float_n100_m500", "random_float", dict(n=100, m=500, seed=2), False, "numeric"),
        ("float_n100_m1000", "random_float", dict(n=100, m=1000, seed=3), False, "numeric"),
        ("ring_float_n80", "ring_plus_random", dict(n=80, extra=240, seed=4), False, "numeric"),
        ("float_n200_m1000", "random_float", dict(n=200, m=1000, seed=0), False, "numeric"),
        ("currency_1000", "currency_graph", dict(n=64, seed=1000), False, "numeric"),
        ("currency_1001", "currency_graph", dict(n=64, seed=1001), False, "numeric"),
        ("int_n50_m250", "random_int", dict(n=50, m=250, seed=1), True, "both"),
        ("int_n120_m700", "random_int", dict(n=120, m=700, seed=2), True, "both"),
        ("int_n300_m1800", "random_int", dict(n=300, m=1800, seed=0), True, "exact"),
        ("int_n2000_m20000", "random_int", dict(n=2000, m=20000, seed=0), True, "exact"),
    ]
    if not skip_slow:
        cases.append(("G_num_1k", "raw_uniform", dict(n=1000, m=5000, seed=0), False, "numeric"))
        cases.append(("float_n1000_m5000_simple", "random_float", dict(n=1000, m=5000, seed=0), False, "numeric"))
    return cases


def gen_seeded(S, skip_slow):
    out = {}
    for name, gen, kw, integer, what in seeded_cases(skip_slow):
        t0 = time.time()
        n, U, V, C, T = synthetic.make(gen, **kw)
        edges = py_edges(U, V, C, T, integer)
        rec = {"generator": gen, "kwargs": kw, "integer": integer, "n": n, "m": len(edges)}
        if what in ("numeric", "both"):
            rec["solve"] = run_solve(S, n, edges)
        if what in ("exact", "both"):
            rec["exact"] = run_exact(S, n, edges)
        out[name] = rec
        print(f"  {name}: {time.time() - t0:.1f}s", flush=True)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-slow", action="store_true")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    S = import_reference()
    todo = args.only.split(",") if args.only else ["kat", "small", "seeded"]
    if "kat" in todo:
        json.dump(gen_kat(S), open(os.path.join(HERE, "kat.json"), "w"), indent=0)
        print("kat.json written", flush=True)
    if "small" in todo:
        json.dump(gen_small(S), open(os.path.join(HERE, "random_small.json"), "w"), indent=0)
        print("random_small.json written", flush=True)
    if "seeded" in todo:
        json.dump(gen_seeded(S, args.skip_slow), open(os.path.join(HERE, "seeded.json"), "w"), indent=0)
        print("seeded.json written", flush=True)


if __name__ == "__main__":
    main()
