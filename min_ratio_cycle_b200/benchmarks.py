"""
DIMACS loader and the benchmark / regression helpers that sit directly on top of the solve path
(reference: min_ratio_cycle/benchmarks.py:19-234), wired to the GPU-backed solver so the
reference's own CI-style checks (tests/test_regression.py, tests/test_analytics.py:50-78) run unchanged.
Same function names, argument meaning and return shapes as the reference.
"""

from __future__ import annotations

import json
import math
from collections.abc import Iterable, Sequence
from pathlib import Path
from statistics import mean, variance
from time import perf_counter

from .solver import MinRatioCycleSolver, SolverConfig


def load_dimacs_graph(lines: Iterable[str], config: SolverConfig | None = None) -> MinRatioCycleSolver:
    """
    Parse DIMACS ``.sp`` lines: ``p [sp] <n> <m>`` and ``a <u> <v> <cost> <time>`` (1-based vertices, weights
    read as floats), comment lines start with ``c`` (benchmarks.py:19-66).
    """
    n_vertices, edges = 0, []
    for line in lines:
        if not line or line.startswith("c"):
            continue
        tok = line.split()
        if not tok:
            continue
        if tok[0] == "p":
            n_vertices = int(tok[1] if len(tok) == 3 else tok[2])
        elif tok[0] == "a":
            edges.append((int(tok[1]) - 1, int(tok[2]) - 1, float(tok[3]), float(tok[4])))
    solver = MinRatioCycleSolver(n_vertices, config)
    solver.add_edges(edges)
    return solver


def compare_with_networkx(solver: MinRatioCycleSolver) -> float:
    """|solver ratio - brute-force optimum over nx.simple_cycles| (benchmarks.py:69-108).  Like the reference,
    parallel edges collapse to the last one added in the DiGraph; meant for tiny graphs only."""
    import networkx as nx

    G = nx.DiGraph()
    for e in solver._edges:
        G.add_edge(e.u, e.v, cost=float(e.cost), time=float(e.time))
    best = math.inf
    for cyc in nx.simple_cycles(G):
        hops = list(zip(cyc, cyc[1:] + [cyc[0]]))
        best = min(best, sum(G[u][v]["cost"] for u, v in hops) / sum(G[u][v]["time"] for u, v in hops))
    return abs(solver.solve().ratio - best)


def benchmark_solver(solver: MinRatioCycleSolver, *, compare: bool = True) -> dict[str, float | None]:
    """{"ratio", "time" (seconds of one solve()), "diff" (vs brute force, or None)} (benchmarks.py:111-140)."""
    t0 = perf_counter()
    result = solver.solve()
    elapsed = perf_counter() - t0
    return {"ratio": result.ratio, "time": elapsed, "diff": compare_with_networkx(solver) if compare else None}


def load_baseline(path: str | Path) -> dict[str, object]:
    with Path(path).open() as fh:
        return json.load(fh)


def save_baseline(path: str | Path, data: dict[str, object]) -> None:
    p = Path(path)
    p.parent.mkdir(parents=True, exist_ok=True)
    with p.open("w") as fh:
        json.dump(data, fh, indent=2)


def _welch_t(a: Sequence[float], b: Sequence[float]) -> float:
    """Welch t statistic (the reference delegates to analytics.compare_solutions, analytics.py:99-128)."""
    if len(a) < 2 or len(b) < 2:
        raise ValueError("need at least two samples per group")
    va, vb = variance(a) / len(a), variance(b) / len(b)
    if va + vb == 0:
        raise ValueError("zero variance")
    return (mean(a) - mean(b)) / math.sqrt(va + vb)


def regression_check(times: Sequence[float], ratio: float, baseline: dict[str, object], *, ratio_tol: float = 1e-9,
                     t_threshold: float = 2.0) -> dict[str, bool]:
    """{"performance": regressed?, "quality": regressed?} against a stored baseline (benchmarks.py:189-234)."""
    perf = False
    if baseline.get("times"):
        try:
            perf = _welch_t(times, baseline["times"]) > t_threshold
        except ValueError:
            perf = mean(times) > mean(baseline["times"]) * 1.5
    qual = baseline.get("ratio") is not None and abs(ratio - float(baseline["ratio"])) > ratio_tol
    return {"performance": perf, "quality": bool(qual)}
