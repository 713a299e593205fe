"""
Host-side mirror of the reference's solver interface (min_ratio_cycle/solver.py) with the parametric
negative-cycle hot path running on a B200 through libmrc_b200.so (include/mrc_b200.h).

What is kept identical to the reference (so this class is a drop-in for that path):
  * types and defaults: SolverMode/LogLevel/SolverConfig/Edge/SolverMetrics/SolverResult (:46-172)
  * MinRatioCycleSolver(n_vertices, config) with add_edge/add_edges/remove_edge/clear_edges (:193-333),
    solve(mode, target_ratio, **kwargs) incl. its recovery ladder and error types (:581-679),
    _solve_exact / _stern_brocot_search return conventions (:685-797), sensitivity_analysis (:1699-1729)
  * private names other code reads: n, config, _edges, _all_int, _arrays_built, _U/_V/_C/_T/_Ci/_Ti
    (destination-sorted, stable), _starts/_counts, _last_result, _build_numpy_arrays_once
What replaces reference code: every `_detect_negative_cycle_numeric`, `_has_negative_cycle_exact`,
`_compute_potentials_exact`, tight-mask and lambda-search call goes to the GPU library.  There is no
CPU fallback: without the library or a CUDA device these methods raise.

Additive (not in the reference): SolverConfig.gpu_* fields, add_edges_arrays() bulk ingestion,
solve_exact_fraction().
"""

from __future__ import annotations

import logging
import math
import time as _time
from dataclasses import dataclass, field
from enum import Enum
from fractions import Fraction
from typing import Any

import numpy as np

from . import _native as N
from .exceptions import ConvergenceError, NumericalInstabilityError, ResourceExhaustionError, TimeoutError

__version__ = "0.1.0"

_DEVICE_BUILD_MIN_EDGES = 200_000   # above this the stable sort by destination runs on the GPU


class SolverMode(Enum):
    AUTO = "auto"
    EXACT = "exact"
    NUMERIC = "numeric"
    APPROXIMATE = "approximate"


class LogLevel(Enum):
    NONE = 0
    ERROR = 1
    WARN = 2
    INFO = 3
    DEBUG = 4


@dataclass
class SolverConfig:
    # reference fields and defaults (solver.py:69-100)
    numeric_max_iter: int = 60
    numeric_tolerance: float = 1e-12
    numeric_cycle_slack: float = 1e-15
    exact_max_denominator: int | None = None
    exact_max_steps: int | None = None
    sparse_threshold: float = 0.1
    enable_preprocessing: bool = True
    enable_early_termination: bool = True
    validate_cycles: bool = True
    repair_cycles: bool = True
    log_level: LogLevel = LogLevel.INFO
    collect_metrics: bool = True
    use_kahan_summation: bool = True   # accepted for compatibility; the device update has no Kahan term
    max_solve_time: float | None = None
    max_memory_mb: int | None = None
    # additive GPU knobs
    gpu_device: int = 0
    gpu_devices: list | None = None    # several GPUs of one box (e.g. [0, 1, 2, 3]): the edge arrays are replicated over
                                       # NVLink and the lambda candidates of the numeric search are sharded over the
                                       # devices inside the library (mrc_comm_init / mrc_mgraph_*); None = gpu_device only
    gpu_candidates: int = 1            # lambdas evaluated per relaxation launch (1..8).  1 = Newton steps on the best cycle's
                                       # ratio + one certifying probe: on ONE GPU that is 1.5-2.8x faster than 4-way
                                       # multisection on most graphs (a negative probe costs 2-6 sweeps, a non-negative
                                       # one next to lambda* 20-100, profiles/r01e_policy.txt section 11); K > 1 is the
                                       # throughput / multi-GPU configuration (bench.py, sharded.py)
    gpu_scenarios_per_sweep: int = 8   # sensitivity_analysis: single-edge scenarios that share one edge sweep (1..8)
    gpu_exact_strategy: int = 1        # 0: every Stern-Brocot step probes the GPU; 1: Newton + replay
    gpu_sparse: bool = True            # late passes walk only the out-edges of the vertices still being lowered
                                       # (worklist + out-edge index; same fixed point and verdicts, shorter solves)
    honor_mode: bool = False           # True: solve(mode="exact") really runs _solve_exact (reference ignores mode)
    gpu_reference_convergence: bool = True   # raise ConvergenceError where max_iter halvings of the initial bracket cannot
                                             # reach the tolerance, as the reference's bisection does (:1102-1109)


@dataclass(frozen=True)
class Edge:
    u: int
    v: int
    cost: float | int | Fraction
    time: float | int | Fraction

    def __post_init__(self):
        if not isinstance(self.u, int) or not isinstance(self.v, int):
            raise ValueError("Vertex indices must be integers")
        if self.u < 0 or self.v < 0:
            raise ValueError("Vertex indices must be non-negative")
        if self.time <= 0:
            raise ValueError("Edge transit time must be strictly positive")


@dataclass
class SolverMetrics:
    solve_time: float = 0.0
    iterations: int = 0
    mode_used: str = ""
    preprocessing_time: float = 0.0
    graph_properties: dict[str, Any] = field(default_factory=dict)
    memory_peak: int | None = None


@dataclass
class SolverResult:
    cycle: list[int]
    sum_cost: float | Fraction
    sum_time: float | Fraction
    ratio: float | Fraction
    success: bool = True
    error_message: str = ""
    metrics: SolverMetrics | None = None
    config_used: SolverConfig | None = None

    def __iter__(self):
        return iter((self.cycle, self.sum_cost, self.sum_time, self.ratio))


def dominated_mask(n: int, U, V, Cw, Tw):
    """
    Boolean mask of edges the reference's _remove_dominated_edges (solver.py:531-575) would delete: another
    edge with the same (u,v) has cost <= and time <= with one strict.  None when nothing is dominated.
    Vectorised: only (u,v) groups with more than one member are examined.
    """
    m = len(U)
    if m < 2:
        return None
    key = U * n + V
    order = np.argsort(key, kind="stable")
    ks = key[order]
    dup = np.zeros(m, dtype=bool)
    same = ks[1:] == ks[:-1]
    dup[1:] |= same
    dup[:-1] |= same
    if not dup.any():
        return None
    drop = np.zeros(m, dtype=bool)
    idx = order[dup]
    kd = ks[dup]
    bounds = np.flatnonzero(np.concatenate([[True], kd[1:] != kd[:-1], [True]]))
    for s, e in zip(bounds[:-1], bounds[1:]):
        grp = idx[s:e]
        c, t = Cw[grp], Tw[grp]
        le = (c[:, None] <= c[None, :]) & (t[:, None] <= t[None, :])
        lt = (c[:, None] < c[None, :]) | (t[:, None] < t[None, :])
        drop[grp[(le & lt).any(axis=0)]] = True
    return drop if drop.any() else None


def _is_integral(x) -> bool:
    """`_all_int` rule of the reference (solver.py:289-298): int, Fraction, or an integral float."""
    return isinstance(x, (int, Fraction)) or (isinstance(x, float) and x.is_integer())


def _mirror(name):
    """Host mirror of a device-resident array (the reference's private `_U`, `_V`, ... that tools read): after the
    device-side build (mrc_graph_build) it is fetched from the device on first read instead of on every build."""
    slot = "_m" + name

    def get(self):
        if getattr(self, slot, None) is None and getattr(self, "_mirrors_pending", False):
            self._materialize_mirrors()
        return getattr(self, slot, None)

    def put(self, value):
        setattr(self, slot, value)
    return property(get, put)


class MinRatioCycleSolver:
    """Minimum cost-to-time ratio cycle; lambda search and Bellman-Ford run on the GPU."""

    _U, _V, _C, _T = _mirror("_U"), _mirror("_V"), _mirror("_C"), _mirror("_T")
    _starts, _counts, _order = _mirror("_starts"), _mirror("_counts"), _mirror("_order")

    def __init__(self, n_vertices: int, config: SolverConfig | None = None):
        if not isinstance(n_vertices, int) or n_vertices <= 0:
            raise ValueError("n_vertices must be a positive integer")
        self.n = n_vertices
        self.config = config or SolverConfig()
        self._edges: list[Edge] = []
        self._bulk: list[tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]] = []
        self._bulk_int: list[tuple[np.ndarray, np.ndarray] | None] = []   # exact int64 weights of a bulk block given as integers
        self._presorted = None
        self._no_repeated_pairs = False
        self._hop_cache: dict = {}
        self._setup_logging()
        self._arrays_built = False
        self._mirrors_pending = False
        self._U = self._V = self._C = self._T = self._Ci = self._Ti = None
        self._starts = self._counts = self._order = None
        self._all_int = True
        self._is_sparse = False
        self._last_result: SolverResult | None = None
        self._metrics = SolverMetrics() if self.config.collect_metrics else None
        self._dev: N.DeviceGraph | None = None
        self._mg = None                    # N.MultiGraph when config.gpu_devices names several GPUs
        self._comm = None
        self._last_exact: dict | None = None
        self.logger.info("Initialized solver with %d vertices", n_vertices)

    # ------------------------------------------------------------------ plumbing
    def _setup_logging(self) -> None:
        self.logger = logging.getLogger(f"{__name__}.{id(self)}")
        level = {LogLevel.NONE: logging.CRITICAL + 1, LogLevel.ERROR: logging.ERROR, LogLevel.WARN: logging.WARNING,
                 LogLevel.INFO: logging.INFO, LogLevel.DEBUG: logging.DEBUG}[self.config.log_level]
        self.logger.setLevel(level)
        if not self.logger.handlers and self.config.log_level != LogLevel.NONE:
            h = logging.StreamHandler()
            h.setFormatter(logging.Formatter("%(asctime)s - %(name)s - %(levelname)s - %(message)s"))
            self.logger.addHandler(h)

    def _build_numpy_arrays_once(self) -> None:
        self._build_arrays_if_needed()

    def _invalidate(self) -> None:
        self._arrays_built = False
        self._presorted = None
        self._no_repeated_pairs = False
        self._mirrors_pending = False

    def _materialize_mirrors(self) -> None:
        """_U/_V/_C/_T (int64 / f64, destination-sorted), _order, _starts, _counts from the device-resident arrays."""
        self._mirrors_pending = False
        so, do, co, to, order = self._dev.fetch_arrays(order=True)
        self._U, self._V, self._C, self._T = so.astype(np.int64), do.astype(np.int64), co, to
        self._order = order.astype(np.int64)
        self._counts = np.bincount(self._V, minlength=self.n)
        self._starts = np.zeros(self.n, dtype=np.int64)
        if self.n > 1:
            np.cumsum(self._counts[:-1], out=self._starts[1:])
        self._hop_cache = {}

    def _device_build(self, remove_dominated: bool) -> bool:
        """
        Large numeric graph given as ONE bulk block: _preprocess_graph + _build_arrays_if_needed + upload as one library
        call (mrc_graph_build) -- dominated-edge removal, stable sort by destination and CSR on the device, the sorted
        arrays stay there; the host mirrors are fetched only if somebody reads them.  Returns False when not applicable.
        """
        if (self._arrays_built or self._edges or len(self._bulk) != 1 or self._all_int
                or len(self._bulk[0][0]) < _DEVICE_BUILD_MIN_EDGES or len(self.config.gpu_devices or []) > 1):
            return False
        t0 = _time.time()
        U, V, Cw, Tw = self._bulk[0]
        if self._dev is not None:
            self._dev.close()
        self._close_multi()
        try:
            self._dev, dups, removed = N.DeviceGraph.build(self.n, U, V, Cw, Tw, remove_dominated=remove_dominated,
                                                           device=self.config.gpu_device)
        except N.NativeError as exc:
            if exc.code == N.ERR_ARG and "endpoint" in exc.message:
                raise ValueError("u and v must be valid vertex indices") from exc
            if exc.code == N.ERR_ARG and "time" in exc.message:
                raise ValueError("Edge transit time must be strictly positive") from exc
            raise
        if removed:
            keep = self._dev.fetch_keep()                      # the reference drops the edges from its list for good
            self._bulk = [tuple(a[keep] for a in self._bulk[0])]
            self._bulk_int = [None]
            self.logger.info("Preprocessing removed %d dominated edges", removed)
        self._no_repeated_pairs = dups == 0 or remove_dominated
        for name in ("_U", "_V", "_C", "_T", "_starts", "_counts", "_order"):
            setattr(self, name, None)
        self._Ci = self._Ti = None
        self._hop_cache = {}
        self._mirrors_pending = True
        self._is_sparse = self._dev.m / (self.n * self.n) < self.config.sparse_threshold
        self._dev.set_option("sparse", 1 if self.config.gpu_sparse else 0)
        self._arrays_built = True
        if self._metrics:
            self._metrics.preprocessing_time = _time.time() - t0
        return True

    def _num_edges(self) -> int:
        return len(self._edges) + sum(len(b[0]) for b in self._bulk)

    def _inv_order(self) -> np.ndarray:
        """insertion index -> position in the destination-sorted arrays (the inverse of _order), kept until _order changes:
        at m = 2 M the scatter costs 30 ms, more than 1,024 perturbation re-solves on the device."""
        order = self._order
        cached = getattr(self, "_inv_order_cache", None)
        if cached is None or cached[0] is not order:
            inv = np.empty(len(order), dtype=np.int64)
            inv[order] = np.arange(len(order))
            self._inv_order_cache = cached = (order, inv)
        return cached[1]

    def _close_multi(self) -> None:
        if self._mg is not None:
            self._mg.close()
            self._mg = None
            self._dev = None               # was a borrowed replica of the multi-graph
        if self._comm is not None:
            self._comm.close()
            self._comm = None

    def __del__(self):
        try:
            if self._mg is not None:
                self._close_multi()
            elif self._dev is not None:
                self._dev.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ edge management (:267-333)
    def add_edge(self, u: int, v: int, cost, time) -> None:
        if not (0 <= u < self.n and 0 <= v < self.n):
            raise ValueError("u and v must be valid vertex indices")
        edge = Edge(u, v, cost, time)
        if self._all_int and not (_is_integral(cost) and _is_integral(time)):
            self._all_int = False
        self._edges.append(edge)
        self._invalidate()

    def add_edges(self, edges) -> None:
        for u, v, cost, time in edges:
            self.add_edge(u, v, cost, time)

    def add_edges_arrays(self, U, V, cost, time) -> None:
        """
        Bulk ingestion (additive; the reference only has the per-edge Python path, :267-310, which costs
        minutes and GBs of Edge objects at 10^7 edges).  Edges added here come after all add_edge() edges
        in insertion order.  Weights count as integers when every value is integral.
        """
        U = np.ascontiguousarray(U, dtype=np.int64)
        V = np.ascontiguousarray(V, dtype=np.int64)
        ints = None
        if np.asarray(cost).dtype.kind in "iu" and np.asarray(time).dtype.kind in "iu":
            # integer weights stay exact (the float64 view below rounds above 2^53; exact mode reads these, :493-498)
            ints = (np.array(cost, dtype=np.int64), np.array(time, dtype=np.int64))
        Cw = np.ascontiguousarray(cost, dtype=np.float64)
        Tw = np.ascontiguousarray(time, dtype=np.float64)
        if not (len(U) == len(V) == len(Cw) == len(Tw)):
            raise ValueError("edge arrays must have equal length")
        if len(U) == 0:
            return
        large = len(U) >= _DEVICE_BUILD_MIN_EDGES
        if self._all_int:
            head = slice(0, 4096)          # a float graph shows in the first few values: no full pass needed
            if not (np.all(Cw[head] == np.trunc(Cw[head])) and np.all(Tw[head] == np.trunc(Tw[head]))):
                self._all_int = False
            elif not (np.all(Cw == np.trunc(Cw)) and np.all(Tw == np.trunc(Tw))):
                self._all_int = False
        if not large or self._all_int:
            if U.min() < 0 or V.min() < 0 or U.max() >= self.n or V.max() >= self.n:
                raise ValueError("u and v must be valid vertex indices")
            if not np.all(Tw > 0):
                raise ValueError("Edge transit time must be strictly positive")
            U, V, Cw, Tw = U.copy(), V.copy(), Cw.copy(), Tw.copy()   # later changes to the caller's arrays must not reach the graph
        # else: a large numeric block is range-checked on the device when the arrays are built (same ValueErrors, raised
        # by solve()); it is borrowed, not copied -- 320 MB at 10^7 edges -- so the caller must not modify it until then
        self._bulk.append((U, V, Cw, Tw))
        self._bulk_int.append(ints)
        self._invalidate()

    def remove_edge(self, u: int, v: int) -> bool:
        """First edge u -> v in insertion order (:312-325), whether it came through add_edge or add_edges_arrays."""
        for i, e in enumerate(self._edges):
            if e.u == u and e.v == v:
                del self._edges[i]
                self._invalidate()
                return True
        for bi, (U, V, Cw, Tw) in enumerate(self._bulk):
            hit = np.flatnonzero((U == u) & (V == v))
            if hit.size:
                k = np.ones(len(U), dtype=bool)
                k[hit[0]] = False
                self._bulk[bi] = (U[k], V[k], Cw[k], Tw[k])
                if self._bulk_int[bi] is not None:
                    self._bulk_int[bi] = tuple(a[k] for a in self._bulk_int[bi])
                if len(self._bulk[bi][0]) == 0:
                    del self._bulk[bi]
                    del self._bulk_int[bi]
                self._invalidate()
                return True
        return False

    def clear_edges(self) -> None:
        self._edges.clear()
        self._bulk.clear()
        self._bulk_int.clear()
        self._invalidate()

    # ------------------------------------------------------------------ graph analysis (:339-444)
    def get_graph_properties(self) -> dict[str, Any]:
        if self._num_edges() == 0:
            return {"error": "No edges in graph"}
        self._build_arrays_if_needed()
        n, m = self.n, self._dev.m
        density = m / (n * n)
        # the degree / weight statistics come from the device-resident arrays (mrc_graph_properties): solve() stores
        # this dictionary in result.metrics on every call and the eight NumPy sweeps cost 160 ms at m = 10^7
        p = self._device().properties()
        return {
            "vertices": n, "edges": m, "density": density, "is_sparse": density < self.config.sparse_threshold,
            "connectivity": {
                "isolated_vertices": int(p["isolated_vertices"]),
                "source_vertices": int(p["source_vertices"]),
                "sink_vertices": int(p["sink_vertices"]),
                "avg_in_degree": m / n, "avg_out_degree": m / n,
                "max_in_degree": int(p["max_in_degree"]), "max_out_degree": int(p["max_out_degree"]),
            },
            "weights": {
                "all_integer": self._all_int,
                "cost_range": (p["cost_min"], p["cost_max"]),
                "time_range": (p["time_min"], p["time_max"]),
                "cost_mean": p["cost_mean"], "time_mean": p["time_mean"],
                "cost_std": p["cost_std"], "time_std": p["time_std"],
            },
        }

    def detect_issues(self) -> list[dict[str, str]]:
        props = self.get_graph_properties()
        if "error" in props:
            return [{"level": "ERROR", "message": props["error"]}]
        issues = []
        iso = props["connectivity"]["isolated_vertices"]
        if iso > 0:
            issues.append({"level": "WARNING", "message": f"{iso} isolated vertices (no cycles possible through them)"})
        if props["density"] < 0.01:
            issues.append({"level": "INFO", "message": f"Very sparse graph (density: {props['density']:.4f})"})
        w = props["weights"]
        if w["cost_range"][1] - w["cost_range"][0] > 1e10:
            issues.append({"level": "WARNING", "message": "Very large cost range may cause numerical issues"})
        if w["time_range"][0] > 0 and w["time_range"][1] / w["time_range"][0] > 1e6:
            issues.append({"level": "WARNING", "message": "Very large time ratio may cause numerical issues"})
        return issues

    # ------------------------------------------------------------------ arrays + preprocessing (:450-575)
    def _insertion_arrays(self):
        m1 = len(self._edges)
        if m1 == 0 and len(self._bulk) == 1:      # one bulk block and nothing else: no copies
            return self._bulk[0]
        U = np.fromiter((e.u for e in self._edges), dtype=np.int64, count=m1)
        V = np.fromiter((e.v for e in self._edges), dtype=np.int64, count=m1)
        Cw = np.fromiter((float(e.cost) for e in self._edges), dtype=np.float64, count=m1)
        Tw = np.fromiter((float(e.time) for e in self._edges), dtype=np.float64, count=m1)
        if self._bulk:
            U = np.concatenate([U] + [b[0] for b in self._bulk])
            V = np.concatenate([V] + [b[1] for b in self._bulk])
            Cw = np.concatenate([Cw] + [b[2] for b in self._bulk])
            Tw = np.concatenate([Tw] + [b[3] for b in self._bulk])
        return U, V, Cw, Tw

    def _build_arrays_if_needed(self) -> None:
        """Destination-sorted COO + CSR-by-destination, then upload (replaces :450-511)."""
        if self._arrays_built:
            return
        t0 = _time.time()
        if self._num_edges() == 0:
            self.logger.warning("No edges to build arrays from")
            return
        if self._presorted is None and self._device_build(remove_dominated=False):
            return
        if self._presorted is not None:
            so, do, co, to, order = self._presorted            # device sort already done by _preprocess_graph
            self._presorted = None
        elif self._num_edges() >= _DEVICE_BUILD_MIN_EDGES:
            U, V, Cw, Tw = self._insertion_arrays()
            so, do, co, to, order, _ = N.sort_edges_by_dst(self.n, U, V, Cw, Tw, device=self.config.gpu_device,
                                                           count_dups=False)
        else:
            so = None
        u32 = v32 = None
        if so is not None:                                     # stable radix sort on the GPU (mrc_sort_edges_by_dst)
            self._U, self._V, self._C, self._T = so.astype(np.int64), do.astype(np.int64), co, to
            order = order.astype(np.int64)
            u32, v32 = so, do                                  # the upload below takes the int32 arrays as they are
        else:
            U, V, Cw, Tw = self._insertion_arrays()
            order = np.argsort(V, kind="stable")
            self._U, self._V, self._C, self._T = U[order], V[order], Cw[order], Tw[order]
        self._order = order
        self._hop_cache = {}
        self._counts = np.bincount(self._V, minlength=self.n)
        self._starts = np.zeros(self.n, dtype=np.int64)
        if self.n > 1:
            np.cumsum(self._counts[:-1], out=self._starts[1:])
        if self._all_int:
            # int(edge.cost), int(edge.time) (:493-498) from the ORIGINAL values: a Python int or an integer array above
            # 2^53 is not representable in the float64 arrays; above 2^63 this raises OverflowError like the reference
            ci = [np.array([int(e.cost) for e in self._edges], dtype=np.int64)]
            ti = [np.array([int(e.time) for e in self._edges], dtype=np.int64)]
            for b, bi in zip(self._bulk, self._bulk_int):
                ci.append(np.trunc(b[2]).astype(np.int64) if bi is None else bi[0])
                ti.append(np.trunc(b[3]).astype(np.int64) if bi is None else bi[1])
            self._Ci = np.concatenate(ci)[order]
            self._Ti = np.concatenate(ti)[order]
        else:
            self._Ci = self._Ti = None
        self._is_sparse = len(self._U) / (self.n * self.n) < self.config.sparse_threshold
        if self._dev is not None:
            self._dev.close()
        self._close_multi()
        devs = [int(d) for d in (self.config.gpu_devices or [])]
        if len(devs) > 1 and not self._all_int:
            # numeric graph on several GPUs: replicas of the sorted arrays on every device (one H2D + NCCL broadcast);
            # the replica on devs[0] doubles as the single-device handle for bounds, properties and probes
            self._comm = N.Comm.group(devs)
            self._mg = N.MultiGraph(self._comm, self.n, self._U if u32 is None else u32, self._V if v32 is None else v32,
                                    self._C, self._T)
            self._dev = self._mg.replica(0)
            self._dev.device = devs[0]
        else:
            self._dev = N.DeviceGraph(self.n, self._U if u32 is None else u32, self._V if v32 is None else v32,
                                      self._C, self._T, self._Ci, self._Ti,
                                      device=devs[0] if devs else self.config.gpu_device)
        for g in ([self._mg.replica(i) for i in range(len(devs))] if self._mg is not None else [self._dev]):
            g.set_option("sparse", 1 if self.config.gpu_sparse else 0)
        self._arrays_built = True
        if self._metrics:
            self._metrics.preprocessing_time = _time.time() - t0

    def _device(self) -> N.DeviceGraph:
        self._build_arrays_if_needed()
        if self._dev is None:
            raise RuntimeError("Edge arrays must be initialized before solving")
        return self._dev

    def _preprocess_graph(self) -> None:
        if not self.config.enable_preprocessing:
            return
        if self._no_repeated_pairs:          # census already done for this edge set: nothing can be dominated
            return
        if self._device_build(remove_dominated=True):
            return
        before = self._num_edges()
        if before >= _DEVICE_BUILD_MIN_EDGES and not self._arrays_built:
            # Large graph: one device pass sorts by destination AND counts repeated (u,v) pairs.  Without repeats
            # nothing can be dominated (:531-575 only compares edges of the same pair) and the sorted arrays are
            # kept for _build_arrays_if_needed; with repeats the host path below runs as usual.
            U, V, Cw, Tw = self._insertion_arrays()
            so, do, co, to, order, dups = N.sort_edges_by_dst(self.n, U, V, Cw, Tw, device=self.config.gpu_device)
            if dups == 0:
                self._presorted = (so, do, co, to, order)
                self._no_repeated_pairs = True
                return
        self._remove_dominated_edges()
        if self._num_edges() != before:
            self.logger.info("Preprocessing removed %d dominated edges", before - self._num_edges())
            self._invalidate()

    def _remove_dominated_edges(self) -> None:
        """
        Drop an edge when another edge with the same (u,v) has cost <= and time <= with one strict
        (:531-575).  Vectorised: only (u,v) groups with more than one member are examined.
        """
        m = self._num_edges()
        if m < 2:
            return
        U, V, Cw, Tw = self._insertion_arrays()
        drop = dominated_mask(self.n, U, V, Cw, Tw)
        if drop is None:
            return
        m1 = len(self._edges)
        self._edges = [e for i, e in enumerate(self._edges) if not drop[i]]
        off = m1
        kept, kept_int = [], []
        for b, bi in zip(self._bulk, self._bulk_int):
            k = ~drop[off: off + len(b[0])]
            off += len(b[0])
            if k.any():
                kept.append(tuple(a[k] for a in b))
                kept_int.append(None if bi is None else tuple(a[k] for a in bi))
        self._bulk, self._bulk_int = kept, kept_int

    # ------------------------------------------------------------------ solve (:581-679)
    def solve(self, mode: SolverMode | str = SolverMode.AUTO, target_ratio: float | None = None, **kwargs) -> SolverResult:
        """solve() of the reference (:581-679): same ladder, errors and result conventions.  `mode` is validated and then
        ignored unless SolverConfig.honor_mode (the reference always runs numeric, :622-627).  `target_ratio` is accepted
        and ignored: the reference's early exit compares it with the cycle's sum_time (a bug, :1063-1071) and a search
        that needs 5-10 probes has nothing to cut short."""
        if isinstance(mode, str):
            mode = SolverMode(mode)
        t_start = _time.time()
        if self.config.max_memory_mb is not None:
            import psutil
            mem = psutil.Process().memory_info().rss / (1024 * 1024)
            if mem > self.config.max_memory_mb:
                raise ResourceExhaustionError("Memory limit exceeded before solve", resource="memory",
                                              limit=self.config.max_memory_mb, usage=mem,
                                              suggested_fix="increase max_memory_mb")
        self._preprocess_graph()
        self._build_arrays_if_needed()
        if self._num_edges() == 0:
            raise ValueError("Graph has no edges")
        run_exact = self.config.honor_mode and (mode == SolverMode.EXACT or (mode == SolverMode.AUTO and self._all_int))
        actual = SolverMode.EXACT if run_exact else SolverMode.NUMERIC   # the reference always runs numeric (:622-623)
        self.logger.info("Starting solve in %s mode", actual.value)
        if run_exact:
            result = self._solve_exact(**kwargs)
        else:
            try:
                result = self._solve_numeric(target_ratio=target_ratio, **kwargs)
            except (NumericalInstabilityError, ConvergenceError) as exc:
                self.logger.warning("Numeric solve failed: %s", exc)
                if self._all_int:
                    result = self._solve_exact(**kwargs)          # :630-632
                else:
                    old = self.config.numeric_tolerance           # :634-638
                    self.config.numeric_tolerance *= 10
                    try:
                        result = self._solve_numeric(target_ratio=target_ratio, **kwargs)
                    finally:
                        self.config.numeric_tolerance = old
                if not result.success:
                    raise NumericalInstabilityError("Solver failed in all recovery attempts", component="solve",
                                                    suggested_fix="check input weights",
                                                    recovery_hint="try exact mode or scale weights") from exc
        if self.config.validate_cycles and result.success:
            result = self._validate_result(result)
        if self._metrics:
            self._metrics.solve_time = _time.time() - t_start
            self._metrics.mode_used = actual.value
            self._metrics.graph_properties = self.get_graph_properties()
            result.metrics = self._metrics
        result.config_used = self.config
        self._last_result = result
        if not result.success:
            self.logger.error("Solve failed: %s", result.error_message)
            raise RuntimeError(result.error_message)
        return result

    # ------------------------------------------------------------------ exact mode (:685-995)
    def _solve_exact(self, **kwargs) -> SolverResult:
        fail = dict(cycle=[], sum_cost=0, sum_time=0, ratio=float("inf"), success=False)
        if not self._all_int:
            return SolverResult(error_message="Exact mode requires integer weights", **fail)
        if self._Ci is None or self._Ti is None:
            raise RuntimeError("Integer weight arrays (_Ci, _Ti) must be initialized before solving")
        max_den = kwargs.get("max_denominator", self.config.exact_max_denominator)
        max_steps = kwargs.get("max_steps", self.config.exact_max_steps)
        try:
            cycle, sum_c, sum_t, _ab, _r = self._stern_brocot_search(max_den, max_steps)
            if cycle and (len(cycle) == 1 or cycle[0] != cycle[-1]):
                cycle.append(cycle[0])
            return SolverResult(cycle=cycle, sum_cost=float(sum_c), sum_time=float(sum_t),
                                ratio=float(sum_c) / float(sum_t), success=True)
        except Exception as exc:  # the reference swallows everything here (:724-732)
            return SolverResult(error_message=f"Exact solver failed: {exc}", **fail)

    def solve_exact_fraction(self, **kwargs):
        """Additive: (closed cycle, Fraction(sum_cost, sum_time)) from the exact path, or raises RuntimeError."""
        self._preprocess_graph()
        self._build_arrays_if_needed()
        r = self._solve_exact(**kwargs)
        if not r.success:
            raise RuntimeError(r.error_message)
        return r.cycle, Fraction(int(r.sum_cost), int(r.sum_time))

    def _stern_brocot_search(self, max_den, max_steps):
        dev = self._device()
        out = dev.solve_exact(max_den, max_steps, strategy=self.config.gpu_exact_strategy)
        self._last_exact = out
        if out["rc"] != N.OK:
            raise RuntimeError(out["message"])
        if self._metrics and out["iterations"] >= 0:
            self._metrics.iterations = int(out["iterations"])
        sc, st = int(out["sum_cost"]), int(out["sum_time"])
        return list(out["cycle"]), sc, st, (int(out["a"]), int(out["b"])), sc / st

    def _has_negative_cycle_exact(self, a: int, b: int) -> bool:
        r = self._device().probe_exact([a], [b])[0]
        return r["verdict"] in (N.V_NEG, N.V_NEG_NOCYCLE)

    def _compute_potentials_exact(self, a: int, b: int):
        return self._device().potentials_exact(a, b)

    def _find_zero_cycle_exact(self, a: int, b: int):
        z = self._device().zero_cycle_exact(a, b)
        return None if z is None else (list(z[0]), int(z[1]), int(z[2]))

    def _compute_cycle_weights_exact(self, cycle):
        """Per hop the first-inserted (u,v) edge (:971-995); lookup through the CSR instead of a scan of _edges."""
        self._build_arrays_if_needed()
        sc = st = 0
        L = len(cycle)
        for i in range(L):
            u, v = cycle[i], cycle[(i + 1) % L]
            s = int(self._starts[v])
            hit = np.flatnonzero(self._U[s: s + int(self._counts[v])] == u)
            if hit.size == 0:
                raise RuntimeError(f"Edge {u} -> {v} not found in cycle extraction")
            sc += int(self._Ci[s + hit[0]])
            st += int(self._Ti[s + hit[0]])
        return cycle, sc, st

    # ------------------------------------------------------------------ numeric mode (:1001-1337)
    def _solve_numeric(self, target_ratio: float | None = None, **kwargs) -> SolverResult:
        lambda_lo, lambda_hi = kwargs.get("lambda_lo"), kwargs.get("lambda_hi")
        max_iter = kwargs.get("max_iter", self.config.numeric_max_iter)
        tol = kwargs.get("tolerance", self.config.numeric_tolerance)
        slack = kwargs.get("cycle_slack", self.config.numeric_cycle_slack)
        try:
            dev = self._device()
            if lambda_lo is None or lambda_hi is None:
                lo, hi = dev.bounds()                      # min/max of C/T -/+ 1 (:1019-1023)
                lambda_lo = lambda_lo or lo                # sic (:1024-1025)
                lambda_hi = lambda_hi or hi
            if self.config.max_memory_mb is not None:
                import psutil
                mem = psutil.Process().memory_info().rss / (1024 * 1024)
                if mem > self.config.max_memory_mb:
                    raise ResourceExhaustionError("Numeric solve exceeded memory limit", resource="memory",
                                                  limit=self.config.max_memory_mb, usage=mem)
            t0 = _time.time()
            limit = self.config.max_solve_time
            if self.config.gpu_reference_convergence and (lambda_hi - lambda_lo) / 2.0 ** max(int(max_iter), 0) > tol:
                # The reference halves [lo, hi] max_iter times and raises when the bracket is still wider than tol
                # (:1102-1109) -- with the defaults that is every graph whose cost/time ratios span more than
                # 1e-12 * 2^60 ~ 1.15e6 (tests/test_solver.py:140-153 pins it).  The Newton steps here would converge,
                # but a drop-in keeps the reference's error behaviour; gpu_reference_convergence=False lifts it.
                raise ConvergenceError("Numeric solver failed to converge", algorithm_name="Lawler",
                                       max_iterations=max_iter, tolerance=tol, final_error=lambda_hi - lambda_lo)
            warm = kwargs.get("_warm_cycle")
            warm_sums = None
            if warm:
                warm_sums = self._compute_cycle_weights_float(warm[:-1])
                if warm_sums[1] > 0 and lambda_lo < warm_sums[0] / warm_sums[1] <= lambda_hi:
                    lambda_hi = warm_sums[0] / warm_sums[1]
                else:
                    warm = None
            if self._mg is not None and not warm and not kwargs.get("_single_device"):
                out = self._mg.solve_numeric(lambda_lo, lambda_hi, max_iter, tol, slack, K_local=self.config.gpu_candidates,
                                             max_seconds=-1.0 if limit is None else max(float(limit), 1e-300))
            else:
                out = dev.solve_numeric(lambda_lo, lambda_hi, max_iter, tol, slack, K=self.config.gpu_candidates,
                                        max_seconds=-1.0 if limit is None else max(float(limit), 1e-300),
                                        hi_is_cycle=bool(warm))
            if warm and out["rc"] in (N.OK, N.ERR_CONVERGENCE) and not out["cycle"]:
                out["cycle"] = list(warm)          # no better cycle exists: the warm-start cycle stands
            if out["rc"] == N.ERR_TIMEOUT:
                raise TimeoutError("Numeric solve exceeded time limit", time_elapsed=_time.time() - t0,
                                   time_limit=limit, operation="numeric_solve")
            if out["rc"] == N.ERR_NO_CYCLE:
                raise NumericalInstabilityError("No negative cycle found", component="numeric_solve")
            cycle = [int(x) for x in out["cycle"]]
            sum_cost, sum_time = self._compute_cycle_weights_float(cycle[:-1])   # reference hop rule (:1289)
            if self._metrics:
                self._metrics.iterations = int(out["iterations"])
            if out["rc"] == N.ERR_CONVERGENCE:
                raise ConvergenceError("Numeric solver failed to converge", algorithm_name="Lawler",
                                       max_iterations=max_iter, tolerance=tol, final_error=out["hi"] - out["lo"])
            return SolverResult(cycle=cycle, sum_cost=sum_cost, sum_time=sum_time, ratio=sum_cost / sum_time,
                                success=True)
        except Exception as exc:   # every failure surfaces as NumericalInstabilityError (:1119-1122)
            raise NumericalInstabilityError("Numeric solver failed", component="numeric_solve") from exc

    def _detect_negative_cycle_numeric(self, lam: float, slack: float = 0.0):
        """(has_neg, (open cycle, sum_cost, sum_time, ratio) | None) like :1124-1294, via one GPU probe."""
        r = self._device().probe_numeric([lam], slack)[0]
        if r["verdict"] in (N.V_NONNEG, N.V_TIE):
            return False, None
        if r["verdict"] == N.V_NEG_NOCYCLE or not r["cycle"]:
            return True, None
        cycle = [int(x) for x in r["cycle"][:-1]]
        sc, st = self._compute_cycle_weights_float(cycle)
        if st <= 0:
            return True, None
        return True, (cycle, sc, st, sc / st)

    def _hop_edges(self, u: int, v: int):
        """Sorted-array indices of the edges u -> v (CSR segment of v), memoised per array build."""
        key = (u, v)
        hit = self._hop_cache.get(key)
        if hit is None:
            s = int(self._starts[v])
            hit = s + np.flatnonzero(self._U[s: s + int(self._counts[v])] == u)
            self._hop_cache[key] = hit
        return hit

    def _compute_cycle_weights_float(self, cycle):
        """
        Per hop, among parallel (u,v) edges the first with minimal cost/time (:1309-1313), summed in cycle
        order from 0.0.  Uses the current CSR arrays (the reference's dict is built once and goes stale,
        SURVEY 0.8 -- deliberately not reproduced).
        """
        self._build_arrays_if_needed()
        if self._mirrors_pending:          # mrc_cycle_weights: same hop rule and summation order on the resident arrays
            sc, st, missing, _ = self._dev.cycle_weights(list(cycle) + [cycle[0]])
            if missing >= 0:
                raise RuntimeError(f"Edge {cycle[missing]} -> {cycle[(missing + 1) % len(cycle)]} not found in cycle")
            return sc, st
        sc = st = 0.0
        L = len(cycle)
        for i in range(L):
            u, v = cycle[i], cycle[(i + 1) % L]
            hit = self._hop_edges(u, v)
            if hit.size == 0:
                raise RuntimeError(f"Edge {u} -> {v} not found in cycle")
            k = hit[int(np.argmin(self._C[hit] / self._T[hit]))] if hit.size > 1 else hit[0]
            sc += float(self._C[k])
            st += float(self._T[k])
        return sc, st

    # ------------------------------------------------------------------ validation (:1495-1582)
    def _verify_cycle_exists(self, cycle):
        if self._mirrors_pending:
            he = self._dev.cycle_weights(cycle)[3]
            issues = [f"Missing edge {u} -> {v}" for (u, v), e in zip(zip(cycle[:-1], cycle[1:]), he) if e < 0]
            return not issues, issues
        issues = [f"Missing edge {u} -> {v}" for u, v in zip(cycle[:-1], cycle[1:]) if self._hop_edges(u, v).size == 0]
        return not issues, issues

    def _validate_result(self, result: SolverResult) -> SolverResult:
        if not result.success or not result.cycle:
            return result
        try:
            if len(result.cycle) < 3:
                if not self.config.repair_cycles:
                    result.success, result.error_message = False, "Cycle too short"
                    return result
                result = self._repair_short_cycle(result)
            ok, issues = self._verify_cycle_exists(result.cycle)
            if not ok:
                if self.config.repair_cycles:
                    return self._repair_missing_edges(result, issues)
                result.success, result.error_message = False, f"Invalid cycle: {', '.join(issues)}"
                return result
            c, t = self._compute_cycle_weights_float(result.cycle[:-1])
            tol = 1e-10 if isinstance(result.ratio, (int, Fraction)) else 1e-6
            if (abs(c - float(result.sum_cost)) > tol or abs(t - float(result.sum_time)) > tol
                    or abs(c / t - float(result.ratio)) > tol):
                if not self.config.repair_cycles:
                    result.success, result.error_message = False, "Weight validation failed"
                    return result
                result.sum_cost, result.sum_time, result.ratio = c, t, c / t
            return result
        except Exception as exc:
            result.success, result.error_message = False, f"Validation failed: {exc}"
            return result

    def _has_edge(self, u: int, v: int) -> bool:
        """:1622-1629, through the CSR segment of v instead of a scan of the edge list."""
        self._build_arrays_if_needed()
        return self._hop_edges(u, v).size > 0

    def _repair_short_cycle(self, result: SolverResult) -> SolverResult:
        """:1584-1608.  Only a closed 2-cycle [u, v, u] is ever extended (to u -> w -> v -> u through the first w that has
        both edges); _validate_result calls this for cycles shorter than that (a self-loop [v, v]), which it leaves alone."""
        if len(result.cycle) == 3 and result.cycle[0] == result.cycle[2]:
            u, v = result.cycle[0], result.cycle[1]
            for w in range(self.n):
                if w != u and w != v and self._has_edge(u, w) and self._has_edge(w, v):
                    new_cycle = [u, w, v, u]
                    cost, time = self._compute_cycle_weights_float(new_cycle[:-1])
                    result.cycle, result.sum_cost, result.sum_time, result.ratio = new_cycle, cost, time, cost / time
                    return result
        return result

    def _repair_missing_edges(self, result: SolverResult, issues: list[str]) -> SolverResult:
        """:1610-1620: no repair is attempted, the result is marked failed with the reference's message."""
        result.success = False
        result.error_message = f"Could not repair cycle: {', '.join(issues)}"
        return result

    # ------------------------------------------------------------------ analytics (:1699-1759)
    def _certify_single_edge_scenarios(self, warm, pos, dc, dt, tol, slack):
        """
        For S scenarios "sorted-array edge pos[s] perturbed by (dc[s], dt[s])": list of (sum_cost, sum_time) of the cycle
        `warm` under the scenario's weights where that cycle is still optimal to within tol there, None elsewhere.
        Everything per scenario is array arithmetic; the probes run inside the library, K scenarios per edge sweep,
        warm-started from one base probe, sharded over the devices of config.gpu_devices (independent units: no
        collective on the data path).
        """
        cyc = warm[:-1]
        sc0, st0 = self._compute_cycle_weights_float(cyc)
        S = len(pos)
        sums = [(sc0, st0)] * S
        lams = np.full(S, sc0 / st0 - 0.75 * tol)
        # scenarios on a (u, v) pair the cycle uses: the hop rule (:1309-1313) may pick another parallel edge there
        hop_keys = np.array([u * self.n + v for u, v in zip(cyc, cyc[1:] + cyc[:1])], dtype=np.int64)
        touched = np.flatnonzero(np.isin(self._U[pos] * self.n + self._V[pos], hop_keys))
        sums = list(sums)
        for j in touched:
            p1 = np.array([pos[j]])
            saved = (self._C[p1].copy(), self._T[p1].copy())   # restore bit-exactly, not by subtracting
            self._C[p1] += dc[j]; self._T[p1] += dt[j]
            try:
                sums[j] = self._compute_cycle_weights_float(cyc)
            finally:
                self._C[p1], self._T[p1] = saved
            lams[j] = sums[j][0] / sums[j][1] - 0.75 * tol if sums[j][1] > 0 else np.inf
        K = max(1, min(int(self.config.gpu_scenarios_per_sweep), N.KMAX))
        lam_base = sc0 / st0 - 0.75 * tol
        graphs = [self._dev] if self._mg is None else [self._mg.replica(i) for i in range(self._comm.nranks)]
        verdict = np.full(S, -1, np.int32)
        if len(graphs) == 1 or S < 4 * K * len(graphs):
            verdict[:], _ = graphs[0].scenarios_certify(pos, dc, dt, lams, lam_base, slack, K)
        else:
            from concurrent.futures import ThreadPoolExecutor
            from .sharded import shard_units
            parts = [shard_units(S, r, len(graphs)) for r in range(len(graphs))]

            def run(r):
                sl = slice(parts[r].start, parts[r].stop)
                return graphs[r].scenarios_certify(pos[sl], dc[sl], dt[sl], lams[sl], lam_base, slack, K)[0]
            with ThreadPoolExecutor(len(graphs)) as ex:      # ctypes releases the GIL: one host thread per device
                for r, v in enumerate(ex.map(run, range(len(graphs)))):
                    verdict[parts[r].start:parts[r].stop] = v
        return [sums[j] if verdict[j] == N.V_NONNEG and sums[j][1] > 0 else None for j in range(S)]

    def sensitivity_analysis(self, edge_perturbations: list[dict[int, tuple[float, float]]]) -> list[SolverResult]:
        """
        One re-solve per scenario {edge index: (delta_cost, delta_time)} (:1699-1729).  Edge indices are
        insertion-order indices into the (preprocessed) edge list, as in the reference.  The perturbation is
        applied to the device-resident arrays (mrc_graph_patch_edges) instead of rebuilding Edge objects and
        arrays per scenario; reported sums use the perturbed weights.
        """
        self._preprocess_graph()
        self._build_arrays_if_needed()
        # warm start: the unperturbed optimum is still a cycle of every perturbed graph, so its (re-costed)
        # ratio is a valid upper end of the bracket and usually certifies in one round
        warm = None
        if self._last_result is not None and self._last_result.success and self._last_result.cycle:
            ok, _ = self._verify_cycle_exists(self._last_result.cycle)
            if ok:
                warm = list(self._last_result.cycle)
        m = self._num_edges()
        inv = self._inv_order()
        scen_arrays = []
        S = len(edge_perturbations)
        single = None   # (pos, dc, dt) arrays when every scenario touches exactly one edge (the sweep of BASELINE config 5)
        if S and all(len(sc) == 1 for sc in edge_perturbations):
            keys = np.fromiter((next(iter(sc)) for sc in edge_perturbations), dtype=np.int64, count=S)
            vals = np.array([next(iter(sc.values())) for sc in edge_perturbations], dtype=np.float64).reshape(S, 2)
            if np.any(keys < 0) or np.any(keys >= m):
                raise IndexError("edge index out of range")
            pos_all = inv[keys]
            if np.any(self._T[pos_all] + vals[:, 1] <= 0):
                raise ValueError("Edge transit time must be strictly positive")
            single = (pos_all, np.ascontiguousarray(vals[:, 0]), np.ascontiguousarray(vals[:, 1]))
            scen_arrays = [(pos_all[i:i + 1], single[1][i:i + 1], single[2][i:i + 1]) for i in range(S)]
        else:
            for scenario in edge_perturbations:
                for idx in scenario:
                    if idx < 0 or idx >= m:
                        raise IndexError("edge index out of range")
                pos = inv[np.fromiter(scenario.keys(), dtype=np.int64, count=len(scenario))]
                dc = np.array([float(d[0]) for d in scenario.values()], dtype=np.float64)
                dt = np.array([float(d[1]) for d in scenario.values()], dtype=np.float64)
                if np.any(self._T[pos] + dt <= 0):
                    raise ValueError("Edge transit time must be strictly positive")
                scen_arrays.append((pos, dc, dt))
        results: list[SolverResult | None] = [None] * len(scen_arrays)
        tol, slack = self.config.numeric_tolerance, self.config.numeric_cycle_slack

        # Fast path, one library call for all single-edge scenarios (mrc_scenarios_certify): each one probes its own
        # certifier, (ratio of the unperturbed optimum under that scenario's weights) - 0.75 tol; no negative cycle there
        # means that cycle is still optimal to within tol -- the answer the sequential path below would give.  Scenarios
        # whose optimum moves (or that touch several edges) fall through to it.
        if warm is not None:
            singles = np.arange(S, dtype=np.int64) if single is not None else \
                np.array([i for i, sa in enumerate(scen_arrays) if len(sa[0]) == 1], dtype=np.int64)
            if len(singles):
                if single is not None:
                    keep = self._certify_single_edge_scenarios(warm, single[0], single[1], single[2], tol, slack)
                else:
                    keep = self._certify_single_edge_scenarios(warm, np.array([scen_arrays[i][0][0] for i in singles]),
                                                               np.array([scen_arrays[i][1][0] for i in singles]),
                                                               np.array([scen_arrays[i][2][0] for i in singles]), tol, slack)
                for i, k in zip(singles, keep):
                    if k is not None:
                        results[i] = SolverResult(cycle=list(warm), sum_cost=k[0], sum_time=k[1], ratio=k[0] / k[1], success=True,
                                                  metrics=SolverMetrics(iterations=1, mode_used="numeric"),
                                                  config_used=self.config)
        for i, (pos, dc, dt) in enumerate(scen_arrays):
            if results[i] is not None:
                continue
            saved = (self._C[pos].copy(), self._T[pos].copy())
            self._dev.patch_edges(pos, dc, dt)
            self._C[pos] += dc
            self._T[pos] += dt
            try:
                res = self._solve_numeric(_warm_cycle=warm, _single_device=True)   # only replica 0 carries the patch
                if self.config.validate_cycles:
                    res = self._validate_result(res)
                res.config_used = self.config
                if self._metrics:
                    res.metrics = SolverMetrics(iterations=self._metrics.iterations, mode_used="numeric")
                results[i] = res
            finally:
                self._dev.restore()
                self._C[pos], self._T[pos] = saved[0], saved[1]
        return results

    def stability_region(self, epsilon: float = 0.01) -> dict[int, bool]:
        """{edge index: does the optimal cycle survive cost*(1+eps), time*(1+eps) on that edge} (:1731-1759).  The
        reference re-solves once per edge; here all m single-edge scenarios go through ONE batched call
        (mrc_scenarios_certify) and only the edges whose scenario moves the optimum are re-solved."""
        base = self._last_result if self._last_result is not None else self.solve()
        self._preprocess_graph()
        self._build_arrays_if_needed()
        m = self._num_edges()
        ok, _ = self._verify_cycle_exists(base.cycle) if base.cycle else (False, [])
        if not ok:
            res = self.sensitivity_analysis([{idx: (float(self._C[p]) * epsilon, float(self._T[p]) * epsilon)}
                                             for idx, p in enumerate(np.argsort(self._order))])
            return {idx: res[idx].cycle == base.cycle for idx in range(m)}
        inv = self._inv_order()                              # insertion index -> sorted position
        keep = self._certify_single_edge_scenarios(list(base.cycle), inv, self._C[inv] * epsilon, self._T[inv] * epsilon,
                                                   self.config.numeric_tolerance, self.config.numeric_cycle_slack)
        out = {}
        moved = [idx for idx in range(m) if keep[idx] is None]
        for idx in range(m):
            out[idx] = keep[idx] is not None
        if moved:
            res = self.sensitivity_analysis([{idx: (float(self._C[inv[idx]]) * epsilon, float(self._T[inv[idx]]) * epsilon)} for idx in moved])
            for idx, r in zip(moved, res):
                out[idx] = r.cycle == base.cycle
        return out

    # ------------------------------------------------------------------ misc (:1631-1693, :1814-1933)
    def get_debug_info(self) -> str:
        lines = ["Min Ratio Cycle Solver - Debug Info", "=" * 50, f"Vertices: {self.n}",
                 f"Edges: {self._num_edges()}", f"Integer weights: {self._all_int}"]
        props = self.get_graph_properties()
        if "error" not in props:
            lines.append(f"Density: {props['density']:.4f}")
            lines.append(f"Is sparse: {props['is_sparse']}")
        for issue in self.detect_issues():
            lines.append(f"  [{issue['level']}] {issue['message']}")
        if self._last_result is not None and self._last_result.success:
            lines.append(f"  Ratio: {self._last_result.ratio}")
            lines.append(f"  Cycle length: {len(self._last_result.cycle) - 1}")
        return "\n".join(lines)

    def visualize_solution(self, *args, **kwargs):
        """Plotting (:1761-1812) is outside the hot path and not provided here; the result objects are the
        reference's, so the reference's own plotting helpers can be fed with them."""
        raise NotImplementedError("visualisation is out of scope of the GPU hot path (SURVEY.md section 2)")

    create_interactive_plot = visualize_solution

    def health_check(self) -> dict[str, Any]:
        checks = []
        ndev = N.device_count()
        checks.append({"name": "cuda_device", "status": "PASS" if ndev > 0 else "FAIL",
                       "details": f"{ndev} CUDA device(s)", "recommendation": None if ndev else "needs a B200"})
        if self._num_edges() == 0:
            checks.append({"name": "graph_state", "status": "WARN", "details": "No edges in graph",
                           "recommendation": "Add edges before solving"})
        else:
            warn = sum(1 for i in self.detect_issues() if i["level"] == "WARNING")
            checks.append({"name": "graph_state", "status": "WARN" if warn else "PASS",
                           "details": f"{warn} warnings found" if warn else "Graph structure looks good",
                           "recommendation": "Review warnings" if warn else None})
        n_pass = sum(c["status"] == "PASS" for c in checks)
        n_warn = sum(c["status"] == "WARN" for c in checks)
        n_fail = sum(c["status"] == "FAIL" for c in checks)
        return {"timestamp": _time.time(), "checks": checks,
                "summary": {"total": len(checks), "passed": n_pass, "warnings": n_warn, "failures": n_fail,
                            "overall_status": "FAIL" if n_fail else ("WARN" if n_warn else "PASS")}}

    def __repr__(self) -> str:
        return f"MinRatioCycleSolver(n_vertices={self.n}, n_edges={self._num_edges()}, integer_mode={self._all_int})"


def solve_min_ratio_cycle(edges, n_vertices: int | None = None, mode: SolverMode | str = SolverMode.AUTO,
                          config: SolverConfig | None = None) -> SolverResult:
    """Convenience wrapper (:1941-1969)."""
    if n_vertices is None:
        n_vertices = max((max(u, v) for u, v, _, _ in edges), default=0) + 1
    solver = MinRatioCycleSolver(n_vertices, config)
    solver.add_edges(edges)
    return solver.solve(mode=mode)
