"""
ctypes binding of libmrc_b200.so (C ABI: include/mrc_b200.h).

There is NO fallback: if the shared library is missing or no CUDA device is present the calls raise
`NativeError` -- the product never routes through a CPU implementation.
"""

from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MRC_B200_LIB") or os.path.join(HERE, "libmrc_b200.so")

KMAX = 8

OK, ERR_CUDA, ERR_ARG, ERR_NO_DEVICE, ERR_OVERFLOW, ERR_NO_CYCLE, ERR_CONVERGENCE = range(7)
ERR_LOWER_NEG, ERR_UPPER_NONNEG, ERR_SB_STEPS, ERR_NOT_INTEGER, ERR_TIMEOUT, ERR_EDGE_MISSING, ERR_COMM = range(7, 14)
UNIQUE_ID_BYTES = 128

V_NONNEG, V_NEG, V_NEG_NOCYCLE, V_TIE, V_SKIPPED_NEG, V_SKIPPED_NONNEG, V_ABANDONED = range(7)
F_MONOTONE_PRUNE = 1
F_STOP_ON_NEG = 2
F_PATIENCE = 4
F_NO_CYCLES = 8
F_CERT_FIRST = 16

EXPORTS = [
    "mrc_abi_version", "mrc_last_error", "mrc_device_count", "mrc_graph_create", "mrc_graph_destroy",
    "mrc_graph_info", "mrc_numeric_bounds", "mrc_graph_properties", "mrc_get_stats", "mrc_reset_stats", "mrc_graph_patch_edges",
    "mrc_graph_restore", "mrc_probe_numeric", "mrc_solve_numeric", "mrc_plan_candidates", "mrc_reduce_round",
    "mrc_probe_exact", "mrc_potentials_exact", "mrc_tight_mask_exact", "mrc_zero_cycle_exact", "mrc_solve_exact",
    "mrc_bench_relax", "mrc_set_option", "mrc_batch_solve_numeric", "mrc_fetch_cycle", "mrc_fetch_cycle_sums", "mrc_sort_edges_by_dst", "mrc_probe_numeric_scenarios",
    "mrc_comm_unique_id", "mrc_comm_init_rank", "mrc_comm_init", "mrc_comm_destroy", "mrc_comm_info",
    "mrc_graph_create_replicated", "mrc_solve_numeric_sharded", "mrc_mgraph_create", "mrc_mgraph_destroy", "mrc_mgraph_graph",
    "mrc_mgraph_solve_numeric", "mrc_shard_fold", "mrc_scenarios_certify",
    "mrc_graph_build", "mrc_graph_fetch_arrays", "mrc_cycle_weights", "mrc_first_round_levels",
]


class NativeError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libmrc_b200 error {code}: {message}")
        self.code = code
        self.message = message


class Stats(C.Structure):
    _fields_ = [
        ("relax_launches", C.c_int64), ("relax_edge_visits", C.c_int64), ("check_launches", C.c_int64),
        ("probes", C.c_int64), ("rounds", C.c_int64), ("other_launches", C.c_int64),
        ("relax_ms", C.c_double), ("check_ms", C.c_double), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
        ("false_cycles_cut", C.c_int64), ("relax_edge_slots", C.c_int64),
        ("sparse_passes", C.c_int64), ("sparse_ms", C.c_double),
        ("kernel_launches", C.c_int64), ("probe_launches", C.c_int64), ("probe_ms", C.c_double),
        ("crowded_sweeps", C.c_int64), ("crowded_ms", C.c_double),
        ("check_doubling_ms", C.c_double), ("check_extract_ms", C.c_double), ("check_rounds", C.c_int64),
        ("relax_stream_bytes", C.c_int64),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None


def lib():
    """Load the shared library (building it is `python -m min_ratio_cycle_b200.build`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeError(-1, f"{LIB_PATH} is missing: build it with `python -m min_ratio_cycle_b200.build` "
                              "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    p32, p64, pd = C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_double)
    pu8, pi = C.POINTER(C.c_uint8), C.POINTER(C.c_int)
    vp = C.c_void_p
    L.mrc_abi_version.restype = C.c_int
    L.mrc_last_error.restype = C.c_char_p
    L.mrc_device_count.argtypes = [pi]
    L.mrc_graph_create.argtypes = [C.c_int64, C.c_int64, p32, p32, pd, pd, p64, p64, C.c_int, C.POINTER(vp)]
    L.mrc_graph_destroy.argtypes = [vp]
    L.mrc_graph_info.argtypes = [vp, p64, p64, pi, pi]
    L.mrc_numeric_bounds.argtypes = [vp, pd, pd]
    L.mrc_graph_properties.argtypes = [vp, pd]
    L.mrc_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.mrc_reset_stats.argtypes = [vp]
    L.mrc_graph_patch_edges.argtypes = [vp, C.c_int64, p64, pd, pd]
    L.mrc_graph_restore.argtypes = [vp]
    L.mrc_probe_numeric.argtypes = [vp, pd, C.c_int, C.c_double, C.c_uint32, p32, p32, p32, C.c_int64, pd, pd, p64]
    L.mrc_solve_numeric.argtypes = [vp, C.c_double, C.c_double, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int,
                                    C.c_double, pd, pd, pd, p32, C.c_int64, p64, p32, pd, pd]
    L.mrc_plan_candidates.argtypes = [C.c_double, C.c_double, C.c_int, C.c_double, C.c_int, pd]
    L.mrc_reduce_round.argtypes = [C.c_int, pd, p32, pd, pd, pd, pi, pi]
    L.mrc_probe_exact.argtypes = [vp, p64, p64, C.c_int, C.c_uint32, p32, p32, p32, C.c_int64, p64, p64, p64]
    L.mrc_potentials_exact.argtypes = [vp, C.c_int64, C.c_int64, pu8, p64]
    L.mrc_tight_mask_exact.argtypes = [vp, C.c_int64, C.c_int64, pu8, pu8]
    L.mrc_zero_cycle_exact.argtypes = [vp, C.c_int64, C.c_int64, pu8, p32, C.c_int64, p64, p64, p64]
    L.mrc_solve_exact.argtypes = [vp, C.c_int64, C.c_int64, C.c_int, p32, C.c_int64, p64, p64, p64, p64, p64, p64,
                                  p64, C.c_int64, p64]
    L.mrc_sort_edges_by_dst.argtypes = [C.c_int64, C.c_int64, p32, p32, pd, pd, C.c_int, p32, p32, pd, pd, p32, p64]
    L.mrc_probe_numeric_scenarios.argtypes = [vp, pd, C.c_int, C.c_double, C.c_uint32, p64, pd, pd, p32, p32, p32, C.c_int64,
                                              pd, pd, p64]
    L.mrc_fetch_cycle.argtypes = [vp, C.c_int, p32, C.c_int64, p64]
    L.mrc_fetch_cycle_sums.argtypes = [vp, C.c_int, p32, C.c_int64, p64, pd, pd]
    L.mrc_bench_relax.argtypes = [vp, pd, C.c_int, C.c_int, C.POINTER(C.c_float)]
    L.mrc_set_option.argtypes = [vp, C.c_char_p, C.c_int]
    L.mrc_batch_solve_numeric.argtypes = [C.c_int64, p32, p64, p32, p32, pd, pd, C.c_int, C.c_double, C.c_double, C.c_int,
                                          C.c_int, pd, pd, pd, p32, p32, C.c_int64, p32, p32, C.POINTER(Stats)]
    L.mrc_comm_unique_id.argtypes = [pu8]
    L.mrc_comm_init_rank.argtypes = [C.c_int, C.c_int, pu8, C.c_int, C.POINTER(vp)]
    L.mrc_comm_init.argtypes = [C.c_int, pi, C.POINTER(vp)]
    L.mrc_comm_destroy.argtypes = [vp]
    L.mrc_comm_info.argtypes = [vp, pi, pi, pi]
    L.mrc_graph_create_replicated.argtypes = [vp, C.c_int, C.c_int64, C.c_int64, p32, p32, pd, pd, C.POINTER(vp)]
    sharded_tail = [C.c_double, C.c_double, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, pd, pd, pd, p32, C.c_int64,
                    p64, p32, pd, pd, p32]
    L.mrc_solve_numeric_sharded.argtypes = [vp, vp] + sharded_tail
    L.mrc_mgraph_create.argtypes = [vp, C.c_int64, C.c_int64, p32, p32, pd, pd, C.POINTER(vp)]
    L.mrc_mgraph_destroy.argtypes = [vp]
    L.mrc_mgraph_graph.argtypes = [vp, C.c_int]
    L.mrc_mgraph_solve_numeric.argtypes = [vp] + sharded_tail
    L.mrc_shard_fold.argtypes = [C.c_int, pd, pd, pd, pi, pi, pu8]
    L.mrc_graph_build.argtypes = [C.c_int64, C.c_int64, vp, vp, C.c_int, pd, pd, C.c_int, C.c_int, C.POINTER(vp), p64, p64, p64]
    L.mrc_graph_fetch_arrays.argtypes = [vp, p32, p32, pd, pd, p32, pu8]
    L.mrc_cycle_weights.argtypes = [vp, C.c_int64, p32, pd, pd, p32, p32]
    L.mrc_first_round_levels.argtypes = [C.c_int, pd]
    L.mrc_scenarios_certify.argtypes = [vp, C.c_int64, p64, pd, pd, pd, C.c_double, C.c_double, C.c_int, p32, p64]
    for name in EXPORTS:
        if name == "mrc_mgraph_graph":
            getattr(L, name).restype = vp
        elif name not in ("mrc_last_error",):
            getattr(L, name).restype = C.c_int
    _lib = L
    return L


def check(rc: int):
    if rc != OK:
        raise NativeError(rc, lib().mrc_last_error().decode("utf-8", "replace"))


def device_count() -> int:
    c = C.c_int(0)
    rc = lib().mrc_device_count(C.byref(c))
    return c.value if rc == OK else 0


def _ptr(a, ct):
    return None if a is None else a.ctypes.data_as(C.POINTER(ct))


class DeviceGraph:
    """
    Device-resident destination-sorted edge arrays (mrc_graph).  Inputs are the reference's
    `_U, _V, _C, _T[, _Ci, _Ti]` (solver.py:488-498) in any integer/float dtype; they are narrowed to
    int32/f64/int64 here.
    """

    def __init__(self, n, U, V, Cw, Tw, Ci=None, Ti=None, device: int = 0):
        L = lib()
        self._h = None
        self.n = int(n)
        self.m = int(len(U))
        src = np.ascontiguousarray(U, dtype=np.int32)
        dst = np.ascontiguousarray(V, dtype=np.int32)
        cost = np.ascontiguousarray(Cw, dtype=np.float64)
        time = np.ascontiguousarray(Tw, dtype=np.float64)
        ic = None if Ci is None else np.ascontiguousarray(Ci, dtype=np.int64)
        it = None if Ti is None else np.ascontiguousarray(Ti, dtype=np.int64)
        h = C.c_void_p()
        check(L.mrc_graph_create(self.n, self.m, _ptr(src, C.c_int32), _ptr(dst, C.c_int32), _ptr(cost, C.c_double),
                                 _ptr(time, C.c_double), _ptr(ic, C.c_int64), _ptr(it, C.c_int64), int(device),
                                 C.byref(h)))
        self._h = h
        self.device = int(device)
        self.has_int = ic is not None

    @classmethod
    def _wrap(cls, handle, n: int, m: int, device: int, owned: bool = True):
        g = cls.__new__(cls)
        g._h, g.n, g.m, g.device, g.has_int, g._owned = handle, int(n), int(m), int(device), False, owned
        return g

    @classmethod
    def replicated(cls, comm: "Comm", n: int, m: int, U=None, V=None, Cw=None, Tw=None, root: int = 0):
        """Replica of a graph on this rank (mrc_graph_create_replicated): the root passes the destination-sorted
        arrays, the other ranks None; one H2D on the root + NCCL broadcast.  Collective over the ranks of `comm`."""
        arrs = [None] * 4
        if U is not None:
            arrs = [np.ascontiguousarray(U, dtype=np.int32), np.ascontiguousarray(V, dtype=np.int32),
                    np.ascontiguousarray(Cw, dtype=np.float64), np.ascontiguousarray(Tw, dtype=np.float64)]
        h = C.c_void_p()
        check(lib().mrc_graph_create_replicated(comm._h, int(root), int(n), int(m), _ptr(arrs[0], C.c_int32),
                                                _ptr(arrs[1], C.c_int32), _ptr(arrs[2], C.c_double),
                                                _ptr(arrs[3], C.c_double), C.byref(h)))
        return cls._wrap(h, n, m, comm.device)

    @classmethod
    def build(cls, n: int, U, V, Cw, Tw, remove_dominated: bool = True, device: int = 0):
        """mrc_graph_build: insertion-order host arrays in; preprocessing, stable sort by destination and CSR on the
        device.  Returns (graph, dup_pairs, removed); graph.m = edges left."""
        U = np.ascontiguousarray(U)
        V = np.ascontiguousarray(V)
        if U.dtype != V.dtype or U.dtype not in (np.int32, np.int64):
            U, V = U.astype(np.int64), V.astype(np.int64)
        Cw = np.ascontiguousarray(Cw, dtype=np.float64)
        Tw = np.ascontiguousarray(Tw, dtype=np.float64)
        h = C.c_void_p()
        mo, dups, rem = C.c_int64(), C.c_int64(), C.c_int64()
        check(lib().mrc_graph_build(int(n), len(U), U.ctypes.data_as(C.c_void_p), V.ctypes.data_as(C.c_void_p), U.dtype.itemsize,
                                    _ptr(Cw, C.c_double), _ptr(Tw, C.c_double), int(bool(remove_dominated)), int(device),
                                    C.byref(h), C.byref(mo), C.byref(dups), C.byref(rem)))
        g = cls._wrap(h, n, mo.value, device)
        g.m_input = len(U)
        return g, dups.value, rem.value

    def fetch_arrays(self, order: bool = False, keep: bool = False):
        """(src int32, dst int32, cost, time[, order int32][, keep uint8]) of the destination-sorted arrays."""
        src, dst = np.empty(self.m, np.int32), np.empty(self.m, np.int32)
        cost, time = np.empty(self.m, np.float64), np.empty(self.m, np.float64)
        o = np.empty(self.m, np.int32) if order else None
        k = np.empty(getattr(self, "m_input", self.m), np.uint8) if keep else None
        check(lib().mrc_graph_fetch_arrays(self._h, _ptr(src, C.c_int32), _ptr(dst, C.c_int32), _ptr(cost, C.c_double),
                                           _ptr(time, C.c_double), _ptr(o, C.c_int32), _ptr(k, C.c_uint8)))
        return tuple(x for x in (src, dst, cost, time, o, k) if x is not None)

    def fetch_keep(self):
        k = np.empty(getattr(self, "m_input", self.m), np.uint8)
        check(lib().mrc_graph_fetch_arrays(self._h, None, None, None, None, None, _ptr(k, C.c_uint8)))
        return k.astype(bool)

    def cycle_weights(self, closed_cycle):
        """mrc_cycle_weights: (sum_cost, sum_time, missing_hop or -1, hop edge indices)."""
        cyc = np.ascontiguousarray(closed_cycle, dtype=np.int32)
        L = max(len(cyc) - 1, 0)
        sc, st, miss = C.c_double(), C.c_double(), C.c_int32(-1)
        he = np.full(L, -1, np.int32)
        check(lib().mrc_cycle_weights(self._h, L, _ptr(cyc, C.c_int32), C.byref(sc), C.byref(st), C.byref(miss), _ptr(he, C.c_int32)))
        return sc.value, st.value, miss.value, he

    def close(self):
        if self._h is not None:
            if getattr(self, "_owned", True):
                lib().mrc_graph_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- bookkeeping ----------------------------------------------------
    def stats(self) -> dict:
        s = Stats()
        check(lib().mrc_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    def reset_stats(self):
        check(lib().mrc_reset_stats(self._h))

    def bounds(self):
        lo, hi = C.c_double(), C.c_double()
        check(lib().mrc_numeric_bounds(self._h, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def properties(self):
        """Degree and weight statistics of get_graph_properties (solver.py:339-387), computed on the device."""
        out = np.zeros(13)
        check(lib().mrc_graph_properties(self._h, _ptr(out, C.c_double)))
        keys = ("isolated_vertices", "source_vertices", "sink_vertices", "max_in_degree", "max_out_degree",
                "cost_min", "cost_max", "cost_mean", "cost_std", "time_min", "time_max", "time_mean", "time_std")
        return dict(zip(keys, out.tolist()))

    def patch_edges(self, idx, dcost, dtime):
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        dc = np.ascontiguousarray(dcost, dtype=np.float64)
        dt = np.ascontiguousarray(dtime, dtype=np.float64)
        check(lib().mrc_graph_patch_edges(self._h, len(idx), _ptr(idx, C.c_int64), _ptr(dc, C.c_double),
                                          _ptr(dt, C.c_double)))

    def restore(self):
        check(lib().mrc_graph_restore(self._h))

    def set_option(self, name: str, value: int):
        check(lib().mrc_set_option(self._h, name.encode(), int(value)))

    def bench_relax(self, lams, passes: int):
        lam = np.ascontiguousarray(lams, dtype=np.float64)
        ms = np.zeros(passes, np.float32)
        check(lib().mrc_bench_relax(self._h, _ptr(lam, C.c_double), len(lam), int(passes), _ptr(ms, C.c_float)))
        return ms

    # -- probes -----------------------------------------------------------
    def probe_numeric(self, lams, slack: float = 1e-15, flags: int = 0, cyc_cap: int | None = None):
        """K-batched analogue of _detect_negative_cycle_numeric.  Returns a list of dicts."""
        lam = np.ascontiguousarray(lams, dtype=np.float64)
        K = len(lam)
        cap = int(cyc_cap if cyc_cap is not None else min(self.n + 1, 1024))   # longer cycles are fetched below
        vd = np.zeros(K, np.int32); cl = np.zeros(K, np.int32); buf = np.empty(K * cap, np.int32)
        sc = np.zeros(K); st = np.zeros(K); ps = np.zeros(K, np.int64)
        check(lib().mrc_probe_numeric(self._h, _ptr(lam, C.c_double), K, float(slack), int(flags), _ptr(vd, C.c_int32),
                                      _ptr(cl, C.c_int32), _ptr(buf, C.c_int32), cap, _ptr(sc, C.c_double),
                                      _ptr(st, C.c_double), _ptr(ps, C.c_int64)))
        out = []
        for k in range(K):
            cyc = None if (int(flags) & F_NO_CYCLES) else self._cycle_of(k, buf, cap, int(cl[k]))
            out.append(dict(lam=float(lam[k]), verdict=int(vd[k]), cycle=cyc, cycle_len=int(cl[k]),
                            sum_cost=float(sc[k]), sum_time=float(st[k]), passes=int(ps[k])))
        return out

    def probe_numeric_scenarios(self, lams, patch_edge, dcost, dtime, slack: float = 1e-15, cyc_cap: int | None = None):
        """K single-edge perturbation scenarios in one batched probe (mrc_probe_numeric_scenarios)."""
        lam = np.ascontiguousarray(lams, dtype=np.float64)
        pe = np.ascontiguousarray(patch_edge, dtype=np.int64)
        dc = np.ascontiguousarray(dcost, dtype=np.float64); dt = np.ascontiguousarray(dtime, dtype=np.float64)
        K = len(lam)
        cap = int(cyc_cap if cyc_cap is not None else min(self.n + 1, 1024))
        vd = np.zeros(K, np.int32); cl = np.zeros(K, np.int32); buf = np.empty(K * cap, np.int32)
        sc = np.zeros(K); st = np.zeros(K); ps = np.zeros(K, np.int64)
        check(lib().mrc_probe_numeric_scenarios(self._h, _ptr(lam, C.c_double), K, float(slack), 0, _ptr(pe, C.c_int64),
                                                _ptr(dc, C.c_double), _ptr(dt, C.c_double), _ptr(vd, C.c_int32),
                                                _ptr(cl, C.c_int32), _ptr(buf, C.c_int32), cap, _ptr(sc, C.c_double),
                                                _ptr(st, C.c_double), _ptr(ps, C.c_int64)))
        return [dict(lam=float(lam[k]), verdict=int(vd[k]), cycle=self._cycle_of(k, buf, cap, int(cl[k])),
                     sum_cost=float(sc[k]), sum_time=float(st[k]), passes=int(ps[k])) for k in range(K)]

    def scenarios_certify(self, edge, dcost, dtime, lams, lam_base: float, slack: float = 1e-15, K: int = 8):
        """mrc_scenarios_certify: (verdicts int32[S], passes int64[S]) for S single-edge scenarios."""
        edge = np.ascontiguousarray(edge, dtype=np.int64)
        dc = np.ascontiguousarray(dcost, dtype=np.float64)
        dt = np.ascontiguousarray(dtime, dtype=np.float64)
        lam = np.ascontiguousarray(lams, dtype=np.float64)
        S = len(edge)
        verdict = np.full(S, -1, np.int32)
        passes = np.zeros(S, np.int64)
        check(lib().mrc_scenarios_certify(self._h, S, _ptr(edge, C.c_int64), _ptr(dc, C.c_double), _ptr(dt, C.c_double),
                                          _ptr(lam, C.c_double), float(lam_base), float(slack), int(K),
                                          _ptr(verdict, C.c_int32), _ptr(passes, C.c_int64)))
        return verdict, passes

    def probe_exact(self, a, b, flags: int = 0, cyc_cap: int | None = None):
        a = np.ascontiguousarray(a, dtype=np.int64); b = np.ascontiguousarray(b, dtype=np.int64)
        K = len(a)
        cap = int(cyc_cap if cyc_cap is not None else min(self.n + 1, 1024))
        vd = np.zeros(K, np.int32); cl = np.zeros(K, np.int32); buf = np.empty(K * cap, np.int32)
        sc = np.zeros(K, np.int64); st = np.zeros(K, np.int64); ps = np.zeros(K, np.int64)
        check(lib().mrc_probe_exact(self._h, _ptr(a, C.c_int64), _ptr(b, C.c_int64), K, int(flags), _ptr(vd, C.c_int32),
                                    _ptr(cl, C.c_int32), _ptr(buf, C.c_int32), cap, _ptr(sc, C.c_int64),
                                    _ptr(st, C.c_int64), _ptr(ps, C.c_int64)))
        out = []
        for k in range(K):
            cyc = None if (int(flags) & F_NO_CYCLES) else self._cycle_of(k, buf, cap, int(cl[k]))
            out.append(dict(a=int(a[k]), b=int(b[k]), verdict=int(vd[k]), cycle=cyc, cycle_len=int(cl[k]),
                            sum_cost=int(sc[k]), sum_time=int(st[k]), passes=int(ps[k])))
        return out

    def fetch_cycle_sums(self, k: int, length: int | None = None):
        """(closed cycle, sum_cost, sum_time) of candidate k of the last probe (probes run with F_NO_CYCLES)."""
        cap = int(length) if length else self.n + 1
        buf = np.empty(cap, np.int32)
        n, sc, st = C.c_int64(), C.c_double(), C.c_double()
        check(lib().mrc_fetch_cycle_sums(self._h, int(k), _ptr(buf, C.c_int32), cap, C.byref(n), C.byref(sc), C.byref(st)))
        return buf[: n.value].tolist(), sc.value, st.value

    def _cycle_of(self, k: int, buf, cap: int, length: int):
        if not length:
            return None
        if length <= cap:
            return buf[k * cap: k * cap + length].tolist()
        full = np.empty(length, np.int32)
        n = C.c_int64()
        check(lib().mrc_fetch_cycle(self._h, k, _ptr(full, C.c_int32), length, C.byref(n)))
        return full[: n.value].tolist()

    # -- searches -----------------------------------------------------------
    def solve_numeric(self, lo: float, hi: float, max_iter: int = 60, tol: float = 1e-12, slack: float = 1e-15,
                      K: int = 4, max_seconds: float = 0.0, hi_is_cycle: bool = False):
        """Returns dict(rc, cycle(closed), sum_cost, sum_time, ratio, iterations, lo, hi); rc may be
        ERR_CONVERGENCE with a valid best cycle, ERR_NO_CYCLE / ERR_TIMEOUT without one."""
        cap = self.n + 1
        buf = np.zeros(cap, np.int32)
        ratio, sc, st, flo, fhi = (C.c_double() for _ in range(5))
        cl = C.c_int64(); it = C.c_int32()
        rc = lib().mrc_solve_numeric(self._h, float(lo), float(hi), int(bool(hi_is_cycle)), int(max_iter), float(tol),
                                     float(slack), int(K),
                                     float(max_seconds or 0.0), C.byref(ratio), C.byref(sc), C.byref(st),
                                     _ptr(buf, C.c_int32), cap, C.byref(cl), C.byref(it), C.byref(flo), C.byref(fhi))
        if rc not in (OK, ERR_CONVERGENCE, ERR_NO_CYCLE, ERR_TIMEOUT):
            check(rc)
        have = rc in (OK, ERR_CONVERGENCE)
        return dict(rc=rc, message=lib().mrc_last_error().decode() if rc else "",
                    cycle=buf[: cl.value].tolist() if have else [], sum_cost=sc.value, sum_time=st.value,
                    ratio=ratio.value, iterations=it.value, lo=flo.value, hi=fhi.value)

    def solve_numeric_sharded(self, comm: "Comm", lo: float, hi: float, max_iter: int = 60, tol: float = 1e-12,
                              slack: float = 1e-15, K_local: int = 4, max_seconds: float = 0.0):
        """mrc_solve_numeric_sharded on this rank's replica; collective over the ranks of `comm`.  Same dict as
        solve_numeric plus `ticks`; every rank gets the same cycle."""
        return _sharded_call(lambda *tail: lib().mrc_solve_numeric_sharded(comm._h, self._h, *tail), self.n, lo, hi, max_iter, tol,
                             slack, K_local, max_seconds)

    def potentials_exact(self, a: int, b: int):
        ok = C.c_uint8()
        dist = np.zeros(self.n, np.int64)
        check(lib().mrc_potentials_exact(self._h, int(a), int(b), C.byref(ok), _ptr(dist, C.c_int64)))
        return bool(ok.value), dist

    def tight_mask_exact(self, a: int, b: int):
        ok = C.c_uint8()
        mask = np.zeros(self.m, np.uint8)
        check(lib().mrc_tight_mask_exact(self._h, int(a), int(b), C.byref(ok), _ptr(mask, C.c_uint8)))
        return bool(ok.value), mask

    def zero_cycle_exact(self, a: int, b: int):
        found = C.c_uint8()
        cap = self.n + 1
        buf = np.zeros(cap, np.int32)
        cl, sc, st = C.c_int64(), C.c_int64(), C.c_int64()
        check(lib().mrc_zero_cycle_exact(self._h, int(a), int(b), C.byref(found), _ptr(buf, C.c_int32), cap,
                                         C.byref(cl), C.byref(sc), C.byref(st)))
        if not found.value:
            return None
        return buf[: cl.value].tolist(), sc.value, st.value

    def solve_exact(self, max_den=None, max_steps=None, strategy: int = 1, probes_cap: int = 4096):
        """_stern_brocot_search analogue.  Returns dict(rc, cycle(open), sum_cost, sum_time, a, b, iterations, probes)."""
        cap = self.n + 1
        buf = np.zeros(cap, np.int32)
        pr = np.zeros(2 * probes_cap, np.int64)
        cl, sc, st, a, b, it, npb = (C.c_int64() for _ in range(7))
        rc = lib().mrc_solve_exact(self._h, -1 if max_den is None else int(max_den),
                                   -1 if max_steps is None else int(max_steps), int(strategy), _ptr(buf, C.c_int32),
                                   cap, C.byref(cl), C.byref(sc), C.byref(st), C.byref(a), C.byref(b), C.byref(it),
                                   _ptr(pr, C.c_int64), probes_cap, C.byref(npb))
        if rc not in (OK, ERR_LOWER_NEG, ERR_UPPER_NONNEG, ERR_SB_STEPS, ERR_OVERFLOW, ERR_EDGE_MISSING,
                      ERR_CONVERGENCE):
            check(rc)
        k = min(npb.value, probes_cap)
        return dict(rc=rc, message=lib().mrc_last_error().decode() if rc else "",
                    cycle=buf[: cl.value].tolist() if rc == OK else [], sum_cost=sc.value, sum_time=st.value,
                    a=a.value, b=b.value, iterations=it.value,
                    probes=[(int(pr[2 * i]), int(pr[2 * i + 1])) for i in range(k)])


def _sharded_call(fn, n, lo, hi, max_iter, tol, slack, K_local, max_seconds):
    cap = int(n) + 1
    buf = np.zeros(cap, np.int32)
    ratio, sc, st, flo, fhi = (C.c_double() for _ in range(5))
    cl = C.c_int64(); it = C.c_int32(); ticks = C.c_int32()
    rc = fn(float(lo), float(hi), int(max_iter), float(tol), float(slack), int(K_local), float(max_seconds or 0.0),
            C.byref(ratio), C.byref(sc), C.byref(st), _ptr(buf, C.c_int32), cap, C.byref(cl), C.byref(it), C.byref(flo),
            C.byref(fhi), C.byref(ticks))
    if rc not in (OK, ERR_CONVERGENCE, ERR_NO_CYCLE, ERR_TIMEOUT):
        check(rc)
    have = rc in (OK, ERR_CONVERGENCE)
    return dict(rc=rc, message=lib().mrc_last_error().decode() if rc else "", cycle=buf[: cl.value].tolist() if have else [],
                sum_cost=sc.value, sum_time=st.value, ratio=ratio.value, iterations=it.value, lo=flo.value, hi=fhi.value,
                ticks=ticks.value)


class Comm:
    """
    NCCL communicator owned by the library (mrc_comm).  Comm.from_torch(): one process per GPU, the unique id
    travels through the job's torch.distributed group.  Comm.group(devices): one process driving several GPUs.
    """

    def __init__(self, handle, nranks: int, rank: int, device: int):
        self._h, self.nranks, self.rank, self.device = handle, nranks, rank, device

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_uint8 * UNIQUE_ID_BYTES)()
        check(lib().mrc_comm_unique_id(buf))
        return bytes(buf)

    @classmethod
    def from_id(cls, nranks: int, rank: int, uid: bytes, device: int):
        buf = (C.c_uint8 * UNIQUE_ID_BYTES).from_buffer_copy(uid)
        h = C.c_void_p()
        check(lib().mrc_comm_init_rank(int(nranks), int(rank), buf, int(device), C.byref(h)))
        return cls(h, nranks, rank, device)

    @classmethod
    def from_torch(cls, device: int, group=None):
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        on_gpu = "nccl" in str(dist.get_backend(group))
        t = torch.zeros(UNIQUE_ID_BYTES, dtype=torch.uint8)
        if rank == 0:
            t = torch.frombuffer(bytearray(cls.unique_id()), dtype=torch.uint8).clone()
        if on_gpu:
            t = t.to(torch.device("cuda", device))
        dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        return cls.from_id(world, rank, bytes(t.cpu().numpy().tobytes()), device)

    @classmethod
    def group(cls, devices):
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        check(lib().mrc_comm_init(len(devices), devs, C.byref(h)))
        return cls(h, len(devices), -1, int(devices[0]))

    def close(self):
        if self._h is not None:
            lib().mrc_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MultiGraph:
    """Replicas of one destination-sorted graph on every device of a group communicator (mrc_mgraph)."""

    def __init__(self, comm: Comm, n, U, V, Cw, Tw):
        self._h = None
        self.comm, self.n, self.m = comm, int(n), int(len(U))
        src = np.ascontiguousarray(U, dtype=np.int32)
        dst = np.ascontiguousarray(V, dtype=np.int32)
        cost = np.ascontiguousarray(Cw, dtype=np.float64)
        time = np.ascontiguousarray(Tw, dtype=np.float64)
        h = C.c_void_p()
        check(lib().mrc_mgraph_create(comm._h, self.n, self.m, _ptr(src, C.c_int32), _ptr(dst, C.c_int32),
                                      _ptr(cost, C.c_double), _ptr(time, C.c_double), C.byref(h)))
        self._h = h

    def replica(self, i: int) -> DeviceGraph:
        """Borrowed handle of the replica on device i of the group (statistics, options, bounds)."""
        h = lib().mrc_mgraph_graph(self._h, int(i))
        if not h:
            raise IndexError(i)
        return DeviceGraph._wrap(C.c_void_p(h), self.n, self.m, -1, owned=False)

    def solve_numeric(self, lo: float, hi: float, max_iter: int = 60, tol: float = 1e-12, slack: float = 1e-15,
                      K_local: int = 4, max_seconds: float = 0.0):
        return _sharded_call(lambda *tail: lib().mrc_mgraph_solve_numeric(self._h, *tail), self.n, lo, hi, max_iter, tol, slack,
                             K_local, max_seconds)

    def close(self):
        if self._h is not None:
            lib().mrc_mgraph_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def shard_fold(rec, lo: float, hi: float, have_cycle: bool):
    """mrc_shard_fold: rec is an (nrec, 4) array of (lambda, state, ratio, length); returns (lo, hi, have_cycle, best_idx, implied)."""
    rec = np.ascontiguousarray(rec, dtype=np.float64).reshape(-1, 4)
    clo, chi = C.c_double(lo), C.c_double(hi)
    hc, best = C.c_int(int(have_cycle)), C.c_int(-1)
    imp = np.zeros(len(rec), np.uint8)
    check(lib().mrc_shard_fold(len(rec), _ptr(rec, C.c_double), C.byref(clo), C.byref(chi), C.byref(hc), C.byref(best),
                               _ptr(imp, C.c_uint8)))
    return clo.value, chi.value, bool(hc.value), best.value, imp.astype(bool)


def first_round_levels(K: int) -> np.ndarray:
    """Quantile levels of the first round of a search without a known cycle (mrc_first_round_levels)."""
    q = np.zeros(int(K), np.float64)
    check(lib().mrc_first_round_levels(int(K), _ptr(q, C.c_double)))
    return q


def plan_candidates(lo: float, hi: float, have_cycle: bool, tol: float, K_total: int) -> np.ndarray:
    out = np.zeros(K_total)
    check(lib().mrc_plan_candidates(float(lo), float(hi), int(bool(have_cycle)), float(tol), int(K_total),
                                    _ptr(out, C.c_double)))
    return out


def reduce_round(lam, verdict, cyc_ratio, lo: float, hi: float, have_cycle: bool):
    lam = np.ascontiguousarray(lam, dtype=np.float64)
    vd = np.ascontiguousarray(verdict, dtype=np.int32)
    cr = np.ascontiguousarray(cyc_ratio, dtype=np.float64)
    clo, chi = C.c_double(lo), C.c_double(hi)
    hc, best = C.c_int(int(bool(have_cycle))), C.c_int(-1)
    check(lib().mrc_reduce_round(len(lam), _ptr(lam, C.c_double), _ptr(vd, C.c_int32), _ptr(cr, C.c_double),
                                 C.byref(clo), C.byref(chi), C.byref(hc), C.byref(best)))
    return clo.value, chi.value, bool(hc.value), best.value


def batch_solve_numeric(nverts, edge_off, src, dst, cost, time, max_iter: int = 60, tol: float = 1e-12,
                        slack: float = 1e-15, K: int = 4, device: int = 0):
    """
    mrc_batch_solve_numeric: B small graphs (concatenated, each destination-sorted with local ids), one CTA
    per graph.  Returns dict of per-graph arrays + stats.
    """
    nv = np.ascontiguousarray(nverts, dtype=np.int32)
    off = np.ascontiguousarray(edge_off, dtype=np.int64)
    s_ = np.ascontiguousarray(src, dtype=np.int32); d_ = np.ascontiguousarray(dst, dtype=np.int32)
    c_ = np.ascontiguousarray(cost, dtype=np.float64); t_ = np.ascontiguousarray(time, dtype=np.float64)
    B = len(nv)
    cap = int(nv.max()) + 1
    ratio = np.zeros(B); sc = np.zeros(B); st = np.zeros(B)
    clen = np.zeros(B, np.int32); cbuf = np.zeros(B * cap, np.int32); it = np.zeros(B, np.int32)
    status = np.zeros(B, np.int32)
    stats = Stats()
    check(lib().mrc_batch_solve_numeric(B, _ptr(nv, C.c_int32), _ptr(off, C.c_int64), _ptr(s_, C.c_int32),
                                        _ptr(d_, C.c_int32), _ptr(c_, C.c_double), _ptr(t_, C.c_double), int(max_iter),
                                        float(tol), float(slack), int(K), int(device), _ptr(ratio, C.c_double),
                                        _ptr(sc, C.c_double), _ptr(st, C.c_double), _ptr(clen, C.c_int32),
                                        _ptr(cbuf, C.c_int32), cap, _ptr(it, C.c_int32), _ptr(status, C.c_int32),
                                        C.byref(stats)))
    cycles = [cbuf[g * cap: g * cap + clen[g]].tolist() for g in range(B)]
    return dict(status=status, ratio=ratio, sum_cost=sc, sum_time=st, cycles=cycles, iterations=it,
                stats=stats.as_dict())


def sort_edges_by_dst(n, U, V, Cw, Tw, device: int = 0, count_dups: bool = True):
    """mrc_sort_edges_by_dst: (src, dst, cost, time, order, dup_pairs) with dst stably sorted on the device."""
    src = np.ascontiguousarray(U, dtype=np.int32); dst = np.ascontiguousarray(V, dtype=np.int32)
    cost = np.ascontiguousarray(Cw, dtype=np.float64); time = np.ascontiguousarray(Tw, dtype=np.float64)
    m = len(src)
    so = np.empty(m, np.int32); do = np.empty(m, np.int32); co = np.empty(m); to = np.empty(m)
    order = np.empty(m, np.int32)
    dups = C.c_int64(-1)
    check(lib().mrc_sort_edges_by_dst(int(n), m, _ptr(src, C.c_int32), _ptr(dst, C.c_int32), _ptr(cost, C.c_double),
                                      _ptr(time, C.c_double), int(device), _ptr(so, C.c_int32), _ptr(do, C.c_int32),
                                      _ptr(co, C.c_double), _ptr(to, C.c_double), _ptr(order, C.c_int32),
                                      C.byref(dups) if count_dups else None))
    return so, do, co, to, order, dups.value
