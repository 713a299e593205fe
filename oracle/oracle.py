"""
ctypes front-end of the CPU oracle (oracle/mrc_oracle.c).

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module.  The
product package (min_ratio_cycle_b200) never imports it and has no CPU fallback.

The oracle restates the reference's hot path (min_ratio_cycle/solver.py:450-1337)
and is pinned against fixtures produced by the Python reference itself
(tests/golden/make_golden.py -> tests/golden/*.json; tests/test_oracle_golden.py).
"""

from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from dataclasses import dataclass
from fractions import Fraction

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libmrc_oracle.so")

ERR_NAMES = {
    0: "ok",
    1: "alloc",
    2: "no_cycle",        # NumericalInstabilityError("No negative cycle found")
    3: "convergence",     # ConvergenceError
    4: "reduceat",        # IndexError in np.minimum.reduceat
    5: "lower_neg",
    6: "upper_nonneg",
    7: "sb_steps",
    8: "edge_missing",
    9: "overflow",
}


class OracleError(RuntimeError):
    def __init__(self, code: int, where: str):
        super().__init__(f"oracle {where}: {ERR_NAMES.get(code, code)}")
        self.code = code
        self.kind = ERR_NAMES.get(code, str(code))


def build(force: bool = False) -> str:
    """Compile oracle/mrc_oracle.c with the committed Makefile."""
    src = os.path.join(_HERE, "mrc_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        p64, pd, pi, pu8 = C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_uint8)
        L.orc_max_threads.restype = C.c_int
        L.orc_graph_build.restype = C.c_void_p
        L.orc_graph_build.argtypes = [C.c_int64, C.c_int64, p64, p64, pd, pd, p64, p64]
        L.orc_graph_free.argtypes = [C.c_void_p]
        L.orc_graph_set_threads.argtypes = [C.c_void_p, C.c_int]
        L.orc_graph_set_quirks.argtypes = [C.c_void_p, C.c_int]
        L.orc_graph_arrays.argtypes = [C.c_void_p, p64, p64, pd, pd, p64, p64, p64]
        L.orc_remove_dominated.restype = C.c_int64
        L.orc_remove_dominated.argtypes = [C.c_int64, p64, p64, pd, pd, pu8]
        L.orc_numeric_bounds.argtypes = [C.c_void_p, pd, pd]
        L.orc_cycle_weights_float_pub.argtypes = [C.c_void_p, p64, C.c_int64, pd, pd]
        L.orc_detect_numeric.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int64, pi, pi, p64,
                                         p64, pd, pd, pd, p64]
        L.orc_time_passes_numeric.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int64, p64]
        L.orc_solve_numeric.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int64, C.c_double, C.c_double,
                                        C.c_int, C.c_double, C.c_int, p64, p64, pd, pd, pd, p64, p64]
        L.orc_has_neg_exact.argtypes = [C.c_void_p, C.c_int64, C.c_int64, pi, p64]
        L.orc_potentials_exact.argtypes = [C.c_void_p, C.c_int64, C.c_int64, pi, p64]
        L.orc_find_zero_cycle_exact.argtypes = [C.c_void_p, C.c_int64, C.c_int64, pi, p64, p64, p64, p64]
        L.orc_stern_brocot.argtypes = [C.c_void_p, C.c_int64, C.c_int64, p64, p64, p64, p64, p64, p64, p64, p64,
                                       C.c_int64, p64, p64]
        _lib = L
    return _lib


def _p64(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def _pd(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


@dataclass
class NumericProbe:
    has_neg: int                 # 1 / 0 / -1 (undecided, only with pass_limit)
    cycle: list | None           # OPEN cycle or None (reference's cycle_data is None)
    sum_cost: float
    sum_time: float
    ratio: float
    passes: int


@dataclass
class NumericSolve:
    cycle: list                  # CLOSED
    sum_cost: float
    sum_time: float
    ratio: float
    iterations: int
    passes: int


@dataclass
class ExactSolve:
    cycle: list                  # CLOSED (as _solve_exact returns it)
    sum_cost: int
    sum_time: int
    ab: tuple                    # reduced (a, b)
    iterations: int
    probes: list                 # [(a, b), ...] handed to _has_negative_cycle_exact, in order
    passes: int

    @property
    def fraction(self) -> Fraction:
        return Fraction(int(self.sum_cost), int(self.sum_time))


class Oracle:
    """
    Arrays in insertion order -> reference preprocessing (solver.py:513-575) -> dst-sorted
    build (solver.py:450-511) -> L1 functions.  `all_int` mirrors `_all_int`
    (solver.py:289-298): pass True only if every cost/time is integral.
    """

    def __init__(self, n, U, V, Cw, Tw, all_int: bool | None = None, preprocess: bool = True, threads: int = 1):
        L = lib()
        U = np.ascontiguousarray(U, dtype=np.int64)
        V = np.ascontiguousarray(V, dtype=np.int64)
        Cf = np.ascontiguousarray(Cw, dtype=np.float64)
        Tf = np.ascontiguousarray(Tw, dtype=np.float64)
        if all_int is None:
            all_int = bool(np.all(Cf == np.floor(Cf)) and np.all(Tf == np.floor(Tf)))
        self.n = int(n)
        self.all_int = all_int
        self.kept = np.ones(len(U), dtype=np.uint8)
        if preprocess and len(U):
            L.orc_remove_dominated(len(U), _p64(U), _p64(V), _pd(Cf), _pd(Tf),
                                   self.kept.ctypes.data_as(C.POINTER(C.c_uint8)))
            k = self.kept.astype(bool)
            U, V, Cf, Tf = U[k].copy(), V[k].copy(), Cf[k].copy(), Tf[k].copy()
        self.m = len(U)
        self.ins = (U, V, Cf, Tf)
        Ci = Ti = None
        if all_int:
            Ci = Cf.astype(np.int64)   # int(edge.cost) truncation, solver.py:493-498
            Ti = Tf.astype(np.int64)
        self._keep = (U, V, Cf, Tf, Ci, Ti)
        self.g = L.orc_graph_build(self.n, self.m, _p64(U), _p64(V), _pd(Cf), _pd(Tf),
                                   _p64(Ci) if all_int else None, _p64(Ti) if all_int else None)
        if not self.g:
            raise MemoryError("orc_graph_build")
        L.orc_graph_set_threads(self.g, threads)

    def __del__(self):
        try:
            if getattr(self, "g", None):
                lib().orc_graph_free(self.g)
                self.g = None
        except Exception:
            pass

    def set_threads(self, t: int):
        lib().orc_graph_set_threads(self.g, int(t))

    def set_quirks(self, on: bool):
        """False = certifying checker: same arithmetic without the reference defects that flip verdicts
        (first-segment cumsum wrap :1179-1183, reduceat IndexError :839)."""
        lib().orc_graph_set_quirks(self.g, int(bool(on)))

    def arrays(self):
        """dst-sorted (_U,_V,_C,_T,_starts,_counts,order) as the reference stores them."""
        m, n = self.m, self.n
        U = np.empty(m, np.int64); V = np.empty(m, np.int64)
        Cf = np.empty(m, np.float64); Tf = np.empty(m, np.float64)
        st = np.empty(n, np.int64); ct = np.empty(n, np.int64); order = np.empty(m, np.int64)
        lib().orc_graph_arrays(self.g, _p64(U), _p64(V), _pd(Cf), _pd(Tf), _p64(st), _p64(ct), _p64(order))
        return U, V, Cf, Tf, st, ct, order

    def bounds(self):
        lo, hi = C.c_double(), C.c_double()
        lib().orc_numeric_bounds(self.g, C.byref(lo), C.byref(hi))
        return lo.value, hi.value

    def cycle_weights_float(self, open_cycle):
        cyc = np.asarray(open_cycle, dtype=np.int64)
        sc, st = C.c_double(), C.c_double()
        rc = lib().orc_cycle_weights_float_pub(self.g, _p64(cyc), len(cyc), C.byref(sc), C.byref(st))
        if rc:
            raise OracleError(rc, "cycle_weights_float")
        return sc.value, st.value

    # -- numeric -------------------------------------------------------
    def detect_negative_cycle_numeric(self, lam: float, slack: float = 0.0, use_kahan: bool = True,
                                      pass_limit: int = -1) -> NumericProbe:
        """solver.py:1124-1294"""
        hn, hc = C.c_int(), C.c_int()
        buf = np.empty(self.n + 1, np.int64)
        cl, ps = C.c_int64(), C.c_int64()
        sc, st, r = C.c_double(), C.c_double(), C.c_double()
        rc = lib().orc_detect_numeric(self.g, lam, slack, int(use_kahan), pass_limit, C.byref(hn), C.byref(hc),
                                      _p64(buf), C.byref(cl), C.byref(sc), C.byref(st), C.byref(r), C.byref(ps))
        if rc:
            raise OracleError(rc, "detect_numeric")
        cyc = [int(x) for x in buf[: cl.value]] if hc.value else None
        return NumericProbe(hn.value, cyc, sc.value, st.value, r.value, ps.value)

    def time_passes_numeric(self, lam: float, passes: int, slack: float = 1e-15, use_kahan: bool = True) -> int:
        done = C.c_int64()
        rc = lib().orc_time_passes_numeric(self.g, lam, slack, int(use_kahan), passes, C.byref(done))
        if rc:
            raise OracleError(rc, "time_passes")
        return done.value

    def solve_numeric(self, lambda_lo=None, lambda_hi=None, max_iter: int = 60, tol: float = 1e-12,
                      slack: float = 1e-15, use_kahan: bool = True, target_ratio=None,
                      early_term: bool = True) -> NumericSolve:
        """solver.py:1001-1122; raises OracleError(kind='no_cycle'|'convergence')."""
        if lambda_lo is None or lambda_hi is None:       # solver.py:1018-1025 (incl. the `or` quirk)
            lo, hi = self.bounds()
            lambda_lo = lambda_lo or lo
            lambda_hi = lambda_hi or hi
        buf = np.empty(self.n + 2, np.int64)
        cl, it, ps = C.c_int64(), C.c_int64(), C.c_int64()
        sc, st, r = C.c_double(), C.c_double(), C.c_double()
        tr = float("nan") if target_ratio is None else float(target_ratio)
        rc = lib().orc_solve_numeric(self.g, lambda_lo, lambda_hi, max_iter, tol, slack, int(use_kahan), tr,
                                     int(early_term), _p64(buf), C.byref(cl), C.byref(sc), C.byref(st),
                                     C.byref(r), C.byref(it), C.byref(ps))
        if rc:
            raise OracleError(rc, "solve_numeric")
        return NumericSolve([int(x) for x in buf[: cl.value]], sc.value, st.value, r.value, it.value, ps.value)

    # -- exact ---------------------------------------------------------
    def has_negative_cycle_exact(self, a: int, b: int) -> bool:
        """solver.py:811-856"""
        hn, ps = C.c_int(), C.c_int64()
        rc = lib().orc_has_neg_exact(self.g, a, b, C.byref(hn), C.byref(ps))
        if rc:
            raise OracleError(rc, "has_neg_exact")
        return bool(hn.value)

    def potentials_exact(self, a: int, b: int):
        """solver.py:922-969"""
        ok = C.c_int()
        dist = np.empty(self.n, np.int64)
        rc = lib().orc_potentials_exact(self.g, a, b, C.byref(ok), _p64(dist))
        if rc:
            raise OracleError(rc, "potentials_exact")
        return bool(ok.value), dist

    def find_zero_cycle_exact(self, a: int, b: int):
        """solver.py:858-920; returns (open_cycle, sum_c, sum_t) or None"""
        f = C.c_int()
        buf = np.empty(self.n + 1, np.int64)
        cl, sc, st = C.c_int64(), C.c_int64(), C.c_int64()
        rc = lib().orc_find_zero_cycle_exact(self.g, a, b, C.byref(f), _p64(buf), C.byref(cl), C.byref(sc),
                                             C.byref(st))
        if rc:
            raise OracleError(rc, "find_zero_cycle_exact")
        if not f.value:
            return None
        return [int(x) for x in buf[: cl.value]], sc.value, st.value

    def solve_exact(self, max_den=None, max_steps=None) -> ExactSolve:
        """_solve_exact + _stern_brocot_search, solver.py:685-797 (cycle closed as :713-714)."""
        if not self.all_int:
            raise OracleError(1, "solve_exact(requires integer weights)")
        buf = np.empty(self.n + 2, np.int64)
        cap = 4096
        probes = np.zeros(2 * cap, np.int64)
        cl, sc, st, a, b, it, npb, ps = (C.c_int64() for _ in range(8))
        rc = lib().orc_stern_brocot(self.g, -1 if max_den is None else max_den,
                                    -1 if max_steps is None else max_steps, _p64(buf), C.byref(cl), C.byref(sc),
                                    C.byref(st), C.byref(a), C.byref(b), C.byref(it), _p64(probes), cap,
                                    C.byref(npb), C.byref(ps))
        if rc:
            raise OracleError(rc, "stern_brocot")
        cyc = [int(x) for x in buf[: cl.value]]
        if cyc and (len(cyc) == 1 or cyc[0] != cyc[-1]):
            cyc.append(cyc[0])
        k = min(npb.value, cap)
        pr = [(int(probes[2 * i]), int(probes[2 * i + 1])) for i in range(k)]
        return ExactSolve(cyc, sc.value, st.value, (a.value, b.value), it.value, pr, ps.value)

    # -- solve() ladder ----------------------------------------------------
    def solve(self, **kw):
        """
        MinRatioCycleSolver.solve(), solver.py:581-679: numeric first (:627); on failure exact
        when all weights are integers (:630-632) else one retry at 10x tolerance (:634-638);
        then _validate_result's weight recomputation (:1523-1550).  Returns a dict shaped like
        the golden fixtures; raises OracleError when the reference raises.
        """
        tol = kw.pop("tol", 1e-12)
        try:
            r = self.solve_numeric(tol=tol, **kw)
            out = dict(cycle=r.cycle, sum_cost=r.sum_cost, sum_time=r.sum_time, ratio=r.ratio,
                       iterations=r.iterations, path="numeric")
        except OracleError as e:
            if e.kind not in ("no_cycle", "convergence", "edge_missing"):
                raise
            if self.all_int:
                x = self.solve_exact()
                out = dict(cycle=x.cycle, sum_cost=float(x.sum_cost), sum_time=float(x.sum_time),
                           ratio=float(x.sum_cost) / float(x.sum_time), iterations=x.iterations, path="exact")
            else:
                r = self.solve_numeric(tol=tol * 10, **kw)
                out = dict(cycle=r.cycle, sum_cost=r.sum_cost, sum_time=r.sum_time, ratio=r.ratio,
                           iterations=r.iterations, path="numeric_relaxed")
        sc, st = self.cycle_weights_float(out["cycle"][:-1])
        if abs(sc - out["sum_cost"]) > 1e-6 or abs(st - out["sum_time"]) > 1e-6 or abs(sc / st - out["ratio"]) > 1e-6:
            out.update(sum_cost=sc, sum_time=st, ratio=sc / st)
        return out

    # -- certificate -----------------------------------------
    def certify(self, closed_cycle, ratio: float, delta: float | None = None, rel: float = 1e-9):
        """
        C(r): every hop exists, recomputed ratio == r (isclose rel=abs=1e-9), and the
        reference detector finds no negative cycle at r - delta.
        """
        cyc = list(closed_cycle)
        if len(cyc) < 2 or cyc[0] != cyc[-1]:
            return False, "cycle not closed"
        try:
            sc, st = self.cycle_weights_float(cyc[:-1])
        except OracleError as e:
            return False, str(e)
        if st <= 0:
            return False, "non-positive time"
        if not math.isclose(sc / st, ratio, rel_tol=rel, abs_tol=rel):
            return False, f"ratio mismatch {sc / st} vs {ratio}"
        if delta is None:
            delta = 1e-7 * max(1.0, abs(ratio))
        pr = self.detect_negative_cycle_numeric(ratio - delta, 1e-15, True)
        if pr.has_neg != 0:
            return False, "negative cycle below ratio"
        return True, "ok"
