"""Three-way parity verdict of SURVEY.md 8c, shared by the GPU parity tests."""

from __future__ import annotations

import math

MATCH, REF_SUBOPTIMAL, MISMATCH = "MATCH", "REF_SUBOPTIMAL", "MISMATCH"


def three_way(oracle_obj, gpu, ref, rel: float = 1e-9):
    """
    gpu / ref: dicts with closed `cycle` and float `ratio`.  oracle_obj: oracle.Oracle of the same graph
    with quirks OFF (certifying checker).
    MATCH           |r_gpu - r_ref| <= 1e-9 * max(1, |r_ref|)   (the reference's isclose(rel=abs=1e-9))
    REF_SUBOPTIMAL  r_gpu < r_ref, C(r_gpu) holds and C(r_ref) fails
    MISMATCH        anything else
    """
    rg, rr = gpu["ratio"], ref["ratio"]
    if abs(rg - rr) <= rel * max(1.0, abs(rr)):
        return MATCH, "within tolerance"
    ok_g, why_g = oracle_obj.certify(gpu["cycle"], rg)
    ok_r, why_r = oracle_obj.certify(ref["cycle"], rr)
    if rg < rr and ok_g and not ok_r:
        return REF_SUBOPTIMAL, f"ref certificate fails: {why_r}"
    return MISMATCH, f"gpu={rg} ({why_g}) ref={rr} ({why_r})"


def cycle_is_valid(closed_cycle, U, V) -> bool:
    """Every hop of the closed cycle is an edge of the graph."""
    if len(closed_cycle) < 2 or closed_cycle[0] != closed_cycle[-1]:
        return False
    edges = set(zip(U.tolist(), V.tolist()))
    return all((a, b) in edges for a, b in zip(closed_cycle[:-1], closed_cycle[1:]))


def isclose9(a, b):
    return math.isclose(a, b, rel_tol=1e-9, abs_tol=1e-9)
