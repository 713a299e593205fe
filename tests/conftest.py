"""
Shared test plumbing.

Markers: `gpu` = needs a real B200 (parity tests proper, they call through the C-ABI).
Everything else runs on CPU.  The oracle (oracle/) is imported here and in the
tests only -- never by the product package.
"""

from __future__ import annotations

import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    config.addinivalue_line("markers", "slow: takes more than a few seconds on CPU")


def load_golden(name: str):
    with open(os.path.join(GOLDEN, name)) as fh:
        return json.load(fh)


def edges_to_arrays(edges):
    U = np.array([e[0] for e in edges], dtype=np.int64)
    V = np.array([e[1] for e in edges], dtype=np.int64)
    C = np.array([float(e[2]) for e in edges], dtype=np.float64)
    T = np.array([float(e[3]) for e in edges], dtype=np.float64)
    return U, V, C, T


def edges_all_int(edges) -> bool:
    """`_all_int` rule of solver.py:289-298 for JSON-decoded weights."""
    def ok(x):
        return isinstance(x, int) or (isinstance(x, float) and x.is_integer())
    return all(ok(e[2]) and ok(e[3]) for e in edges)


def rotations_equal(a, b) -> bool:
    """Closed cycles equal up to rotation."""
    a, b = list(a), list(b)
    if len(a) != len(b):
        return False
    if len(a) <= 1:
        return a == b
    oa, ob = a[:-1], b[:-1]
    if not oa:
        return True
    for s in range(len(ob)):
        if ob[s:] + ob[:s] == oa:
            return True
    return False


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle
    oracle.build()
    return oracle


def has_cuda() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
