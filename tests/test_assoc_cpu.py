"""stemseg_b200/csrc/assoc.cuh compiled for the host: the assignment solver against scipy (the reference's
online_chainer.py:330) on tie-heavy matrices, and the CPython set-order emulation against the interpreter itself
(online_chainer.py:307-308 builds the label lists through `set`)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
from scipy.optimize import linear_sum_assignment

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("assoc") / "libassoc_host.so")
    src = os.path.join(ROOT, "tests", "native", "assoc_host.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, src], check=True)
    lib = ctypes.CDLL(out)
    lib.assoc_pyset_order.restype = ctypes.c_int
    lib.assoc_lsap.restype = ctypes.c_int
    return lib


def _order(lib, values):
    arr = (ctypes.c_longlong * max(1, len(values)))(*values)
    out = (ctypes.c_longlong * max(1, len(values)))()
    n = lib.assoc_pyset_order(arr, len(values), out)
    assert n >= 0
    return [out[i] for i in range(n)]


def test_pyset_order_matches_the_interpreter(lib):
    rng = np.random.default_rng(0)
    cases = [[], [-1], [3], [-1, 9, 2], [9, 2], list(range(1, 30)), [-1] + list(range(5, 200, 7))]
    for _ in range(3000):
        n = int(rng.integers(0, 70))
        hi = int(rng.choice([8, 24, 64, 200, 1000, 5000]))
        vals = sorted(set(rng.integers(1, hi + 1, size=n).tolist()))
        if rng.random() < 0.6:
            vals = [-1] + vals
        cases.append(vals)
    for vals in cases:
        expect = list(set(vals) - {-1})                      # what online_chainer.py:307-308 evaluates
        assert _order(lib, vals) == expect, vals


def _solve(lib, cost):
    nr, nc = cost.shape
    c = np.ascontiguousarray(cost, dtype=np.float64)
    rows = (ctypes.c_int * max(1, min(nr, nc)))()
    cols = (ctypes.c_int * max(1, min(nr, nc)))()
    n = lib.assoc_lsap(c.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), nr, nc, rows, cols)
    assert n == min(nr, nc)
    return [rows[i] for i in range(n)], [cols[i] for i in range(n)]


def test_lsap_matches_scipy_including_ties(lib):
    rng = np.random.default_rng(1)
    for trial in range(4000):
        nr, nc = int(rng.integers(1, 24)), int(rng.integers(1, 24))
        kind = trial % 4
        if kind == 0:        # IoU-like: mostly 1.0 (no overlap) with a few matches -> massive ties
            cost = np.ones((nr, nc), np.float32)
            for _ in range(int(rng.integers(0, min(nr, nc) + 1))):
                cost[rng.integers(0, nr), rng.integers(0, nc)] = np.float32(1.0 - rng.random())
        elif kind == 1:      # few distinct values
            cost = rng.choice(np.array([0.0, 0.25, 0.5, 1.0], np.float32), size=(nr, nc))
        elif kind == 2:      # all equal
            cost = np.full((nr, nc), 1.0, np.float32)
        else:
            cost = rng.random((nr, nc)).astype(np.float32)
        r_ref, c_ref = linear_sum_assignment(cost)
        r, c = _solve(lib, cost.astype(np.float64))
        assert r == r_ref.tolist() and c == c_ref.tolist(), (trial, cost)


def test_lsap_empty_sides(lib):
    rows = (ctypes.c_int * 1)()
    cols = (ctypes.c_int * 1)()
    assert lib.assoc_lsap(None, 0, 5, rows, cols) == 0
    assert lib.assoc_lsap(None, 4, 0, rows, cols) == 0
