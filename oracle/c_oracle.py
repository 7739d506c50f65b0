"""ctypes wrapper of oracle/liboracle.so (C restatement) — TEST / BASELINE INFRASTRUCTURE ONLY.
Used by tests (second checker) and by bench.py's cpu_baseline / `--impl reference` legs."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None
_D = C.POINTER(C.c_double)
_I = C.POINTER(C.c_int64)


def load():
    global _lib
    if _lib is None:
        so = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(so):
            subprocess.run(["make", "-C", HERE], check=True, capture_output=True)
        _lib = C.CDLL(so)
        _lib.oracle_first_touch.restype = C.c_int64
        _lib.oracle_pattern.restype = C.c_int64
        _lib.oracle_pcg.restype = C.c_int64
        _lib.oracle_apply_dirichlet.restype = C.c_double
        _lib.oracle_energy.restype = C.c_double
        _lib.oracle_free.argtypes = [C.c_void_p]
    return _lib


def _d(a):
    return a.ctypes.data_as(_D)


def _i(a):
    return a.ctypes.data_as(_I)


class CProblem:
    """Whole path on the CPU, timing each stage (seconds in self.t)."""

    def __init__(self, points, cells):
        lib = load()
        self.lib = lib
        self.points = np.ascontiguousarray(points, dtype=np.float64)
        self.cells = np.ascontiguousarray(cells, dtype=np.int64)
        self.ne, self.npc = self.cells.shape
        self.nn = self.points.shape[0]
        self.nb = 3 * self.npc
        self.t = {}
        t0 = time.perf_counter()
        self.node_first_dof = np.empty(self.nn, dtype=np.int64)
        self.cell_dofs = np.empty((self.ne, self.nb), dtype=np.int64)
        self.n = lib.oracle_first_touch(C.c_int64(self.ne), self.npc, _i(self.cells), C.c_int64(self.nn), _i(self.node_first_dof), _i(self.cell_dofs))
        self.colptr = np.empty(self.n + 1, dtype=np.int64)
        rv = C.c_void_p()
        self.nnz = lib.oracle_pattern(C.c_int64(self.ne), self.nb, _i(self.cell_dofs), C.c_int64(self.n), _i(self.colptr), C.byref(rv))
        self.rowval = np.ctypeslib.as_array(C.cast(rv, _I), shape=(self.nnz,)).copy()
        lib.oracle_free(rv)
        self.t["setup"] = time.perf_counter() - t0
        self.nzval = np.zeros(self.nnz)
        self.f = np.zeros(self.n)

    def _mat(self, lam_mu=None, simp=None, density=None):
        if simp is None:
            return 0, np.array([lam_mu[0], lam_mu[1], 0, 0], dtype=np.float64), np.zeros(1)
        return 1, np.array(simp, dtype=np.float64), np.ascontiguousarray(density, dtype=np.float64)

    def ke_batch(self, first0, count, **mat):
        mode, par, dens = self._mat(**mat)
        out = np.empty((count, self.nb, self.nb))
        rc = self.lib.oracle_ke_batch(C.c_int64(first0), C.c_int64(count), self.npc, _i(self.cells), _d(self.points), mode, _d(par), _d(dens), _d(out))
        assert rc == 0
        return out

    def assemble(self, **mat):
        mode, par, dens = self._mat(**mat)
        t0 = time.perf_counter()
        rc = self.lib.oracle_assemble(C.c_int64(self.ne), self.npc, _i(self.cells), _d(self.points), _i(self.cell_dofs), mode, _d(par), _d(dens),
                                      C.c_int64(self.n), _i(self.colptr), _i(self.rowval), _d(self.nzval))
        self.t["assemble"] = time.perf_counter() - t0
        if rc != 0:
            raise ValueError("det(J) is not positive")
        self.f[:] = 0.0

    def apply_force(self, nodes1, F):
        nodes1 = np.asarray(nodes1, dtype=np.int64)
        for g in nodes1:
            d0 = self.node_first_dof[g - 1]
            if d0 >= 0:
                self.f[d0:d0 + 3] += np.asarray(F, dtype=np.float64) / nodes1.size

    def apply_dirichlet(self, dofs0):
        flag = np.zeros(self.n, dtype=np.uint8)
        flag[np.asarray(dofs0, dtype=np.int64)] = 1
        return self.lib.oracle_apply_dirichlet(C.c_int64(self.n), _i(self.colptr), _i(self.rowval), _d(self.nzval), _d(self.f),
                                               flag.ctypes.data_as(C.POINTER(C.c_ubyte)))

    def pcg(self, tol=1e-8, itmax=10000, history=False):
        x = np.empty(self.n)
        solved = C.c_int()
        res = np.empty(itmax + 1) if history else None
        t0 = time.perf_counter()
        k = self.lib.oracle_pcg(C.c_int64(self.n), _i(self.colptr), _i(self.rowval), _d(self.nzval), _d(self.f), C.c_double(tol), C.c_double(tol),
                                C.c_int64(itmax), _d(x), C.byref(solved), _d(res) if history else None)
        self.t["pcg"] = time.perf_counter() - t0
        self.niter, self.solved = k, bool(solved.value)
        return x, k, bool(solved.value), (res[:k + 1] if history else None)

    def energy(self, u):
        t0 = time.perf_counter()
        e = self.lib.oracle_energy(C.c_int64(self.n), _i(self.colptr), _i(self.rowval), _d(self.nzval), _d(u))
        self.t["energy"] = time.perf_counter() - t0
        return e
