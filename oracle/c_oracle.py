"""ctypes front end of oracle/fea_oracle_c.c (OpenMP C restatement of the reference path).

TEST INFRASTRUCTURE ONLY: the checker and the multi-core CPU baseline of bench.py, never the
product.  The shared object is built next to the source on first use (gcc -O3 -fopenmp) and by
`__graft_entry__.build()`; it is git-ignored and travels to the GPU box with the snapshot.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "fea_oracle_c.c")
LIB = os.path.join(HERE, "libfea_oracle_c.so")
_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.isfile(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        # no -march=native: the object is built in one container and may run on another host
        cmd = ["gcc", "-O3", "-fopenmp", "-shared", "-fPIC", SRC, "-o", LIB, "-lm"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("gcc failed:\n" + " ".join(cmd) + "\n" + res.stderr)
    return LIB


def load():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build())
        P, i64, f64 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_double
        lib.fea_c_threads.restype = ctypes.c_int
        lib.fea_c_hex8_ke.restype = ctypes.c_int
        lib.fea_c_hex8_ke.argtypes = [P, f64, f64, P]
        lib.fea_c_hex8_ke_batch.restype = i64
        lib.fea_c_hex8_ke_batch.argtypes = [P, P, i64, f64, f64, P]
        lib.fea_c_assemble_hex8.restype = i64
        lib.fea_c_assemble_hex8.argtypes = [P, P, i64, f64, f64, P, P, P]
        lib.fea_c_spmv.restype = None
        lib.fea_c_spmv.argtypes = [i64, P, P, P, P, P]
        lib.fea_c_jacobi_pcg.restype = i64
        lib.fea_c_jacobi_pcg.argtypes = [i64, P, P, P, P, P, f64, i64, P]
        _lib = lib
    return _lib


def threads() -> int:
    return int(load().fea_c_threads())


def _ptr(a: np.ndarray) -> int:
    return a.ctypes.data


def hex8_ke_batched(nodes, elements, E: float, nu: float) -> np.ndarray:
    """(M, 24, 24) Ke in the reference's dense order of operations (utils.py:127-239), all cores."""
    nodes = np.ascontiguousarray(nodes, dtype=np.float64)
    elements = np.ascontiguousarray(elements, dtype=np.int64)
    M = elements.shape[0]
    ke = np.empty((M, 24, 24))
    bad = load().fea_c_hex8_ke_batch(_ptr(nodes), _ptr(elements), M, float(E), float(nu), _ptr(ke))
    if bad:
        raise ValueError("Jacobian determinant is non-positive. Check the element shape.")
    return ke


def dof_pattern(elements, n_nodes: int, d: int):
    """Structural DOF-level CSR pattern (int32 indptr, sorted int32 indices) of the assembled
    matrix: the node-level pattern (scipy coo -> csr on node pairs) expanded d x d.  Equals
    `fea_oracle.structural_pattern` (checked in tests/test_oracle_golden.py) at 1/d^2 of the cost."""
    elements = np.asarray(elements, dtype=np.int64)
    M, npe = elements.shape
    rows = np.repeat(elements, npe, axis=1).ravel()
    cols = np.tile(elements, (1, npe)).ravel()
    A = sp.coo_matrix((np.ones(rows.size, dtype=np.int8), (rows, cols)), shape=(n_nodes, n_nodes)).tocsr()
    A.sum_duplicates()
    A.sort_indices()
    cnt = np.diff(A.indptr).astype(np.int64)
    seg = (d * A.indices.astype(np.int64)[:, None] + np.arange(d)).ravel()  # one DOF row per node, concatenated
    seg_start = d * A.indptr[:-1].astype(np.int64)
    row_len = np.repeat(d * cnt, d)            # length of every DOF row
    row_src = np.repeat(seg_start, d)          # where its columns start in `seg`
    indptr = np.zeros(d * n_nodes + 1, dtype=np.int64)
    np.cumsum(row_len, out=indptr[1:])
    pos = np.arange(indptr[-1]) - np.repeat(indptr[:-1], row_len) + np.repeat(row_src, row_len)
    return indptr.astype(np.int32), seg[pos].astype(np.int32)


def assemble_hex8(nodes, elements, E: float, nu: float, pattern=None) -> sp.csr_matrix:
    """Ke + `K[ix_(d,d)] += Ke` (cubebeam.py:82-90) into the structural CSR pattern, all cores."""
    nodes = np.ascontiguousarray(nodes, dtype=np.float64)
    elements = np.ascontiguousarray(elements, dtype=np.int64)
    n = 3 * nodes.shape[0]
    indptr, indices = pattern if pattern is not None else dof_pattern(elements, nodes.shape[0], 3)
    data = np.zeros(indices.shape[0])
    bad = load().fea_c_assemble_hex8(_ptr(nodes), _ptr(elements), elements.shape[0], float(E), float(nu),
                                     _ptr(indptr), _ptr(indices), _ptr(data))
    if bad:
        raise ValueError("Jacobian determinant is non-positive. Check the element shape.")
    return sp.csr_matrix((data, indices, indptr), shape=(n, n))


def spmv(A: sp.csr_matrix, x: np.ndarray) -> np.ndarray:
    indptr = np.ascontiguousarray(A.indptr, dtype=np.int32)
    indices = np.ascontiguousarray(A.indices, dtype=np.int32)
    data = np.ascontiguousarray(A.data, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty(A.shape[0])
    load().fea_c_spmv(A.shape[0], _ptr(indptr), _ptr(indices), _ptr(data), _ptr(x), _ptr(y))
    return y


def jacobi_pcg(A: sp.csr_matrix, b: np.ndarray, tol: float = 1e-12, maxiter: int | None = None):
    """Same recurrence and stop rule as fea_oracle.jacobi_pcg.  Returns (x, iterations, relres)."""
    n = A.shape[0]
    indptr = np.ascontiguousarray(A.indptr, dtype=np.int32)
    indices = np.ascontiguousarray(A.indices, dtype=np.int32)
    data = np.ascontiguousarray(A.data, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.empty(n)
    rel = ctypes.c_double(0.0)
    it = load().fea_c_jacobi_pcg(n, _ptr(indptr), _ptr(indices), _ptr(data), _ptr(b), _ptr(x), float(tol),
                                 int(10 * n if maxiter is None else maxiter), ctypes.addressof(rel))
    return x, int(it), float(rel.value)
