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
        lib.fea_c_set_threads.restype = None
        lib.fea_c_set_threads.argtypes = [ctypes.c_int]
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
        lib.fea_c_reduce_csr.restype = None
        lib.fea_c_reduce_csr.argtypes = [i64, P, P, P, P, P, P, P]
        lib.fea_c_expand_pattern.restype = None
        lib.fea_c_expand_pattern.argtypes = [i64, ctypes.c_int32, P, P, P, P]
        _lib = lib
    return _lib


def threads() -> int:
    return int(load().fea_c_threads())


def set_threads(n: int) -> int:
    """Pin the OpenMP team size (whatever OMP_NUM_THREADS the launcher exported); returns it."""
    load().fea_c_set_threads(int(n))
    return threads()


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
    if d * d * A.nnz >= 2**31:
        raise ValueError("pattern does not fit int32")
    indptr_n = np.ascontiguousarray(A.indptr, dtype=np.int32)
    indices_n = np.ascontiguousarray(A.indices, dtype=np.int32)
    indptr = np.empty(d * n_nodes + 1, dtype=np.int32)
    indices = np.empty(d * d * A.nnz, dtype=np.int32)
    load().fea_c_expand_pattern(n_nodes, d, _ptr(indptr_n), _ptr(indices_n), _ptr(indptr), _ptr(indices))
    return indptr, indices


def reduce_csr(K: sp.csr_matrix, free: np.ndarray) -> sp.csr_matrix:
    """`K[np.ix_(free, free)]` (cubebeam.py:92-96) without scipy's fancy-indexing temporaries:
    two parallel passes over the rows (count, fill).  `free` ascending."""
    n = K.shape[0]
    free = np.asarray(free, dtype=np.int64)
    dof_map = np.full(n, -1, dtype=np.int32)
    dof_map[free] = np.arange(free.size, dtype=np.int32)
    indptr = np.ascontiguousarray(K.indptr, dtype=np.int32)
    indices = np.ascontiguousarray(K.indices, dtype=np.int32)
    data = np.ascontiguousarray(K.data, dtype=np.float64)
    out_indptr = np.zeros(free.size + 1, dtype=np.int64)
    lib = load()
    lib.fea_c_reduce_csr(n, _ptr(indptr), _ptr(indices), _ptr(data), _ptr(dof_map), _ptr(out_indptr), None, None)
    np.cumsum(out_indptr, out=out_indptr)
    nnz = int(out_indptr[-1])
    out_indices = np.empty(nnz, dtype=np.int32)
    out_data = np.empty(nnz, dtype=np.float64)
    lib.fea_c_reduce_csr(n, _ptr(indptr), _ptr(indices), _ptr(data), _ptr(dof_map), _ptr(out_indptr),
                         _ptr(out_indices), _ptr(out_data))
    return sp.csr_matrix((out_data, out_indices, out_indptr.astype(np.int32)), shape=(free.size, free.size))


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
