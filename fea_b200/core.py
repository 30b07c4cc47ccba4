"""Host-side plumbing over the C ABI: device buffers (torch), pattern / matrix objects, PCG.

PyTorch is used for device memory, streams and (in dist.py) torch.distributed only; every
computation is a hand-written sm_100a kernel behind include/fea_b200.h.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib


def device() -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.FeaLibraryError("fea_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t) -> int | None:
    return None if t is None else t.data_ptr()


def to_device(a, dtype: torch.dtype) -> torch.Tensor:
    """numpy / torch (host or device) -> contiguous CUDA tensor of `dtype` (conversion on device)."""
    dev = device()
    if isinstance(a, torch.Tensor):
        t = a
    else:
        arr = np.asarray(a)
        if not arr.flags.c_contiguous:
            arr = np.ascontiguousarray(arr)
        if not arr.flags.writeable:
            arr = arr.copy()
        t = torch.from_numpy(arr)
    if t.device != dev:
        t = t.to(dev, non_blocking=True)
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def _status_slot() -> torch.Tensor:
    return torch.zeros(2, dtype=torch.int32, device=device())


def _check_status(status: torch.Tensor) -> None:
    _lib.raise_for_status(status.cpu().numpy())


# ------------------------------------------------------------------------------------------------
# symbolic pass
# ------------------------------------------------------------------------------------------------
@dataclass
class Pattern:
    """Node-block CSR pattern (device).  DOF-level CSR of `d` DOF per node is implied:
    row d*i+a = {d*j+b : j in node_colidx[node_rowptr[i]:node_rowptr[i+1]], b < d}."""

    n_nodes: int
    nodes_per_elem: int
    nnz_blocks: int
    max_coupled: int
    max_incident: int
    n2e_ptr: torch.Tensor
    n2e: torch.Tensor
    node_rowptr: torch.Tensor
    node_colidx: torch.Tensor
    _csr: dict = field(default_factory=dict)

    def csr(self, d: int):
        """(rowptr, colidx) int32 device tensors of the DOF-level CSR pattern."""
        if d not in self._csr:
            nnz = d * d * self.nnz_blocks
            if nnz >= 2**31:
                raise ValueError("nnz does not fit int32")
            rowptr = torch.empty(self.n_nodes * d + 1, dtype=torch.int32, device=self.node_rowptr.device)
            colidx = torch.empty(nnz, dtype=torch.int32, device=self.node_rowptr.device)
            lib = _lib.load()
            _lib.check(lib.fea_csr_expand(self.n_nodes, d, _p(self.node_rowptr), _p(self.node_colidx), _p(rowptr),
                                          _p(colidx), _stream()), "fea_csr_expand")
            self._csr[d] = (rowptr, colidx)
        return self._csr[d]


def symbolic(elements: torch.Tensor, n_nodes: int) -> Pattern:
    """connectivity (M, npe) int32 on device -> Pattern (replaces the index bookkeeping of
    cubebeam.py:80-90)."""
    lib = _lib.load()
    dev = device()
    assert elements.dtype == torch.int32 and elements.is_cuda and elements.is_contiguous()
    n_elem, npe = elements.shape
    ws_bytes = lib.fea_csr_symbolic_workspace(n_nodes, n_elem, npe)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    n2e_ptr = torch.empty(n_nodes + 1, dtype=torch.int32, device=dev)
    n2e = torch.empty(max(n_elem * npe, 1), dtype=torch.int32, device=dev)
    node_rowptr = torch.empty(n_nodes + 1, dtype=torch.int32, device=dev)
    sizes = (ctypes.c_int64 * 4)()
    _lib.check(lib.fea_csr_symbolic_count(_p(elements), n_elem, npe, n_nodes, _p(n2e_ptr), _p(n2e), _p(node_rowptr),
                                          ctypes.addressof(sizes), _p(ws), ws_bytes, _stream()),
               "fea_csr_symbolic_count")
    nnzb, max_coupled, max_incident = int(sizes[0]), int(sizes[1]), int(sizes[2])
    node_colidx = torch.empty(max(nnzb, 1), dtype=torch.int32, device=dev)
    _lib.check(lib.fea_csr_symbolic_fill(_p(elements), n_elem, npe, n_nodes, _p(n2e_ptr), _p(n2e), _p(node_rowptr),
                                         _p(node_colidx), max_incident, _stream()), "fea_csr_symbolic_fill")
    return Pattern(n_nodes, npe, nnzb, max_coupled, max_incident, n2e_ptr, n2e, node_rowptr, node_colidx)


# ------------------------------------------------------------------------------------------------
# matrix object
# ------------------------------------------------------------------------------------------------
@dataclass
class BlockCSR:
    """Assembled K on the device: `values` in DOF-level CSR order over `pattern`."""

    pattern: Pattern
    dof_per_node: int
    values: torch.Tensor
    dinv: torch.Tensor | None = None   # Jacobi 1/K_ii, 0 on constrained DOF
    fixed: torch.Tensor | None = None  # uint8 per DOF

    @property
    def n_dof(self) -> int:
        return self.pattern.n_nodes * self.dof_per_node

    @property
    def nnz(self) -> int:
        return self.dof_per_node**2 * self.pattern.nnz_blocks

    def matvec(self, x: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """y = K x (cubebeam.py:106)."""
        lib = _lib.load()
        x = x.contiguous()
        y = torch.empty(self.n_dof, dtype=torch.float64, device=x.device) if out is None else out
        pt = self.pattern
        _lib.check(lib.fea_spmv(pt.n_nodes, self.dof_per_node, _p(pt.node_rowptr), _p(pt.node_colidx),
                                _p(self.values), pt.max_coupled, _p(x), _p(y), _stream()), "fea_spmv")
        return y

    def matmat(self, X: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """Y = K X, X (n_dof, n_rhs) row-major."""
        lib = _lib.load()
        X = X.contiguous()
        Y = torch.empty_like(X) if out is None else out
        pt = self.pattern
        _lib.check(lib.fea_spmm(pt.n_nodes, self.dof_per_node, _p(pt.node_rowptr), _p(pt.node_colidx),
                                _p(self.values), _p(X), _p(Y), X.shape[1], _stream()), "fea_spmm")
        return Y

    def to_scipy(self):
        """Host copy as scipy.sparse.csr_matrix (for inspection and parity tests)."""
        import scipy.sparse as sp

        rowptr, colidx = self.pattern.csr(self.dof_per_node)
        return sp.csr_matrix(
            (self.values.cpu().numpy(), colidx.cpu().numpy(), rowptr.cpu().numpy()), shape=(self.n_dof, self.n_dof)
        )


def _fixed_mask(constraints, n_dof: int) -> torch.Tensor | None:
    if constraints is None:
        return None
    if isinstance(constraints, torch.Tensor):
        c = constraints.to(device()).reshape(-1)
        fixed = (c != 0).to(torch.uint8)
    else:
        fixed = to_device((np.asarray(constraints).reshape(-1) != 0).astype(np.uint8), torch.uint8)
    if fixed.numel() != n_dof:
        raise ValueError("constraints must have one entry per DOF")
    return fixed.contiguous()


def assemble_hex8(nodes: torch.Tensor, elements: torch.Tensor, E: float, nu: float, pattern: Pattern | None = None,
                  fixed: torch.Tensor | None = None, mode: int = _lib.ASSEMBLE_FULL, check: bool = True) -> BlockCSR:
    """Ke evaluation + assembly (utils.py:127-239 + cubebeam.py:80-90), fused; K never dense."""
    lib = _lib.load()
    n_nodes = nodes.shape[0]
    if pattern is None:
        pattern = symbolic(elements, n_nodes)
    values = torch.empty(9 * pattern.nnz_blocks, dtype=torch.float64, device=nodes.device)
    dinv = torch.empty(3 * n_nodes, dtype=torch.float64, device=nodes.device)
    status = _status_slot()
    _lib.check(lib.fea_assemble_hex8(_p(nodes), _p(elements), elements.shape[0], n_nodes, float(E), float(nu),
                                     _p(pattern.n2e_ptr), _p(pattern.n2e), _p(pattern.node_rowptr),
                                     _p(pattern.node_colidx), max(pattern.max_coupled, 1), _p(fixed), mode,
                                     _p(values), _p(dinv), _p(status), _stream()), "fea_assemble_hex8")
    if check:
        _check_status(status)
    return BlockCSR(pattern, 3, values, dinv, fixed)


def assemble_beam(EI: torch.Tensor, length: torch.Tensor, elements: torch.Tensor, n_nodes: int,
                  pattern: Pattern | None = None, fixed: torch.Tensor | None = None,
                  mode: int = _lib.ASSEMBLE_FULL) -> BlockCSR:
    """euler_bernoulli.py:22-49 for per-element EI, L."""
    lib = _lib.load()
    if pattern is None:
        pattern = symbolic(elements, n_nodes)
    values = torch.empty(4 * pattern.nnz_blocks, dtype=torch.float64, device=EI.device)
    dinv = torch.empty(2 * n_nodes, dtype=torch.float64, device=EI.device)
    _lib.check(lib.fea_assemble_beam(_p(EI), _p(length), _p(elements), elements.shape[0], n_nodes,
                                     _p(pattern.n2e_ptr), _p(pattern.n2e), _p(pattern.node_rowptr),
                                     _p(pattern.node_colidx), _p(fixed), mode, _p(values), _p(dinv), _stream()),
               "fea_assemble_beam")
    return BlockCSR(pattern, 2, values, dinv, fixed)


def assemble_truss(nodes: torch.Tensor, members: torch.Tensor, k: torch.Tensor, pattern: Pattern | None = None,
                   fixed: torch.Tensor | None = None, mode: int = _lib.ASSEMBLE_FULL, check: bool = True) -> BlockCSR:
    """Linearised pin-jointed members (tangent of truss.py:78-92), 3 DOF per node."""
    lib = _lib.load()
    n_nodes = nodes.shape[0]
    if pattern is None:
        pattern = symbolic(members, n_nodes)
    values = torch.empty(9 * pattern.nnz_blocks, dtype=torch.float64, device=nodes.device)
    dinv = torch.empty(3 * n_nodes, dtype=torch.float64, device=nodes.device)
    status = _status_slot()
    _lib.check(lib.fea_assemble_truss(_p(nodes), _p(members), _p(k), members.shape[0], n_nodes, _p(pattern.n2e_ptr),
                                      _p(pattern.n2e), _p(pattern.node_rowptr), _p(pattern.node_colidx), _p(fixed),
                                      mode, _p(values), _p(dinv), _p(status), _stream()), "fea_assemble_truss")
    if check:
        _check_status(status)
    return BlockCSR(pattern, 3, values, dinv, fixed)


# ------------------------------------------------------------------------------------------------
# solver
# ------------------------------------------------------------------------------------------------
@dataclass
class SolveInfo:
    iterations: int
    rel_residual: float
    bnorm: float
    status: int
    history: np.ndarray | None = None


def pcg(A: BlockCSR, b: torch.Tensor, tol: float = 1e-12, max_iter: int | None = None, history: bool = False,
        raise_on_failure: bool = True):
    """Jacobi-PCG on the free DOF of A (those with dinv != 0); x0 = 0; recurrence-residual stop
    ||r|| <= tol ||b||.  Replaces np.linalg.solve(reduced_K, reduced_forces) (cubebeam.py:98)."""
    lib = _lib.load()
    n = A.n_dof
    if A.dinv is None:
        raise ValueError("matrix has no Jacobi diagonal")
    if max_iter is None:
        max_iter = min(10 * n, 2**31 - 1)
    b = b.contiguous()
    x = torch.empty(n, dtype=torch.float64, device=b.device)
    ws_bytes = lib.fea_pcg_workspace(n)
    work = torch.empty(ws_bytes, dtype=torch.uint8, device=b.device)
    hist = torch.zeros(max_iter, dtype=torch.float64, device=b.device) if history else None
    res = _lib.PcgResult()
    pt = A.pattern
    _lib.check(lib.fea_pcg_solve(pt.n_nodes, A.dof_per_node, _p(pt.node_rowptr), _p(pt.node_colidx), _p(A.values),
                                 pt.max_coupled,
                                 _p(A.dinv), _p(b), _p(x), float(tol), int(max_iter), _p(work), ws_bytes, _p(hist),
                                 ctypes.byref(res), _stream()), "fea_pcg_solve")
    info = SolveInfo(res.iterations, res.rel_residual, res.bnorm, res.status,
                     hist[: res.iterations].cpu().numpy() if history else None)
    if raise_on_failure and res.status != _lib.FEA_OK:
        _lib.raise_for_status(np.array([res.status, 0x7FFFFFFF - res.iterations]))
    return x, info


def chain_solve(A: BlockCSR, b: torch.Tensor, extended: bool = True):
    """Direct solve of a chain mesh (block-tridiagonal K, 1 or 2 DOF per node) by parallel cyclic
    reduction; constrained DOF (A.fixed) come back exactly 0.  Replaces np.linalg.solve of
    euler_bernoulli.py:69.  `extended`: eliminate in double-double arithmetic (default; the beam's
    cond(K) ~ 5 n^4 leaves FP64 elimination nothing beyond n of a few thousand).  ValueError if the
    pattern is not a chain, LinAlgError for a singular pivot."""
    lib = _lib.load()
    pt = A.pattern
    n = pt.n_nodes
    b = b.contiguous()
    x = torch.empty(A.n_dof, dtype=torch.float64, device=b.device)
    ext = 1 if extended else 0
    ws_bytes = lib.fea_chain_solve_workspace(n, A.dof_per_node, ext)
    work = torch.empty(ws_bytes, dtype=torch.uint8, device=b.device)
    status = _status_slot()
    _lib.check(lib.fea_chain_solve(n, A.dof_per_node, _p(pt.node_rowptr), _p(pt.node_colidx), _p(A.values),
                                   _p(A.fixed), _p(b), _p(x), ext, _p(work), ws_bytes, _p(status), _stream()),
               "fea_chain_solve")
    _check_status(status)
    free = A.dinv != 0 if A.dinv is not None else None
    r = b - A.matvec(x)
    if free is not None:
        r = r[free]
        bn = b[free].norm()
    else:
        bn = b.norm()
    rel = float(r.norm() / bn) if float(bn) > 0 else 0.0
    return x, SolveInfo(0, rel, float(bn), _lib.FEA_OK, None)


def pcg_multi(A: BlockCSR, B: torch.Tensor, tol: float = 1e-12, max_iter: int | None = None,
              raise_on_failure: bool = True):
    """Batched multi-RHS Jacobi-PCG (BASELINE config 5): B, X are (n_dof, n_rhs) row-major."""
    lib = _lib.load()
    n, k = B.shape
    if n != A.n_dof:
        raise ValueError("B has the wrong number of rows")
    if max_iter is None:
        max_iter = min(10 * n, 2**31 - 1)
    B = B.contiguous()
    X = torch.empty_like(B)
    ws_bytes = lib.fea_pcg_multi_workspace(n, k)
    work = torch.empty(ws_bytes, dtype=torch.uint8, device=B.device)
    iters = (ctypes.c_int32 * k)()
    res = _lib.PcgResult()
    pt = A.pattern
    _lib.check(lib.fea_pcg_solve_multi(pt.n_nodes, A.dof_per_node, _p(pt.node_rowptr), _p(pt.node_colidx),
                                       _p(A.values), _p(A.dinv), _p(B), _p(X), k, float(tol), int(max_iter),
                                       _p(work), ws_bytes, ctypes.addressof(iters), ctypes.byref(res), _stream()),
               "fea_pcg_solve_multi")
    info = SolveInfo(res.iterations, res.rel_residual, res.bnorm, res.status, np.array(iters[:], dtype=np.int64))
    if raise_on_failure and res.status != _lib.FEA_OK:
        _lib.raise_for_status(np.array([res.status, 0x7FFFFFFF - res.iterations]))
    return X, info


def solve_system(A: BlockCSR, loads: torch.Tensor, tol: float = 1e-12, max_iter: int | None = None,
                 history: bool = False):
    """Reduce + solve + expand + reactions (cubebeam.py:92-108) on an assembled FULL matrix whose
    dinv encodes the constraints.  Returns (u, reactions = K_full u, info)."""
    u, info = pcg(A, loads.reshape(-1), tol=tol, max_iter=max_iter, history=history)
    reactions = A.matvec(u)
    return u, reactions, info
