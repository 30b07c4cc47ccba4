"""Counterpart of the reference's euler_bernoulli.py: Hermite beam elements, assembly, solve and
the script's own moment/shear post-processing, all on the device.

The reference computes everything at import time and exposes the results as module globals
(euler_bernoulli.py:5-102).  Here the same names (`E, I, L, q, n_elements, n_nodes, element_length,
element_stiffness_matrix, global_stiffness_matrix, load_vector, fixed_dofs, free_dofs,
displacement_vector, moment_vector, shear_vector`) are computed lazily on first access, so that
importing the module does not run a solve; `run()` is the script body, `solve_beam` the callable
for arbitrary sizes (BASELINE config 2: cantilever, 100k elements, tip load).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, core

# Beam properties (euler_bernoulli.py:5-15)
E = 210e9
I = 1e-6
L = 1.0
q = 1000
n_elements = 100
n_nodes = n_elements + 1
element_length = L / n_elements


def beam_stiffness_matrices(EI, lengths) -> torch.Tensor:
    """(M, 4, 4) Hermite stiffness per element, DOF (w1, th1, w2, th2) (euler_bernoulli.py:22-39)."""
    lib = _lib.load()
    EI_d = core.to_device(np.atleast_1d(EI) if not isinstance(EI, torch.Tensor) else EI, torch.float64)
    L_d = core.to_device(np.atleast_1d(lengths) if not isinstance(lengths, torch.Tensor) else lengths, torch.float64)
    m = EI_d.numel()
    ke = torch.empty((m, 4, 4), dtype=torch.float64, device=EI_d.device)
    _lib.check(lib.fea_ke_beam(EI_d.data_ptr(), L_d.data_ptr(), m, ke.data_ptr(), core._stream()), "fea_ke_beam")
    return ke


def beam_elements(n: int) -> np.ndarray:
    """Connectivity [i, i+1]; DOF [2i, 2i+1, 2i+2, 2i+3] (euler_bernoulli.py:44)."""
    i = np.arange(n, dtype=np.int64)
    return np.stack([i, i + 1], axis=1)


def uniform_load_vector(q_load: float, n: int, le: float) -> np.ndarray:
    """Consistent load vector as the reference accumulates it, element by element, with its
    L/6 moment arm (euler_bernoulli.py:51-57)."""
    fe = q_load * le / 2 * np.array([1, le / 6, 1, -le / 6])
    f = np.zeros(2 * (n + 1))
    for c in range(4):  # same order of additions per DOF as the reference's loop
        np.add.at(f, 2 * np.arange(n) + c, fe[c])
    return f


def solve_beam(elements, EI, lengths, constraints, loads, tol: float = 1e-12, max_iter: int | None = None,
               return_matrix: bool = False, method: str = "direct", extended: bool = True):
    """Assemble (euler_bernoulli.py:42-49), eliminate constrained DOF (:61-66), solve (:69),
    expand (:72-73).  elements (M,2); EI, lengths (M,); constraints, loads (n_nodes, 2).
    Returns u (n_nodes, 2) [, BlockCSR, SolveInfo].

    method "direct" (default): block-tridiagonal parallel cyclic reduction on the device
    (fea_chain_solve) -- the counterpart of the reference's dense LU (`np.linalg.solve`,
    euler_bernoulli.py:69), eliminating in double-double arithmetic unless extended=False; it needs a
    chain mesh (element i joins nodes i and i+1 in any order) and falls back to "pcg" otherwise.  method "pcg": Jacobi-PCG, usable while cond(K) ~ 5 n^4 stays far
    below 1/eps (n of a few hundred)."""
    elements_d = core.to_device(elements, torch.int32)
    EI_d = core.to_device(EI, torch.float64)
    L_d = core.to_device(lengths, torch.float64)
    n_nodes_ = int(np.shape(constraints)[0])
    fixed = core._fixed_mask(constraints, 2 * n_nodes_)
    b = core.to_device(loads, torch.float64).reshape(-1)
    K = core.assemble_beam(EI_d, L_d, elements_d, n_nodes_, fixed=fixed)
    if method == "direct" and K.pattern.max_coupled <= 3:
        try:
            u, info = core.chain_solve(K, b, extended=extended)
        except ValueError:  # three couplings per node, but not a chain
            u, info = core.pcg(K, b, tol=tol, max_iter=max_iter)
    else:
        u, info = core.pcg(K, b, tol=tol, max_iter=max_iter)
    u_host = u.cpu().numpy().reshape(n_nodes_, 2)
    if return_matrix:
        return u_host, K, info
    return u_host


def moment_shear(displacements, EI, lengths):
    """moment_vector, shear_vector exactly as the reference defines them (euler_bernoulli.py:76-102):
    n_nodes entries each, the last left at 0."""
    lib = _lib.load()
    u = core.to_device(np.asarray(displacements, dtype=np.float64).reshape(-1), torch.float64)
    EI_d = core.to_device(EI, torch.float64)
    L_d = core.to_device(lengths, torch.float64)
    n = EI_d.numel()
    m = torch.empty(n + 1, dtype=torch.float64, device=u.device)
    v = torch.empty(n + 1, dtype=torch.float64, device=u.device)
    _lib.check(lib.fea_beam_moment_shear(u.data_ptr(), EI_d.data_ptr(), L_d.data_ptr(), n, m.data_ptr(),
                                         v.data_ptr(), core._stream()), "fea_beam_moment_shear")
    return m.cpu().numpy(), v.cpu().numpy()


def fixed_fixed_case(n: int = n_elements, E_: float = E, I_: float = I, L_: float = L, q_: float = q):
    """The shipped problem: both ends clamped, uniform load (euler_bernoulli.py:5-19, 51-61)."""
    le = L_ / n
    constraints = np.zeros((n + 1, 2), dtype=int)
    constraints[0] = 1
    constraints[-1] = 1
    return beam_elements(n), np.full(n, E_ * I_), np.full(n, le), constraints, uniform_load_vector(q_, n, le).reshape(-1, 2)


def cantilever_case(n: int, E_: float = E, I_: float = I, L_: float = L, P: float = -1000.0):
    """BASELINE config 2: DOF {0, 1} fixed, tip load P on DOF 2n."""
    le = L_ / n
    constraints = np.zeros((n + 1, 2), dtype=int)
    constraints[0] = 1
    loads = np.zeros((n + 1, 2))
    loads[-1, 0] = P
    return beam_elements(n), np.full(n, E_ * I_), np.full(n, le), constraints, loads


_RESULTS: dict | None = None


def run() -> dict:
    """The reference's script body (euler_bernoulli.py:17-102) on the device."""
    global _RESULTS
    elements, EI, Ls, constraints, loads = fixed_fixed_case()
    u, K, info = solve_beam(elements, EI, Ls, constraints, loads, return_matrix=True)
    fixed = [0, 1, 2 * n_nodes - 2, 2 * n_nodes - 1]
    m, v = moment_shear(u, EI, Ls)
    _RESULTS = {
        "element_stiffness_matrix": beam_stiffness_matrices(E * I, element_length)[0].cpu().numpy(),
        "global_stiffness_matrix": K.to_scipy().toarray(),
        "load_vector": loads.reshape(-1).copy(),
        "fixed_dofs": fixed,
        "free_dofs": [d for d in range(2 * n_nodes) if d not in set(fixed)],
        "displacement_vector": u.reshape(-1),
        "moment_vector": m,
        "shear_vector": v,
        "info": info,
    }
    return _RESULTS


def __getattr__(name):  # PEP 562: the reference's module-level result names, computed on demand
    lazy = ("element_stiffness_matrix", "global_stiffness_matrix", "load_vector", "fixed_dofs", "free_dofs",
            "displacement_vector", "moment_vector", "shear_vector")
    if name in lazy:
        res = _RESULTS if _RESULTS is not None else run()
        return res[name]
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")


def main():
    res = run()
    print("Displacements:", res["displacement_vector"])
    print("Bending Moments:", res["moment_vector"])
    print("Shear Forces:", res["shear_vector"])
    return res


if __name__ == "__main__":
    main()
