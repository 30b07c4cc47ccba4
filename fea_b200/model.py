"""The solve() bodies shared by the reference-named modules (cubebeam.py:79-108 == fea.py:86-115)."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, core

PSI = 6894.76
E_DEFAULT = 10_000_000 * PSI  # cubebeam.py:84
NU_DEFAULT = 0.3


def solve_hex8(nodes, elements, constraints, forces, E: float = E_DEFAULT, nu: float = NU_DEFAULT,
               tol: float = 1e-12, max_iter: int | None = None, return_info: bool = False, history: bool = False):
    """solve(nodes, elements, constraints, forces) -> (displacements (N,3), forces (N,3)).

    Host arrays in, host arrays out, like the reference (cubebeam.py:79-108): assemble K from hex8
    elements, drop constrained DOF (constraints != 0, homogeneous), solve K_ff u_f = f_f, expand,
    and return the nodal forces K_full u (applied loads on free DOF, reactions on constrained DOF).
    Inputs are not modified.  ValueError for an inverted element (utils.py:212-215),
    numpy.linalg.LinAlgError for a singular reduced system (what np.linalg.solve raises).

    Under an initialised NCCL process group with more than one rank (torchrun, one rank per GPU) the
    call is COLLECTIVE: every rank passes the same host arrays, the mesh is split into slabs of node
    layers, and rank 0 returns the full host arrays while the other ranks return (None, None)
    (fea_b200/dist.py:solve_hex8).  FEA_DIST=0 keeps every rank on its own single-GPU solve.
    """
    import os

    from . import dist as fdist

    if fdist.active_world() > 1 and os.environ.get("FEA_DIST", "1") != "0" and not history:
        res = fdist.solve_hex8_or_none(nodes, elements, constraints, forces, E, nu, tol=tol, max_iter=max_iter,
                                       return_info=return_info)
        if res is not None:
            return res
    nodes_d = core.to_device(nodes, torch.float64)
    if nodes_d.ndim != 2 or nodes_d.shape[1] != 3:
        raise ValueError("nodes must be (N, 3)")
    elements_d = core.to_device(elements, torch.int32)
    if elements_d.ndim != 2 or elements_d.shape[1] != 8:
        raise ValueError("elements must be (M, 8)")
    n_nodes = nodes_d.shape[0]
    fixed = core._fixed_mask(constraints, 3 * n_nodes)
    loads = core.to_device(forces, torch.float64).reshape(-1)
    if loads.numel() != 3 * n_nodes:
        raise ValueError("forces must be (N, 3)")
    K = core.assemble_hex8(nodes_d, elements_d, E, nu, fixed=fixed, mode=_lib.ASSEMBLE_FULL)
    u, reactions, info = core.solve_system(K, loads, tol=tol, max_iter=max_iter, history=history)
    shape = tuple(np.shape(nodes)) if not isinstance(nodes, torch.Tensor) else tuple(nodes.shape)
    out = torch.stack([u, reactions]).cpu().numpy()
    displacements, nodal_forces = out[0].reshape(shape), out[1].reshape(shape)
    if return_info:
        return displacements, nodal_forces, info, K
    return displacements, nodal_forces
