"""Counterpart of the reference's truss.py: pin-jointed members as axial springs.

`vec2` and `compute_forces` keep the reference's signatures (truss.py:9, 78); the endless
relaxation loop of the script (truss.py:95-119) is `relax(n_steps)`; and the stiffness-matrix
route the reference never takes -- the tangent of `compute_forces` at rest, K u = f
(SURVEY.md T1') -- is `solve_linear`, which is what BASELINE configs 1 and 5 run.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, core

stiffness = 1000.0


def vec2(x=0.0, y=0.0):
    return np.array([x, y, 0.0], dtype=np.float32)


nodes = np.array([vec2(0.0, 0.0), vec2(0.0, 1.0), vec2(1.0, 0.5)])
members = [[0, 2], [1, 2]]
loads = [[2, vec2(0.0, -100.0)]]


def _member_forces_device(nodes_d, members_d, k_d, displaced_d):
    lib = _lib.load()
    out = torch.zeros_like(nodes_d)
    _lib.check(lib.fea_truss_member_forces(nodes_d.data_ptr(), members_d.data_ptr(), k_d.data_ptr(),
                                           members_d.shape[0], displaced_d.data_ptr(), out.data_ptr(),
                                           core._stream()), "fea_truss_member_forces")
    return out


def compute_forces(nodes, members, displaced_nodes, forces, member_stiffness=None):
    """Accumulate the axial member forces into `forces` IN PLACE (truss.py:78-92):
    dl = |X_b - X_a| - |x_b - x_a|, F = -k dl, forces[a] += F e, forces[b] -= F e with e the
    current member axis.  Returns None.  Evaluated in FP64 on the device and added to the
    caller's array in its own dtype (the script's arrays are float32, truss.py:9-10).
    `member_stiffness` (scalar or per member) defaults to the module constant `stiffness`."""
    members_arr = np.asarray(members, dtype=np.int64).reshape(-1, 2)
    k = np.broadcast_to(np.asarray(stiffness if member_stiffness is None else member_stiffness, dtype=np.float64),
                        (members_arr.shape[0],))
    nodes_d = core.to_device(np.asarray(nodes, dtype=np.float64), torch.float64)
    disp_d = core.to_device(np.asarray(displaced_nodes, dtype=np.float64), torch.float64)
    f = _member_forces_device(nodes_d, core.to_device(members_arr, torch.int32),
                              core.to_device(np.ascontiguousarray(k), torch.float64), disp_d)
    forces += f.cpu().numpy().astype(forces.dtype)


def relax(nodes, members, loads, n_steps, member_stiffness=None):
    """`n_steps` passes of the script's loop (truss.py:97-119): member forces, residual at the
    first loaded node (what the script prints), then x_i += (load_i + f_i) / k for loaded nodes
    only.  Returns (displaced_nodes, residual history); dtype follows `nodes`."""
    displaced = np.array(nodes, copy=True)
    k = stiffness if member_stiffness is None else member_stiffness
    history = []
    for _ in range(n_steps):
        forces = np.zeros_like(displaced)
        compute_forces(nodes, members, displaced, forces, k)
        history.append(float(np.linalg.norm(loads[0][1] + forces[loads[0][0]])))
        for i, load in loads:
            displaced[i] += (load + forces[i, :]) / displaced.dtype.type(stiffness)
    return displaced, np.array(history)


def member_stiffness_matrices(nodes, members, k) -> torch.Tensor:
    """(M, 6, 6) tangent stiffness k [[cc^T, -cc^T], [-cc^T, cc^T]] per member (SURVEY.md T1')."""
    lib = _lib.load()
    host = lambda a, dt: a if isinstance(a, torch.Tensor) else np.asarray(a, dtype=dt)
    nodes_d = core.to_device(host(nodes, np.float64), torch.float64)
    members_d = core.to_device(host(members, np.int64), torch.int32).reshape(-1, 2)
    k_d = core.to_device(host(k, np.float64), torch.float64)
    m = members_d.shape[0]
    ke = torch.empty((m, 6, 6), dtype=torch.float64, device=nodes_d.device)
    status = core._status_slot()
    _lib.check(lib.fea_ke_truss(nodes_d.data_ptr(), members_d.data_ptr(), k_d.data_ptr(), m, ke.data_ptr(),
                                status.data_ptr(), core._stream()), "fea_ke_truss")
    core._check_status(status)
    return ke


def solve_linear(nodes, members, k, constraints, loads, tol: float = 1e-12, max_iter: int | None = None,
                 return_matrix: bool = False):
    """Linear truss K u = f.  nodes (N,3); members (M,2); k (M,); constraints (N,3) with 1 = fixed;
    loads (N,3) for one load case or (3N, R) for R load cases (batched multi-RHS PCG).
    Returns u with the shape of `loads` [, BlockCSR, SolveInfo]."""
    nodes_d = core.to_device(np.asarray(nodes, dtype=np.float64), torch.float64)
    members_d = core.to_device(np.asarray(members).reshape(-1, 2), torch.int32)
    k_d = core.to_device(np.asarray(k, dtype=np.float64), torch.float64)
    n = nodes_d.shape[0]
    fixed = core._fixed_mask(constraints, 3 * n)
    K = core.assemble_truss(nodes_d, members_d, k_d, fixed=fixed)
    loads_arr = loads if isinstance(loads, torch.Tensor) else np.asarray(loads, dtype=np.float64)
    multi = loads_arr.ndim == 2 and tuple(loads_arr.shape) != (n, 3)
    b = core.to_device(loads_arr, torch.float64)
    if multi:
        u, info = core.pcg_multi(K, b, tol=tol, max_iter=max_iter)
    else:
        u, info = core.pcg(K, b.reshape(-1), tol=tol, max_iter=max_iter)
    u_host = u.cpu().numpy().reshape(loads_arr.shape)
    if return_matrix:
        return u_host, K, info
    return u_host


def shipped_case():
    """truss.py:6-24 as a linear problem: nodes that carry no load never move in the script
    (truss.py:112-119), i.e. they are pinned; z is identically 0."""
    constraints = np.ones((3, 3), dtype=int)
    constraints[2, :2] = 0
    f = np.zeros((3, 3))
    f[loads[0][0]] = loads[0][1]
    return nodes.astype(np.float64), np.array(members), np.full(2, stiffness), constraints, f


LATTICE_DIRECTIONS = np.array(
    [[1, 0, 0], [0, 1, 0], [0, 0, 1],
     [1, 1, 0], [1, -1, 0], [1, 0, 1], [1, 0, -1], [0, 1, 1], [0, 1, -1],
     [1, 1, 1], [1, 1, -1], [1, -1, 1], [1, -1, -1]], dtype=np.int64)


def lattice_truss(n: int, n_rhs: int = 64, h: float = 1.0):
    """BASELINE config 5 generator, frozen in SURVEY.md §8(d): jittered cubic lattice of n^3 nodes
    (id = (iz*n + iy)*n + ix) with members along the 13 half-space neighbour directions (rigid),
    jitter U(-0.1h, 0.1h) and k ~ U(500, 1500) from default_rng(0), z = 0 layer pinned, loads =
    default_rng(1).standard_normal((3 n^3, n_rhs)).  n = 93 gives 10,224,788 members."""
    rng = np.random.default_rng(0)
    iz, iy, ix = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    grid = np.stack([ix.ravel(), iy.ravel(), iz.ravel()], axis=1)
    pts = grid * h + rng.uniform(-0.1 * h, 0.1 * h, size=(n**3, 3))
    ids = np.arange(n**3, dtype=np.int64)
    chunks = []
    for d in LATTICE_DIRECTIONS:
        tgt = grid + d
        ok = np.all((tgt >= 0) & (tgt < n), axis=1)
        chunks.append(np.stack([ids[ok], (tgt[ok, 2] * n + tgt[ok, 1]) * n + tgt[ok, 0]], axis=1))
    mem = np.concatenate(chunks).astype(np.int64)
    k = rng.uniform(500.0, 1500.0, size=mem.shape[0])
    constraints = np.zeros((n**3, 3), dtype=int)
    constraints[grid[:, 2] == 0] = 1
    f = np.random.default_rng(1).standard_normal((3 * n**3, n_rhs))
    return pts, mem, k, constraints, f
