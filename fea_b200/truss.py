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


def _dtype_of(nodes, fp32):
    """float32 when asked for, or when the caller's arrays are float32 like the script's (truss.py:9-10)."""
    if fp32 is None:
        fp32 = np.asarray(nodes).dtype == np.float32
    return (np.float32, torch.float32, 1) if fp32 else (np.float64, torch.float64, 0)


def _incidence(members_d, n_nodes):
    """node -> (member, end) incidence lists, ascending member order (the symbolic pass)."""
    return core.symbolic(members_d, n_nodes)


def compute_forces(nodes, members, displaced_nodes, forces, member_stiffness=None, fp32=None):
    """Accumulate the axial member forces into `forces` IN PLACE (truss.py:78-92):
    dl = |X_b - X_a| - |x_b - x_a|, F = -k dl, forces[a] += F e, forces[b] -= F e with e the
    current member axis.  Returns None.  Evaluated on the device node by node, every node summing its
    members in list order (deterministic, the reference's own summation order), in the dtype of `nodes`
    (the script's arrays are float32, truss.py:9-10) unless `fp32` says otherwise.
    `member_stiffness` (scalar or per member) defaults to the module constant `stiffness`."""
    lib = _lib.load()
    np_t, th_t, flag = _dtype_of(nodes, fp32)
    members_arr = np.asarray(members, dtype=np.int64).reshape(-1, 2)
    k = np.ascontiguousarray(np.broadcast_to(
        np.asarray(stiffness if member_stiffness is None else member_stiffness, dtype=np_t), (members_arr.shape[0],)))
    nodes_d = core.to_device(np.asarray(nodes, dtype=np_t), th_t)
    disp_d = core.to_device(np.asarray(displaced_nodes, dtype=np_t), th_t)
    members_d = core.to_device(members_arr, torch.int32)
    k_d = core.to_device(k, th_t)
    n = nodes_d.shape[0]
    inc = _incidence(members_d, n)
    out = torch.zeros_like(nodes_d)
    _lib.check(lib.fea_truss_member_forces(nodes_d.data_ptr(), members_d.data_ptr(), k_d.data_ptr(),
                                           members_d.shape[0], n, inc.n2e_ptr.data_ptr(), inc.n2e.data_ptr(),
                                           disp_d.data_ptr(), out.data_ptr(), flag, core._stream()),
               "fea_truss_member_forces")
    forces += out.cpu().numpy().astype(forces.dtype)


def relax(nodes, members, loads, n_steps, member_stiffness=None, fp32=None, return_residual=False):
    """`n_steps` passes of the script's loop (truss.py:97-119) ON THE DEVICE, without a host round trip
    per step: member forces at the loaded nodes from the current positions, residual at the first loaded
    node (what the script prints), then x_i += (load_i + f_i) / stiffness for loaded nodes only.
    `loads` is the script's list of [node, vector] (distinct nodes).  Returns
    (displaced_nodes, residual history) in the dtype of `nodes` (float32 in the script)."""
    lib = _lib.load()
    np_t, th_t, flag = _dtype_of(nodes, fp32)
    nodes_np = np.asarray(nodes)
    members_arr = np.asarray(members, dtype=np.int64).reshape(-1, 2)
    k = np.ascontiguousarray(np.broadcast_to(
        np.asarray(stiffness if member_stiffness is None else member_stiffness, dtype=np_t), (members_arr.shape[0],)))
    load_nodes = np.array([int(i) for i, _ in loads], dtype=np.int32)
    if np.unique(load_nodes).size != load_nodes.size:
        raise ValueError("loads must name distinct nodes")
    load_vecs = np.ascontiguousarray(np.stack([np.asarray(v, dtype=np_t).reshape(3) for _, v in loads]))
    nodes_d = core.to_device(nodes_np.astype(np_t), th_t)
    members_d = core.to_device(members_arr, torch.int32)
    k_d = core.to_device(k, th_t)
    n = nodes_d.shape[0]
    inc = _incidence(members_d, n)
    displaced = nodes_d.clone()  # truss.py:95
    ln_d, lv_d = core.to_device(load_nodes, torch.int32), core.to_device(load_vecs, th_t)
    residual = torch.zeros_like(lv_d)
    history = torch.zeros(max(n_steps, 1), dtype=torch.float64, device=nodes_d.device)
    _lib.check(lib.fea_truss_relax(nodes_d.data_ptr(), members_d.data_ptr(), k_d.data_ptr(), members_d.shape[0], n,
                                   inc.n2e_ptr.data_ptr(), inc.n2e.data_ptr(), ln_d.data_ptr(), lv_d.data_ptr(),
                                   load_nodes.size, float(stiffness), int(n_steps), displaced.data_ptr(),
                                   residual.data_ptr(), history.data_ptr(), flag, core._stream()), "fea_truss_relax")
    out = displaced.cpu().numpy().astype(nodes_np.dtype if nodes_np.dtype.kind == "f" else np_t)
    hist = history[:n_steps].cpu().numpy()
    if return_residual:
        return out, hist, residual.cpu().numpy()
    return out, hist


def relax_device(steps=40):
    """The script's own case (truss.py:6-24, float32) for `steps` passes; (history, displaced)."""
    displaced, hist = relax(nodes, members, loads, steps)
    return hist, displaced


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


def lattice_truss_device(n: int, h: float = 1.0, seed: int = 0):
    """Nodes, members and spring rates of `lattice_truss(n)` generated directly in device memory
    (SURVEY.md §8(f) N2): numpy's PCG64 stream of default_rng(seed) is reproduced on the device by LCG
    jump-ahead, so the arrays equal the host generator's element for element.  Returns CUDA tensors
    (nodes (n^3,3) f64, members (M,2) int32, k (M,) f64, constraints (n^3,3) uint8).  The load cases stay a
    host draw: standard_normal is a rejection sampler (ziggurat), inherently sequential."""
    import ctypes

    lib = _lib.load()
    dev = core.device()
    st = np.random.PCG64(seed).state["state"]
    mask = (1 << 64) - 1
    words = (ctypes.c_uint64 * 4)(st["state"] >> 64, st["state"] & mask, st["inc"] >> 64, st["inc"] & mask)
    m = int(lib.fea_mesh_lattice_members(n))
    pts = torch.empty((n**3, 3), dtype=torch.float64, device=dev)
    mem = torch.empty((m, 2), dtype=torch.int32, device=dev)
    k = torch.empty(m, dtype=torch.float64, device=dev)
    scratch = torch.empty(14, dtype=torch.int64, device=dev)
    _lib.check(lib.fea_mesh_lattice(n, float(h), ctypes.addressof(words), pts.data_ptr(), mem.data_ptr(), k.data_ptr(),
                                    scratch.data_ptr(), core._stream()), "fea_mesh_lattice")
    constraints = torch.zeros((n**3, 3), dtype=torch.uint8, device=dev)
    constraints[: n * n] = 1  # iz = 0: the first n^2 node ids
    return pts, mem, k, constraints
