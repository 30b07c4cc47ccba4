"""Counterpart of the reference's utils.py hot-path callables (plot helpers are out of scope)."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, core


def hexahedral_stiffness_matrices(nodes, elements, E: float, nu: float) -> torch.Tensor:
    """Batched Ke, (M, 24, 24) float64 on the device.  Same mathematics as
    `hexahedral_stiffness_matrix` (utils.py:127-239) for every row of `elements` (M, 8)."""
    lib = _lib.load()
    nodes_d = core.to_device(nodes, torch.float64)
    elements_d = core.to_device(elements, torch.int32)
    if elements_d.ndim != 2 or elements_d.shape[1] != 8:
        raise ValueError("elements must be (M, 8)")
    m = elements_d.shape[0]
    ke = torch.empty((m, 24, 24), dtype=torch.float64, device=nodes_d.device)
    if m == 0:
        return ke
    status = core._status_slot()
    _lib.check(lib.fea_ke_hex8(nodes_d.data_ptr(), elements_d.data_ptr(), m, float(E), float(nu), ke.data_ptr(),
                               status.data_ptr(), core._stream()), "fea_ke_hex8")
    core._check_status(status)
    return ke


def hexahedral_stiffness_matrix(nodes, E, nu):
    """Ke (24, 24) of one 8-node hexahedron: nodes (8, 3), Young's modulus E, Poisson ratio nu
    (utils.py:127-239).  Raises ValueError("Jacobian determinant is non-positive. ...") like the
    reference (utils.py:212-215)."""
    nodes = np.asarray(nodes, dtype=np.float64)
    if nodes.shape != (8, 3):
        raise ValueError("nodes must be (8, 3)")
    ke = hexahedral_stiffness_matrices(nodes, np.arange(8, dtype=np.int32)[None, :], E, nu)
    return ke[0].cpu().numpy()


def stack_faces_2d(nodes2d, faces2d, z_heights):
    """Extrude 2-D quads into hex8 elements (utils.py:356-376): node id = layer * n2d + i,
    element = [face + lo, face + hi]; `z_heights` lists the NODE layers.  Host arrays, vectorised."""
    nodes2d = np.asarray(nodes2d, dtype=np.float64)
    faces2d = np.asarray(faces2d)
    z = np.asarray(z_heights, dtype=np.float64)
    n2d, nl = nodes2d.shape[0], z.shape[0]
    nodes3d = np.zeros((n2d * nl, 3))
    nodes3d[:, :2] = np.tile(nodes2d, (nl, 1))
    nodes3d[:, 2] = np.repeat(z, n2d)
    lo = (np.arange(max(nl - 1, 0)) * n2d)[:, None, None]
    bottom = faces2d[None, :, :] + lo
    elements = np.concatenate([bottom, bottom + n2d], axis=2).reshape(-1, 8)
    return nodes3d, elements


def stack_faces_2d_device(nodes2d, faces2d, z_heights):
    """`stack_faces_2d` with the mesh generated directly in device memory.
    Returns (nodes3d float64 (N,3), elements int32 (M,8)) CUDA tensors."""
    lib = _lib.load()
    n2 = core.to_device(nodes2d, torch.float64)
    f2 = core.to_device(faces2d, torch.int32)
    z = core.to_device(np.asarray(z_heights, dtype=np.float64), torch.float64)
    n2d, nf, nl = n2.shape[0], f2.shape[0], z.shape[0]
    nodes3d = torch.empty((n2d * nl, 3), dtype=torch.float64, device=n2.device)
    elements = torch.empty((nf * max(nl - 1, 0), 8), dtype=torch.int32, device=n2.device)
    _lib.check(lib.fea_mesh_extrude(n2.data_ptr(), n2d, f2.data_ptr(), nf, z.data_ptr(), nl, nodes3d.data_ptr(),
                                    elements.data_ptr(), core._stream()), "fea_mesh_extrude")
    return nodes3d, elements


_QUAD = np.array([[0, 1, 2, 3]], dtype=int)
_HEX_FACES = np.array(
    [[0, 1, 2, 3], [4, 5, 6, 7], [0, 1, 5, 4], [1, 2, 6, 5], [2, 3, 7, 6], [3, 0, 4, 7]], dtype=int
)


def faces_from_nodes2d(selection):
    """(1, 4) face of a quad selection (utils.py:379-387)."""
    return np.asarray(selection)[_QUAD]


def faces_from_nodes(selection):
    """(6, 4) quad faces of a hex8 node selection (utils.py:390-403)."""
    return np.asarray(selection)[_HEX_FACES]
