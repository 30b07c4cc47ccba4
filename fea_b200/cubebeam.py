"""Counterpart of the reference's cubebeam.py: square-bar cantilever meshed with hex8 elements.

Same names and call signatures for the callables (`generate_quad_grid`, `solve`) and the same
module constants; the script body that the reference runs at import time (cubebeam.py:60-66,
111-124) is `main()` here, and `cantilever_case` builds BASELINE configs 3/4 at any size.
"""
from __future__ import annotations

import numpy as np

from . import model
from .utils import stack_faces_2d

psi = 6894.76
lbf = 4.44822
ft = 0.3048
inch = 0.0254

n_elements_width = 4
n_elements_height = 50

beam_width = 0.1
beam_length = 1.0
face_area = beam_width * beam_length
linear_load = 100.0 * lbf / ft
total_load = linear_load * beam_length
pressure = total_load / face_area

number_elements_face = (n_elements_width + 1) * (n_elements_height + 1)
force_per_element = total_load / number_elements_face


def generate_quad_grid(nx, ny, width, height):
    """Regular grid of quads (cubebeam.py:28-57): (nx+1)(ny+1) nodes with x fastest, elements
    [n1, n2, n4, n3] counter-clockwise.  Returns (nodes (n,2) float, elements (m,4) int)."""
    x = np.linspace(0, width, nx + 1)
    y = np.linspace(0, height, ny + 1)
    nodes = np.empty(((nx + 1) * (ny + 1), 2), dtype=float)
    nodes[:, 0] = np.tile(x, ny + 1)
    nodes[:, 1] = np.repeat(y, nx + 1)
    n1 = (np.arange(ny)[:, None] * (nx + 1) + np.arange(nx)[None, :]).reshape(-1)
    elements = np.stack([n1, n1 + 1, n1 + nx + 2, n1 + nx + 1], axis=1).astype(int)
    return nodes, elements


def generate_quad_grid_device(nx, ny, width, height):
    """`generate_quad_grid` written straight into device memory: CUDA tensors
    (nodes (n,2) f64, elements (m,4) int32) equal to the host builder's arrays element for element."""
    import torch

    from . import _lib, core

    lib = _lib.load()
    dev = core.device()
    nodes2d = torch.empty(((nx + 1) * (ny + 1), 2), dtype=torch.float64, device=dev)
    quads = torch.empty((nx * ny, 4), dtype=torch.int32, device=dev)
    _lib.check(lib.fea_mesh_quad_grid(nx, ny, float(width), float(height), nodes2d.data_ptr(), quads.data_ptr(),
                                      core._stream()), "fea_mesh_quad_grid")
    return nodes2d, quads


def cantilever_case_device(A, b, width=beam_width, length=beam_length):
    """`cantilever_case` built entirely on the device: (nodes (N,3) f64, elements (M,8) int32,
    fixed (3N,) uint8, loads (3N,) f64) CUDA tensors, ready for core.assemble_hex8 / core.solve_system."""
    import torch

    from .utils import stack_faces_2d_device

    nodes2d, quads = generate_quad_grid_device(b, b, width, width)
    z = np.linspace(0, length, A + 1)
    nodes, elements = stack_faces_2d_device(nodes2d, quads, z)
    fixed = (nodes[:, 2] == 0).repeat_interleave(3).to(torch.uint8)
    loads = torch.zeros(nodes.shape, dtype=torch.float64, device=nodes.device)
    loads[:, 1] = (nodes[:, 1] == 0).to(torch.float64) * (linear_load * length / ((b + 1) * (A + 1)))
    return nodes, elements, fixed, loads.reshape(-1)


def solve(nodes, elements, constraints, forces):
    """(displacements, forces) = solve(nodes, elements, constraints, forces), cubebeam.py:79-108,
    with E = 1e7 psi and nu = 0.3 as hard-wired there (cubebeam.py:84)."""
    return model.solve_hex8(nodes, elements, constraints, forces, 10_000_000 * psi, 0.3)


def bar_case(n_width, n_node_layers, f_node, width=beam_width, length=beam_length):
    """Mesh + boundary conditions of the bar: all DOF fixed on z = 0 (cubebeam.py:112-114), load
    (0, f_node, 0) on every node with y = 0 (cubebeam.py:116-118)."""
    nodes2d, face2ds = generate_quad_grid(n_width, n_width, width, width)
    nodes, elements = stack_faces_2d(nodes2d, face2ds, np.linspace(0, length, n_node_layers))
    constraints = np.zeros(nodes.shape, dtype=int)
    constraints[nodes[:, 2] == 0] = 1
    forces = np.zeros(nodes.shape, dtype=float)
    forces[nodes[:, 1] == 0] += np.array([0, f_node, 0])
    return nodes, elements, constraints, forces


def shipped_case():
    """The exact inputs of the reference's own run (4x4 quads, linspace(0, 1, 50) node layers)."""
    return bar_case(n_elements_width, n_elements_height, force_per_element)


def cantilever_case(A, b, width=beam_width, length=beam_length):
    """BASELINE configs 3/4: A element layers along z, b x b elements in the section,
    f = total_load / ((b+1)(A+1)) on the nodes of the y = 0 face."""
    return bar_case(b, A + 1, linear_load * length / ((b + 1) * (A + 1)), width, length)


def main():
    nodes, elements, constraints, forces = shipped_case()
    displacements, nodal_forces = solve(nodes, elements, constraints, forces)
    np.set_printoptions(precision=5, linewidth=200, suppress=True)
    print("forces", nodal_forces / lbf, sep="\n")
    print("displacements", displacements / inch, sep="\n")
    return displacements, nodal_forces


if __name__ == "__main__":
    main()
