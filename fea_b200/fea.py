"""Counterpart of the reference's fea.py: thin-walled tube (periodic cross-section) under a
cosine-distributed load.  `solve` is the same callable as cubebeam.solve (fea.py:86-115 is
byte-identical to cubebeam.py:79-108)."""
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
beam_length = 1.0

n_elements = 26
outer_radius = 4 * inch
inner_radius = 3.9 * inch


def tube_section(n_seg=n_elements, r_in=inner_radius, r_out=outer_radius):
    """Cross-section of fea.py:28-48: inner ring then outer ring, faces
    [i, i+n, (i+1)%n+n, (i+1)%n], plus the 2-D load of fea.py:51-55 (first half of the outer ring)."""
    thetas = np.linspace(0, np.pi * 2, n_seg, endpoint=False).reshape(-1, 1)
    unit_points = np.hstack([np.cos(thetas), np.sin(thetas)])
    nodes2d = np.vstack([unit_points * r_in, unit_points * r_out])
    i = np.arange(n_seg)
    face2ds = np.stack([i, i + n_seg, (i + 1) % n_seg + n_seg, (i + 1) % n_seg], axis=1)
    forces2d = np.zeros_like(nodes2d)
    half = slice(n_seg, (3 * n_seg) // 2)
    forces2d[half, 1] = -np.cos(np.pi / 2 * nodes2d[half, 0] / r_out) * np.pi / 4 / r_out
    return nodes2d, face2ds, forces2d


def tube_section_device(n_seg=n_elements, r_in=inner_radius, r_out=outer_radius):
    """Nodes and periodic quads of `tube_section` generated in device memory.
    The connectivity equals the host builder's; the coordinates agree to the last bit or two (the
    device's cos / sin are not numpy's)."""
    import torch

    from . import _lib, core

    lib = _lib.load()
    dev = core.device()
    nodes2d = torch.empty((2 * n_seg, 2), dtype=torch.float64, device=dev)
    quads = torch.empty((n_seg, 4), dtype=torch.int32, device=dev)
    _lib.check(lib.fea_mesh_tube_section(n_seg, float(r_in), float(r_out), nodes2d.data_ptr(), quads.data_ptr(),
                                         core._stream()), "fea_mesh_tube_section")
    return nodes2d, quads


def shipped_case():
    """Inputs of the reference's own run, including its load layout: fea.py:71 repeats the 2-D
    load along axis 0 although nodes are layer-major (quirk Q5) -- reproduced as is."""
    nodes2d, face2ds, forces2d = tube_section()
    nodes, elements = stack_faces_2d(nodes2d, face2ds, np.linspace(0, beam_length, n_elements_height))
    forces = np.zeros_like(nodes)
    forces[:, :2] = forces2d.repeat(n_elements_height, axis=0)
    constraints = np.zeros(nodes.shape, dtype=int)
    constraints[nodes[:, 2] == 0] = 1
    return nodes, elements, constraints, forces


def solve(nodes, elements, constraints, forces):
    """(displacements, forces) = solve(nodes, elements, constraints, forces), fea.py:86-115."""
    return model.solve_hex8(nodes, elements, constraints, forces, 10_000_000 * psi, 0.3)


def main():
    nodes, elements, constraints, forces = shipped_case()
    displacements, nodal_forces = solve(nodes, elements, constraints, forces)
    np.set_printoptions(precision=5, linewidth=200, suppress=True)
    print("forces", nodal_forces / lbf, sep="\n")
    print("displacements", displacements / inch, sep="\n")
    return displacements, nodal_forces


if __name__ == "__main__":
    main()
