"""Oracle vs the LIVE reference; only runs where /root/reference exists (the build container)."""
import numpy as np
import pytest

from oracle import fea_oracle as fo
from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present")


def test_ke_random_hexes_vs_reference():
    U = ref_loader.load_utils()
    rng = np.random.default_rng(11)
    base = fo.HEX8_SIGNS * 0.5
    for _ in range(10):
        x = base * rng.uniform(0.5, 2.0, size=3) + rng.uniform(-0.15, 0.15, size=(8, 3))
        ref = U.hexahedral_stiffness_matrix(x, 2.5e9, 0.27)
        assert np.abs(fo.hex8_ke(x, 2.5e9, 0.27) - ref).max() <= 1e-15 * np.abs(ref).max()
        got = fo.hex8_ke_batched(x, np.arange(8)[None], 2.5e9, 0.27)[0]
        assert np.abs(got - ref).max() <= 1e-13 * np.abs(ref).max()


def test_mesh_builders_vs_reference():
    U = ref_loader.load_utils()
    n2, q2 = fo.generate_quad_grid(3, 5, 0.2, 0.4)
    z = np.linspace(0, 2, 7)
    rn, re = U.stack_faces_2d(n2, q2, z)
    on, oe = fo.stack_faces_2d(n2, q2, z)
    assert np.array_equal(rn, on) and np.array_equal(re, oe)


def test_dense_reference_solve_small():
    """The reference's own dense solve() body (cubebeam.py:79-108), re-run through its functions on a
    small mesh, against the sparse restatement."""
    U = ref_loader.load_utils()
    nodes, elements, cons, forces = fo.cantilever_case(6, 2)
    K = np.zeros((nodes.size, nodes.size))
    for el in elements:
        ke = U.hexahedral_stiffness_matrix(nodes[el], fo.E_HEX, fo.NU_HEX)
        d = np.array([i * 3 + j for i in el for j in range(3)])
        K[np.ix_(d, d)] += ke
    free = np.where(cons.flatten() == 0)[0]
    uf = np.linalg.solve(K[np.ix_(free, free)], forces.flatten()[free])
    u = np.zeros(nodes.size)
    u[free] = uf
    ou, of, info = fo.solve_hex8(nodes, elements, cons, forces, method="direct")
    assert np.abs(info["K"].toarray() - K).max() <= 1e-14 * np.abs(K).max()
    assert np.abs(ou.ravel() - u).max() <= 1e-9 * np.abs(u).max()
    assert np.abs(of.ravel() - K @ u).max() <= 1e-8 * np.abs(K @ u).max()
