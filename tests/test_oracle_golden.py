"""Pin oracle/fea_oracle.py to the committed outputs of the reference (tests/golden, K1..K8)."""
import numpy as np
import pytest

from oracle import fea_oracle as fo


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


def test_k1_cube_ke(golden):
    g = golden("hex8_single.npz")
    ke = fo.hex8_ke(g["cube"], 1000, 0.0)
    assert rel(ke, g["ke_cube"]) < 1e-14
    assert rel(fo.hex8_ke_batched(g["cube"], np.arange(8)[None], 1000, 0.0)[0], g["ke_cube"]) < 1e-14
    # SURVEY.md K1 properties of the reference's own matrix
    K = g["ke_cube"]
    assert abs(np.trace(K) - 10666.666666666666) < 1e-9
    np.testing.assert_allclose(K[0, :6], [444.444444, 83.333333, 83.333333, -111.111111, -83.333333, -83.333333], rtol=1e-8)
    assert np.abs(K - K.T).max() < 1e-12
    w = np.linalg.eigvalsh((K + K.T) / 2)
    assert (np.abs(w) < 1e-9).sum() == 6 and abs(w.max() - 1000) < 1e-9


def test_k2_round_trip(golden):
    g = golden("hex8_single.npz")
    ke = fo.hex8_ke(g["cube"], 1000, 0.0)
    f = (ke @ g["disp"].flatten()).reshape(-1, 3)
    assert rel(f, g["f_cube"]) < 1e-13
    cons = np.zeros((8, 3), dtype=int)
    cons[:4] = 1
    K = fo.assemble_csr(np.arange(8)[None], ke[None], 8, 3)
    u, _, _ = fo.solve_system(K, cons, f, method="dense")
    assert np.abs(u.reshape(8, 3) - g["u_back"]).max() < 1e-14
    assert np.abs(u.reshape(8, 3) - g["disp"]).max() < 1e-14


def test_k3_inverted_raises(golden):
    g = golden("hex8_single.npz")
    assert str(g["inverted_message"]) == fo.JACOBIAN_MESSAGE
    with pytest.raises(ValueError, match="Jacobian determinant is non-positive"):
        fo.hex8_ke(g["inverted"], 1000, 0.0)
    with pytest.raises(ValueError, match="Jacobian determinant is non-positive"):
        fo.hex8_ke_batched(g["inverted"], np.arange(8)[None], 1000, 0.0)


def test_distorted_ke(golden):
    g = golden("hex8_single.npz")
    E, nu = float(g["E"]), float(g["nu"])
    X = g["dist_nodes"]
    nodes = X.reshape(-1, 3)
    elements = np.arange(nodes.shape[0]).reshape(-1, 8)
    kb = fo.hex8_ke_batched(nodes, elements, E, nu)
    assert rel(kb, g["ke_dist"]) < 1e-12
    for x, k in zip(X[:4], g["ke_dist"]):
        assert rel(fo.hex8_ke(x, E, nu), k) < 1e-14


def test_affine_closed_form_ke(golden):
    """The closed form the CUDA assembly uses on exactly affine elements (oracle.hex8_ke_affine) against the
    reference-generated Ke of the cube (K1) and against the Gauss-point oracle on sheared parallelepipeds; the
    trilinear coefficients that decide affinity vanish exactly on grid elements and not on distorted ones."""
    g = golden("hex8_single.npz")
    assert rel(fo.hex8_ke_affine(g["cube"], 1000, 0.0), g["ke_cube"]) < 1e-13
    rng = np.random.default_rng(0)
    cube = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], float)
    for _ in range(5):
        T = np.eye(3) + 0.3 * rng.standard_normal((3, 3))
        if np.linalg.det(T) < 0:
            T[0] *= -1
        X = cube @ T.T * 0.37 + rng.standard_normal(3)
        ref = fo.hex8_ke(X, fo.E_HEX, fo.NU_HEX)
        assert np.abs(fo.hex8_ke_affine(X, fo.E_HEX, fo.NU_HEX) - ref).max() / np.abs(ref).max() < 1e-12
    nodes, elements, _, _ = fo.cantilever_case(6, 3)
    for e in elements:
        assert np.all(fo.hex8_trilinear_coefficients(nodes[e])[[3, 5, 6, 7]] == 0.0)  # exactly affine: grid rows share values
    tn, te = fo.tube_case()[:2]
    assert all(np.abs(fo.hex8_trilinear_coefficients(tn[e])[[3, 5, 6, 7]]).max() > 0 for e in te[:50])


def test_k4_stack_faces(golden):
    g = golden("stack_faces.npz")
    n, e = fo.stack_faces_2d(np.array([[0.0, 0], [1, 0], [1, 1], [0, 1]]), np.array([[0, 1, 2, 3]]), [0.0, 1.0, 2.0])
    assert np.array_equal(n, g["nodes"]) and np.array_equal(e, g["elements"])
    assert e.dtype == np.int64


def test_k5_cubebeam(golden):
    g = golden("cubebeam.npz")
    nodes, elements, cons, forces = fo.cubebeam_case()
    assert np.array_equal(nodes, g["nodes"])
    assert np.array_equal(elements, g["elements"])
    assert np.array_equal(cons, g["constraints"])
    assert np.array_equal(forces, g["forces_in"])
    qn, qe = fo.generate_quad_grid(3, 2, 0.3, 0.2)
    assert np.array_equal(qn, g["quad_nodes"]) and np.array_equal(qe, g["quad_elements"])
    assert rel(fo.hex8_ke(nodes[elements[0]], fo.E_HEX, fo.NU_HEX), g["ke0"]) < 1e-14
    u, f, info = fo.solve_hex8(nodes, elements, cons, forces, method="direct")
    assert rel(u, g["displacements"]) < 1e-9
    assert rel(f, g["forces_out"]) < 1e-9
    # SURVEY.md K5 numbers
    assert abs(np.abs(g["displacements"]).max() - 3.050405508343811e-04) < 1e-12
    assert abs(g["forces_out"][nodes[:, 2] == 0][:, 1].sum() + 1402.158792651803) < 1e-6
    # structural pattern (H5): 225,108 entries, more than the numerically non-zero 223,417
    assert info["K"].nnz == 225108
    up, _, ip = fo.solve_hex8(nodes, elements, cons, forces, method="pcg")
    assert rel(up, g["displacements"]) < 1e-8
    assert ip["history"][-1] <= 1e-12


def test_k6_tube(golden):
    g = golden("fea_tube.npz")
    nodes, elements, cons, forces = fo.tube_case()
    assert np.array_equal(nodes, g["nodes"])
    assert np.array_equal(elements, g["elements"])
    assert np.array_equal(cons, g["constraints"])
    assert np.array_equal(forces, g["forces_in"])
    u, f, _ = fo.solve_hex8(nodes, elements, cons, forces, method="direct")
    assert rel(u, g["displacements"]) < 1e-8
    assert rel(f, g["forces_out"]) < 1e-8


def test_k7_euler_bernoulli(golden):
    g = golden("euler_bernoulli.npz")
    E, I, L, q, n, Le = g["params"]
    n = int(n)
    assert np.array_equal(fo.beam_ke(E * I, Le), g["element_stiffness_matrix"])
    elements, EI, Ls, cons, loads = fo.beam_fixed_fixed_case(n, E, I, L, q)
    assert np.array_equal(loads.ravel(), g["load_vector"])
    assert np.array_equal(fo.free_dofs(cons), g["free_dofs"])
    u, K, _ = fo.solve_beam(elements, EI, Ls, cons, loads, method="dense")
    assert rel(K.toarray(), g["global_stiffness_matrix"]) < 1e-15
    assert rel(u.ravel(), g["displacement_vector"]) < 1e-9
    M, V = fo.beam_moment_shear(g["displacement_vector"], EI, Ls)
    assert rel(M, g["moment_vector"]) < 1e-12 and rel(V, g["shear_vector"]) < 1e-12
    assert abs(g["displacement_vector"][n] - 1.2400793650878108e-05) < 1e-15


def test_k8_truss(golden):
    g = golden("truss.npz")
    nodes32, members = g["nodes"], g["members"]
    assert nodes32.dtype == np.float32
    # float32 relaxation reproduces the reference's printed residual history bit for bit
    disp = nodes32.copy()
    loads = [(int(g["load_node"]), g["load"])]
    for it in range(len(g["residual_history"])):
        r = fo.truss_relax_step(nodes32, members, disp, loads, float(g["stiffness"]))
        assert r == pytest.approx(float(g["residual_history"][it]), rel=1e-6)
        np.testing.assert_allclose(disp, g["displaced_history"][it], rtol=1e-6, atol=1e-7)
    # float64 probe of compute_forces
    n64 = nodes32.astype(np.float64)
    f = np.zeros_like(n64)
    fo.truss_compute_forces(n64, members, g["probe"], f, float(g["stiffness"]))
    assert rel(f, g["f_probe"]) < 1e-14
    # linearisation T1': K from truss_ke equals the finite-difference tangent of the reference
    Ke = fo.truss_ke_batched(n64, members, np.full(len(members), float(g["stiffness"])))
    K = fo.assemble_csr(members, Ke, 3, 3).toarray()
    assert np.abs(K - g["tangent_fd"]).max() < 1e-4
    nodes, mem, k, cons, loads = fo.truss_shipped_case()
    K = fo.assemble_csr(mem, fo.truss_ke_batched(nodes, mem, k), 3, 3)
    u, _, _ = fo.solve_system(K, cons, loads, method="dense")
    np.testing.assert_allclose(u.reshape(3, 3)[2], [0.0, -0.25, 0.0], atol=1e-15)
    free = fo.free_dofs(cons)
    np.testing.assert_allclose(K.toarray()[np.ix_(free, free)], np.diag([1600.0, 400.0]), atol=1e-12)


def test_pattern_is_structural():
    nodes, elements, cons, forces = fo.cantilever_case(6, 3)
    indptr, indices = fo.structural_pattern(elements, nodes.shape[0], 3)
    n1, n2, n3 = 4, 4, 7
    assert indices.size == 9 * (3 * n1 - 2) * (3 * n2 - 2) * (3 * n3 - 2)
    K = fo.assemble_csr(elements, fo.hex8_ke_batched(nodes, elements, fo.E_HEX, fo.NU_HEX), nodes.shape[0], 3)
    assert np.array_equal(K.indptr, indptr) and np.array_equal(K.indices, indices)


def test_lattice_truss_counts_and_rigidity():
    n = 4
    nodes, members, k, cons, loads = fo.lattice_truss_case(n, n_rhs=3)
    assert members.shape[0] == 3 * n * n * (n - 1) + 6 * n * (n - 1) ** 2 + 4 * (n - 1) ** 3
    K = fo.assemble_csr(members, fo.truss_ke_batched(nodes, members, k), n**3, 3)
    free = fo.free_dofs(cons)
    w = np.linalg.eigvalsh(K[free][:, free].toarray())
    assert w.min() > 1e-3  # rigid once the z=0 face is pinned
    X, iters = fo.jacobi_pcg_multi(K[free][:, free].tocsr(), loads[free], tol=1e-12)
    Xd = np.linalg.solve(K[free][:, free].toarray(), loads[free])
    assert rel(X, Xd) < 1e-9


def test_multi_rhs_matches_single():
    nodes, elements, cons, forces = fo.cantilever_case(5, 2)
    K = fo.assemble_csr(elements, fo.hex8_ke_batched(nodes, elements, 1.0, 0.3), nodes.shape[0], 3)
    free = fo.free_dofs(cons)
    A = K[free][:, free].tocsr()
    B = np.random.default_rng(5).standard_normal((free.size, 3))
    X, iters = fo.jacobi_pcg_multi(A, B, tol=1e-12)
    for j in range(3):
        x, it, _ = fo.jacobi_pcg(A, B[:, j], tol=1e-12)
        assert abs(it - iters[j]) <= 3  # dot-product summation order differs (einsum vs @)
        assert rel(X[:, j], x) < 1e-9


# ------------------------------------------------------------------------------------------------
# oracle/fea_oracle_c.c (OpenMP C restatement, the multi-core CPU baseline of bench.py): pinned to
# the same reference-generated fixtures and to the numpy oracle.
# ------------------------------------------------------------------------------------------------
def test_c_oracle_ke_golden(golden):
    from oracle import c_oracle as co

    g = golden("hex8_single.npz")
    assert rel(co.hex8_ke_batched(g["cube"], np.arange(8)[None], 1000, 0.0)[0], g["ke_cube"]) < 1e-14
    E, nu = float(g["E"]), float(g["nu"])
    nodes = g["dist_nodes"].reshape(-1, 3)
    elements = np.arange(nodes.shape[0]).reshape(-1, 8)
    assert rel(co.hex8_ke_batched(nodes, elements, E, nu), g["ke_dist"]) < 1e-12
    with pytest.raises(ValueError, match="Jacobian determinant is non-positive"):
        co.hex8_ke_batched(g["inverted"], np.arange(8)[None], 1000, 0.0)


def test_c_oracle_assembly_and_pcg_vs_numpy_oracle(golden):
    from oracle import c_oracle as co

    g = golden("cubebeam.npz")
    nodes, elements, cons, forces = fo.cubebeam_case()
    ip, ix = fo.structural_pattern(elements, nodes.shape[0], 3)
    ip2, ix2 = co.dof_pattern(elements, nodes.shape[0], 3)
    assert np.array_equal(ip, ip2) and np.array_equal(ix, ix2)
    K = fo.assemble_csr(elements, fo.hex8_ke_batched(nodes, elements, fo.E_HEX, fo.NU_HEX), nodes.shape[0], 3)
    Kc = co.assemble_hex8(nodes, elements, fo.E_HEX, fo.NU_HEX)
    assert np.array_equal(K.indptr, Kc.indptr) and np.array_equal(K.indices, Kc.indices)
    assert rel(Kc.data, K.data) < 1e-13
    x = np.random.default_rng(0).standard_normal(K.shape[0])
    assert rel(co.spmv(Kc, x), K @ x) < 1e-13
    free = fo.free_dofs(cons)
    Kff, ff = Kc[free][:, free].tocsr(), forces.flatten()[free]
    u, it, relres = co.jacobi_pcg(Kff, ff, tol=1e-12)
    un, itn, _ = fo.jacobi_pcg(Kff, ff, tol=1e-12)
    # On this symmetric uniform mesh the ITERATION COUNT is decided by rounding noise (it decides when
    # mathematically repeated eigenvalues split): the OpenMP reduction order alone moves it between
    # ~480 and ~520 from run to run.  Both must converge to the same solution; on a generic (jittered)
    # mesh below the counts agree to a few percent (the last iterations towards 1e-12 still feel the
    # summation order: 603..629 against 626 over repeated runs).
    assert relres <= 1e-12 and 0.5 * itn <= it <= 1.5 * itn
    assert rel(u, un) < 1e-8
    Kr = co.reduce_csr(Kc, free)  # the parallel two-pass reduction equals scipy's fancy indexing
    assert np.array_equal(Kr.indptr, Kff.indptr) and np.array_equal(Kr.indices, Kff.indices)
    assert np.array_equal(Kr.data, Kff.data)
    rng = np.random.default_rng(3)
    jn = nodes + rng.uniform(-0.15, 0.15, nodes.shape) * (0.1 / 4) * (nodes[:, 2:3] > 0)
    jf = forces + 0.1 * np.abs(forces).max() * rng.standard_normal(forces.shape)
    Kj = co.reduce_csr(co.assemble_hex8(jn, elements, fo.E_HEX, fo.NU_HEX), free)
    uj, itj, relj = co.jacobi_pcg(Kj, jf.flatten()[free], tol=1e-12)
    unj, itnj, _ = fo.jacobi_pcg(Kj, jf.flatten()[free], tol=1e-12)
    assert relj <= 1e-12 and abs(itj - itnj) <= max(10, itnj // 10) and rel(uj, unj) < 1e-8
    # and against the reference's own displacements (K5 fixture)
    full = np.zeros(K.shape[0])
    full[free] = u
    assert rel(full.reshape(-1, 3), g["displacements"]) < 1e-8
