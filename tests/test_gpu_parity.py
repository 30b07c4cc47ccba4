"""GPU parity: the CUDA path (through the C ABI) against the oracle and the reference-generated
golden fixtures.  Bars (BASELINE.json north_star): CSR pattern and DOF numbering bit-exact;
Ke / K values within 1e-10 relative; displacements within 1e-8 relative at CG residual 1e-12."""
import os

import numpy as np
import pytest
import torch

from oracle import fea_oracle as fo

pytestmark = pytest.mark.gpu

KE_RTOL = 1e-10
U_RTOL = 1e-8


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.fixture(scope="module")
def mods():
    from fea_b200 import core, cubebeam, euler_bernoulli, fea, truss, utils

    return dict(core=core, cubebeam=cubebeam, eb=euler_bernoulli, fea=fea, truss=truss, utils=utils)


# ---------------------------------------------------------------------------------- element Ke
def test_ke_hex8_golden(mods, golden):
    g = golden("hex8_single.npz")
    U = mods["utils"]
    ke = U.hexahedral_stiffness_matrix(g["cube"], 1000, 0.0)
    assert ke.shape == (24, 24) and ke.dtype == np.float64
    assert rel(ke, g["ke_cube"]) < KE_RTOL
    E, nu = float(g["E"]), float(g["nu"])
    for x, k in zip(g["dist_nodes"], g["ke_dist"]):
        assert rel(U.hexahedral_stiffness_matrix(x, E, nu), k) < KE_RTOL
    # K2 round trip through the device Ke
    f = (ke @ g["disp"].flatten()).reshape(-1, 3)
    assert rel(f, g["f_cube"]) < 1e-10


def test_ke_hex8_inverted_raises(mods, golden):
    g = golden("hex8_single.npz")
    with pytest.raises(ValueError, match="Jacobian determinant is non-positive. Check the element shape."):
        mods["utils"].hexahedral_stiffness_matrix(g["inverted"], 1000, 0.0)


def test_ke_hex8_batched_vs_oracle(mods):
    rng = np.random.default_rng(2)
    m = 1003  # not a multiple of 4: exercises the ragged last warp group
    base = fo.HEX8_SIGNS * 0.5
    X = base[None] * rng.uniform(0.5, 2.0, size=(m, 1, 3)) + rng.uniform(-0.2, 0.2, size=(m, 8, 3))
    nodes = X.reshape(-1, 3)
    elements = np.arange(8 * m).reshape(m, 8)
    ke = mods["utils"].hexahedral_stiffness_matrices(nodes, elements, fo.E_HEX, fo.NU_HEX).cpu().numpy()
    ref = fo.hex8_ke_batched(nodes, elements, fo.E_HEX, fo.NU_HEX)
    err = np.abs(ke - ref).reshape(m, -1).max(axis=1) / np.abs(ref).reshape(m, -1).max(axis=1)
    assert err.max() < KE_RTOL
    assert np.abs(ke - ke.transpose(0, 2, 1)).max() / np.abs(ke).max() < 1e-13


def test_ke_empty_batch(mods):
    ke = mods["utils"].hexahedral_stiffness_matrices(np.zeros((8, 3)), np.zeros((0, 8), dtype=np.int64), 1.0, 0.3)
    assert ke.shape == (0, 24, 24)


def test_ke_beam_and_truss(mods, golden):
    g = golden("euler_bernoulli.npz")
    E, I, L, q, n, Le = g["params"]
    ke = mods["eb"].beam_stiffness_matrices(E * I, Le)[0].cpu().numpy()
    assert rel(ke, g["element_stiffness_matrix"]) < KE_RTOL
    rng = np.random.default_rng(3)
    EI, Ls = rng.uniform(1e3, 1e6, 77), rng.uniform(0.01, 2.0, 77)
    assert rel(mods["eb"].beam_stiffness_matrices(EI, Ls).cpu().numpy(), fo.beam_ke_batched(EI, Ls)) < KE_RTOL
    nodes, members, k, _, _ = fo.lattice_truss_case(4, 1)
    kt = mods["truss"].member_stiffness_matrices(nodes, members, k).cpu().numpy()
    assert rel(kt, fo.truss_ke_batched(nodes, members, k)) < KE_RTOL
    # tangent of the reference's compute_forces (finite differences, golden)
    t = golden("truss.npz")
    n64 = t["nodes"].astype(np.float64)
    kk = mods["truss"].member_stiffness_matrices(n64, t["members"], np.full(2, float(t["stiffness"]))).cpu().numpy()
    K = fo.assemble_csr(t["members"], kk, 3, 3).toarray()
    assert np.abs(K - t["tangent_fd"]).max() < 1e-4


# ------------------------------------------------------------------------ symbolic + numeric
def _hex_cases():
    yield "cantilever 6x3x3", fo.cantilever_case(6, 3)
    yield "shipped cubebeam", fo.cubebeam_case()
    yield "shipped tube (periodic section)", fo.tube_case()
    yield "single element", fo.cantilever_case(1, 1)
    # distorted + shuffled element order + an unreferenced node
    nodes, elements, cons, forces = fo.cantilever_case(5, 4)
    rng = np.random.default_rng(7)
    h = 0.1 / 4
    nodes = nodes + rng.uniform(-0.2 * h, 0.2 * h, size=nodes.shape)
    elements = elements[rng.permutation(elements.shape[0])]
    nodes = np.vstack([nodes, [[9.0, 9.0, 9.0]]])
    cons = np.vstack([cons, [[1, 1, 1]]])
    forces = np.vstack([forces, [[0.0, 0.0, 0.0]]])
    yield "distorted, shuffled, isolated node", (nodes, elements, cons, forces)


@pytest.mark.parametrize("name,case", list(_hex_cases()), ids=[n for n, _ in _hex_cases()])
def test_hex8_pattern_and_values(mods, name, case):
    core = mods["core"]
    nodes, elements, cons, forces = case
    nd = core.to_device(nodes, torch.float64)
    el = core.to_device(elements, torch.int32)
    pat = core.symbolic(el, nodes.shape[0])
    rowptr, colidx = (t.cpu().numpy() for t in pat.csr(3))
    Kref = fo.assemble_csr(elements, fo.hex8_ke_batched(nodes, elements, fo.E_HEX, fo.NU_HEX), nodes.shape[0], 3)
    indptr, indices = fo.structural_pattern(elements, nodes.shape[0], 3)
    # bit-exact pattern and DOF numbering
    assert rowptr.dtype == np.int32 and colidx.dtype == np.int32
    assert np.array_equal(rowptr, indptr) and np.array_equal(colidx, indices)
    assert np.array_equal(rowptr, Kref.indptr) and np.array_equal(colidx, Kref.indices)
    K = core.assemble_hex8(nd, el, fo.E_HEX, fo.NU_HEX, pattern=pat)
    vals = K.values.cpu().numpy()
    assert np.abs(vals - Kref.data).max() / np.abs(Kref.data).max() < KE_RTOL
    # deterministic: a second assembly is bit-identical
    K2 = core.assemble_hex8(nd, el, fo.E_HEX, fo.NU_HEX, pattern=pat)
    assert torch.equal(K.values, K2.values)
    # Dirichlet: dinv encodes the constraints; eliminated mode = identity rows/cols
    from fea_b200 import _lib

    fixed = core._fixed_mask(cons, nodes.size)
    Kf = core.assemble_hex8(nd, el, fo.E_HEX, fo.NU_HEX, pattern=pat, fixed=fixed)
    dinv = Kf.dinv.cpu().numpy()
    diag = Kref.diagonal()
    free = fo.free_dofs(cons)
    isfree = np.zeros(nodes.size, bool)
    isfree[free] = True
    isfree &= diag != 0
    assert np.all(dinv[~isfree] == 0.0)
    assert rel(dinv[isfree], 1.0 / diag[isfree]) < KE_RTOL
    Ke = core.assemble_hex8(nd, el, fo.E_HEX, fo.NU_HEX, pattern=pat, fixed=fixed, mode=_lib.ASSEMBLE_ELIMINATED)
    Ke = Ke.to_scipy()
    sub = Ke[free][:, free]
    assert np.abs(sub - Kref[free][:, free]).max() / np.abs(Kref.data).max() < KE_RTOL
    fx = np.setdiff1d(np.arange(nodes.size), free)
    fx = fx[diag[fx] != 0]
    assert np.abs(Ke[fx][:, free]).max() == 0 and np.abs(Ke[free][:, fx]).max() == 0
    assert np.array_equal(Ke[fx][:, fx].diagonal(), np.ones(fx.size))


def test_hex8_affine_pass_vs_general(mods, monkeypatch):
    """Nodes whose incident elements are all exactly affine take the closed-form pass
    (assemble_hex8_affine_kernel), the others the Gauss-point kernel: same matrix to rounding, rows of
    the non-affine part bit-identical whether or not the affine pass ran, and the split is per node
    (a half-jittered mesh exercises affine, mixed and general nodes in one assembly)."""
    core = mods["core"]
    nodes, elements, cons, forces = fo.cantilever_case(12, 5)
    rng = np.random.default_rng(11)
    h = 0.1 / 5
    moved = nodes[:, 2] > 0.55
    nodes = nodes.copy()
    nodes[moved] += rng.uniform(-0.2 * h, 0.2 * h, size=(int(moved.sum()), 3))
    nd, el = core.to_device(nodes, torch.float64), core.to_device(elements, torch.int32)
    pat = core.symbolic(el, nodes.shape[0])
    fixed = core._fixed_mask(cons, nodes.size)
    Kref = fo.assemble_csr(elements, fo.hex8_ke_batched(nodes, elements, fo.E_HEX, fo.NU_HEX), nodes.shape[0], 3)
    out = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("FEA_ASSEMBLE_AFFINE", flag)
        K = core.assemble_hex8(nd, el, fo.E_HEX, fo.NU_HEX, pattern=pat, fixed=fixed)
        out[flag] = (K.values.cpu().numpy(), K.dinv.cpu().numpy())
        assert np.abs(out[flag][0] - Kref.data).max() / np.abs(Kref.data).max() < KE_RTOL
    v1, v0 = out["1"][0], out["0"][0]
    assert np.abs(v1 - v0).max() / np.abs(v0).max() < 1e-14
    assert rel(out["1"][1], out["0"][1]) < 1e-13
    # rows of nodes with a jittered incident element: untouched by the affine pass, bit for bit
    touched = np.zeros(nodes.shape[0], bool)
    touched[elements[moved[elements].any(axis=1)].ravel()] = True
    rp = pat.node_rowptr.cpu().numpy().astype(np.int64)
    general_rows = np.concatenate([np.arange(9 * rp[i], 9 * rp[i + 1]) for i in np.flatnonzero(touched)])
    assert np.array_equal(v1[general_rows], v0[general_rows])
    assert 0 < touched.sum() < nodes.shape[0] and not np.array_equal(v1, v0)


def test_inverted_element_in_assembly_raises(mods):
    nodes, elements, cons, forces = fo.cantilever_case(3, 2)
    elements = elements.copy()
    elements[5] = elements[5][[4, 5, 6, 7, 0, 1, 2, 3]]
    with pytest.raises(ValueError, match="Jacobian determinant is non-positive") as exc:
        mods["cubebeam"].solve(nodes, elements, cons, forces)
    assert exc.value.element == 5


def test_bad_connectivity_rejected(mods):
    nodes, elements, cons, forces = fo.cantilever_case(2, 1)
    elements = elements.copy()
    elements[0, 0] = nodes.shape[0] + 5
    with pytest.raises(ValueError):
        mods["cubebeam"].solve(nodes, elements, cons, forces)


def test_beam_and_truss_assembly(mods):
    core = mods["core"]
    elements, EI, Ls, cons, loads = fo.beam_cantilever_case(257)
    rng = np.random.default_rng(4)
    EI = EI * rng.uniform(0.5, 2.0, EI.shape)
    Ls = Ls * rng.uniform(0.5, 2.0, Ls.shape)
    K = core.assemble_beam(core.to_device(EI, torch.float64), core.to_device(Ls, torch.float64),
                           core.to_device(elements, torch.int32), 258)
    Kref = fo.assemble_csr(elements, fo.beam_ke_batched(EI, Ls), 258, 2)
    rp, ci = (t.cpu().numpy() for t in K.pattern.csr(2))
    assert np.array_equal(rp, Kref.indptr) and np.array_equal(ci, Kref.indices)
    assert rel(K.values.cpu().numpy(), Kref.data) < KE_RTOL
    assert K.nnz == 4 * (3 * 258 - 2)

    nodes, members, k, cons, loads = fo.lattice_truss_case(6, 2)
    perm = np.random.default_rng(5).permutation(members.shape[0])
    members, k = members[perm], k[perm]
    K = core.assemble_truss(core.to_device(nodes, torch.float64), core.to_device(members, torch.int32),
                            core.to_device(k, torch.float64))
    Kref = fo.assemble_csr(members, fo.truss_ke_batched(nodes, members, k), nodes.shape[0], 3)
    rp, ci = (t.cpu().numpy() for t in K.pattern.csr(3))
    assert np.array_equal(rp, Kref.indptr) and np.array_equal(ci, Kref.indices)
    assert np.abs(K.values.cpu().numpy() - Kref.data).max() / np.abs(Kref.data).max() < KE_RTOL


# --------------------------------------------------------------------------------- SpMV / PCG
def test_spmv_spmm_vs_scipy(mods):
    core = mods["core"]
    nodes, elements, cons, forces = fo.cantilever_case(9, 4)
    K = core.assemble_hex8(core.to_device(nodes, torch.float64), core.to_device(elements, torch.int32), 1.0, 0.3)
    Kref = K.to_scipy()
    rng = np.random.default_rng(6)
    x = rng.standard_normal(K.n_dof)
    y = K.matvec(core.to_device(x, torch.float64)).cpu().numpy()
    assert rel(y, Kref @ x) < 1e-13
    for r in (1, 3, 32, 64, 70):
        X = rng.standard_normal((K.n_dof, r))
        Y = K.matmat(core.to_device(X, torch.float64)).cpu().numpy()
        assert rel(Y, Kref @ X) < 1e-13
    # D = 1: the same kernel as a plain CSR SpMV on the DOF-level arrays
    from fea_b200 import _lib

    rowptr, colidx = K.pattern.csr(3)
    y1 = torch.empty(K.n_dof, dtype=torch.float64, device="cuda")
    xd = core.to_device(x, torch.float64)
    for maxc in (0, 3 * K.pattern.max_coupled):  # 0 = generic kernel, else the bulk-copy pipeline
        rc = _lib.load().fea_spmv(K.n_dof, 1, rowptr.data_ptr(), colidx.data_ptr(), K.values.data_ptr(), maxc,
                                  xd.data_ptr(), y1.data_ptr(), None)
        assert rc == 0
        assert rel(y1.cpu().numpy(), Kref @ x) < 1e-13
    # D = 3, generic kernel (max_coupled unknown) against the pipeline kernel
    pt = K.pattern
    rc = _lib.load().fea_spmv(pt.n_nodes, 3, pt.node_rowptr.data_ptr(), pt.node_colidx.data_ptr(),
                              K.values.data_ptr(), 0, xd.data_ptr(), y1.data_ptr(), None)
    assert rc == 0 and rel(y1.cpu().numpy(), Kref @ x) < 1e-13
    # rigid translations are in the null space of the unconstrained K (row sums vanish)
    t = np.tile([1.0, 0.0, 0.0], nodes.shape[0])
    assert np.abs(K.matvec(core.to_device(t, torch.float64)).cpu().numpy()).max() < 1e-12 * np.abs(Kref.data).max()


def test_pcg_vs_oracle(mods, monkeypatch):
    core = mods["core"]
    monkeypatch.setenv("FEA_PCG_ALGO", "0")  # the classical recurrence is the one the oracle states
    nodes, elements, cons, forces = fo.cantilever_case(20, 4)
    nd, el = core.to_device(nodes, torch.float64), core.to_device(elements, torch.int32)
    fixed = core._fixed_mask(cons, nodes.size)
    K = core.assemble_hex8(nd, el, fo.E_HEX, fo.NU_HEX, fixed=fixed)
    b = core.to_device(forces, torch.float64).reshape(-1)
    uo, fo_, io = fo.solve_hex8(nodes, elements, cons, forces, method="pcg", tol=1e-12)
    ud, _, _ = fo.solve_hex8(nodes, elements, cons, forces, method="direct")
    # (1) the solver on the GPU-assembled K: converges to the reference solution
    u_own, info_own = core.pcg(K, b, tol=1e-12)
    assert info_own.status == 0 and info_own.rel_residual <= 1e-12
    assert rel(u_own.cpu().numpy(), ud.ravel()) < U_RTOL
    # (2) the recurrence itself, on the ORACLE's matrix values (same CSR order): the iteration count of
    # CG on this symmetric uniform mesh depends on the rounding noise in K (it decides when the
    # mathematically repeated eigenvalues split), so the histories are compared on identical values.
    Ko = io["K"]
    assert np.array_equal(Ko.indptr, K.pattern.csr(3)[0].cpu().numpy())
    K.values.copy_(torch.from_numpy(Ko.data))
    diag = torch.from_numpy(Ko.diagonal()).to(K.values.device)
    K.dinv.copy_(torch.where(fixed != 0, torch.zeros_like(diag), 1.0 / diag))
    u, info = core.pcg(K, b, tol=1e-12, history=True)
    assert info.status == 0 and info.rel_residual <= 1e-12
    assert abs(info.iterations - io["iterations"]) <= max(5, io["iterations"] // 50)
    assert rel(u.cpu().numpy(), uo.ravel()) < U_RTOL
    assert rel(u.cpu().numpy(), ud.ravel()) < U_RTOL
    # same recurrence: early residual history tracks the oracle's
    h, ho = info.history, np.array(io["history"])
    k = min(40, len(h), len(ho))
    assert np.allclose(h[:k], ho[:k], rtol=1e-6)
    assert np.all(u.cpu().numpy()[cons.ravel() != 0] == 0.0)


def test_pcg_single_reduction_variant(mods, monkeypatch):
    """FEA_PCG_ALGO=1 (Chronopoulos-Gear: 2 kernels, 1 reduction per iteration) against the classical
    recurrence: same iterates in exact arithmetic -> same solution (1e-8), iteration count within a
    few, residual history tracking; max_iter, zero right-hand side and history bookkeeping."""
    core = mods["core"]
    nodes, elements, cons, forces = fo.cantilever_case(20, 4)
    # jittered interior nodes and a random load: on the symmetric uniform mesh the iteration count is
    # decided by rounding noise (311 / 417 / 420 for three roundings of the same K, see
    # test_pcg_vs_oracle); a generic mesh makes the two recurrences comparable iteration by iteration
    rng = np.random.default_rng(3)
    h = 0.1 / 4
    nodes = nodes + rng.uniform(-0.15 * h, 0.15 * h, nodes.shape) * (nodes[:, 2:3] > 0)
    forces = forces + 0.1 * np.abs(forces).max() * rng.standard_normal(forces.shape)
    nd, el = core.to_device(nodes, torch.float64), core.to_device(elements, torch.int32)
    K = core.assemble_hex8(nd, el, fo.E_HEX, fo.NU_HEX, fixed=core._fixed_mask(cons, nodes.size))
    b = core.to_device(forces, torch.float64).reshape(-1)
    monkeypatch.setenv("FEA_PCG_ALGO", "0")
    u0, i0 = core.pcg(K, b, tol=1e-12, history=True)
    monkeypatch.setenv("FEA_PCG_ALGO", "1")
    u1, i1 = core.pcg(K, b, tol=1e-12, history=True)
    assert i1.status == 0 and i1.rel_residual <= 1e-12
    assert abs(i1.iterations - i0.iterations) <= max(5, i0.iterations // 50)
    assert rel(u1.cpu().numpy(), u0.cpu().numpy()) < U_RTOL
    k = min(40, len(i0.history), len(i1.history))
    assert np.allclose(i1.history[:k], i0.history[:k], rtol=1e-6)
    assert len(i1.history) == i1.iterations and abs(i1.history[-1] - i1.rel_residual) < 1e-15
    assert np.all(u1.cpu().numpy()[cons.ravel() != 0] == 0.0)
    # iteration cap: exactly max_iter iterations, FEA_ERR_MAXITER
    u2, i2 = core.pcg(K, b, tol=1e-12, max_iter=37, raise_on_failure=False)
    assert i2.iterations == 37 and i2.status != 0
    # the capped iterate equals the classical one after the same number of iterations
    monkeypatch.setenv("FEA_PCG_ALGO", "0")
    u3, i3 = core.pcg(K, b, tol=1e-12, max_iter=37, raise_on_failure=False)
    assert i3.iterations == 37 and rel(u2.cpu().numpy(), u3.cpu().numpy()) < 1e-9
    monkeypatch.setenv("FEA_PCG_ALGO", "1")
    uz, iz = core.pcg(K, torch.zeros_like(b), tol=1e-12)
    assert iz.status == 0 and iz.iterations == 0 and float(uz.abs().max()) == 0.0
    # end to end through the reference-facing call
    ud, _, _ = fo.solve_hex8(nodes, elements, cons, forces, method="direct")
    u_api, f_api = mods["cubebeam"].solve(nodes, elements, cons, forces)
    assert rel(u_api.ravel(), ud.ravel()) < U_RTOL


def test_pcg_persistent_kernel_vs_two_kernel_path(mods, monkeypatch):
    """Below 3 M DOF fea_pcg_solve runs the single-reduction recurrence in ONE persistent kernel per 128
    iterations (pcg_fused.cuh: grid barriers, every CTA reduces the partial sums itself).  Against the
    two-kernel path (FEA_PCG_FUSED=0) on the same matrix: same recurrence, another order of the partial sums
    -> solution 1e-10, crossing of 1e-8 within 2 % of the iterations, history of the first 100 iterations 1e-6; iteration caps on
    and around a launch boundary; residual history across launches."""
    core = mods["core"]
    nodes, elements, cons, forces = fo.cantilever_case(30, 6)
    rng = np.random.default_rng(5)
    h = 0.1 / 6
    nodes = nodes + rng.uniform(-0.15 * h, 0.15 * h, nodes.shape) * (nodes[:, 2:3] > 0)
    forces = forces + 0.1 * np.abs(forces).max() * rng.standard_normal(forces.shape)
    nd, el = core.to_device(nodes, torch.float64), core.to_device(elements, torch.int32)
    K = core.assemble_hex8(nd, el, fo.E_HEX, fo.NU_HEX, fixed=core._fixed_mask(cons, nodes.size))
    b = core.to_device(forces, torch.float64).reshape(-1)
    monkeypatch.setenv("FEA_PCG_ALGO", "1")
    monkeypatch.setenv("FEA_PCG_FUSED", "0")
    u0, i0 = core.pcg(K, b, tol=1e-12, history=True)
    monkeypatch.setenv("FEA_PCG_FUSED", "1")
    u1, i1 = core.pcg(K, b, tol=1e-12, history=True)
    assert i0.status == 0 and i1.status == 0 and i1.rel_residual <= 1e-12
    assert i0.iterations > 300  # several launches of the persistent kernel
    # the recurrence residual creeps under 1e-12 at the end (slope ~5 % per iteration): where exactly it crosses
    # depends on the rounding of the dot products; the crossing of 1e-8 is sharp
    first_below = lambda hist: int(np.argmax(np.asarray(hist) < 1e-8))
    assert abs(first_below(i1.history) - first_below(i0.history)) <= max(3, first_below(i0.history) // 50)
    assert abs(i1.iterations - i0.iterations) <= i0.iterations // 6
    assert rel(u1.cpu().numpy(), u0.cpu().numpy()) < 1e-10
    k = min(100, len(i0.history), len(i1.history))
    assert np.allclose(i1.history[:k], i0.history[:k], rtol=1e-6)
    assert len(i1.history) == i1.iterations and abs(i1.history[-1] - i1.rel_residual) < 1e-15
    # across the launch boundary at 128 (by iteration 200 the two roundings have drifted 1 % apart on this case)
    assert np.allclose(i1.history[100:140], i0.history[100:140], rtol=1e-4)
    u1b, i1b = core.pcg(K, b, tol=1e-12)
    assert torch.equal(u1, u1b) and i1b.iterations == i1.iterations  # deterministic
    for cap in (1, 64, 127, 128, 129):  # (later the two roundings of this case drift apart: 1 % by iteration 200)
        monkeypatch.setenv("FEA_PCG_FUSED", "1")
        uc, ic = core.pcg(K, b, tol=1e-12, max_iter=cap, raise_on_failure=False)
        monkeypatch.setenv("FEA_PCG_FUSED", "0")
        ud, idd = core.pcg(K, b, tol=1e-12, max_iter=cap, raise_on_failure=False)
        assert ic.iterations == cap == idd.iterations and ic.status == idd.status != 0
        assert rel(uc.cpu().numpy(), ud.cpu().numpy()) < (1e-9 if cap <= 64 else 1e-5)


def test_pcg_zero_rhs_and_singular(mods):
    core = mods["core"]
    nodes, elements, cons, forces = fo.cantilever_case(3, 2)
    u, f = mods["cubebeam"].solve(nodes, elements, cons, np.zeros_like(forces))
    assert np.all(u == 0) and np.all(f == 0)
    # unconstrained body: the reference's np.linalg.solve raises LinAlgError (singular K)
    with pytest.raises(np.linalg.LinAlgError):
        from fea_b200 import model

        model.solve_hex8(nodes, elements, np.zeros_like(cons), forces, max_iter=2000)


def test_pcg_stagnation_exit(mods):
    """An under-constrained body with the DEFAULT max_iter (10 n): the residual stops improving and the
    solve ends with FEA_ERR_STAGNATION -> LinAlgError after ~10 k iterations instead of 10 n."""
    from fea_b200 import _lib

    core = mods["core"]
    nodes, elements, cons, forces = fo.cantilever_case(40, 8)  # 9,963 DOF: max_iter = 99,630
    nd, el = core.to_device(nodes, torch.float64), core.to_device(elements, torch.int32)
    b = core.to_device(forces, torch.float64).reshape(-1)
    K = core.assemble_hex8(nd, el, fo.E_HEX, fo.NU_HEX, fixed=core._fixed_mask(np.zeros_like(cons), nodes.size))
    for algo in ("0", "1"):
        os.environ["FEA_PCG_ALGO"] = algo
        try:
            _, info = core.pcg(K, b, raise_on_failure=False)
        finally:
            del os.environ["FEA_PCG_ALGO"]
        assert info.status in (_lib.FEA_ERR_STAGNATION, _lib.FEA_ERR_BREAKDOWN)
        assert info.iterations < 40_000
    with pytest.raises(np.linalg.LinAlgError):
        core.pcg(K, b)
    # a constrained solve of the same mesh is untouched by the guard
    Kc = core.assemble_hex8(nd, el, fo.E_HEX, fo.NU_HEX, fixed=core._fixed_mask(cons, nodes.size))
    _, info = core.pcg(Kc, b)
    assert info.status == 0 and info.rel_residual <= 1e-12


# ---------------------------------------------------------------- end-to-end, reference surface
def test_k5_cubebeam_shipped(mods, golden):
    g = golden("cubebeam.npz")
    nodes, elements, cons, forces = mods["cubebeam"].shipped_case()
    before = (nodes.copy(), elements.copy(), cons.copy(), forces.copy())
    u, f = mods["cubebeam"].solve(nodes, elements, cons, forces)
    assert u.shape == nodes.shape and f.shape == nodes.shape and u.dtype == np.float64
    for a, b in zip((nodes, elements, cons, forces), before):
        assert np.array_equal(a, b)  # inputs untouched (cubebeam.py:102-108)
    assert rel(u, g["displacements"]) < U_RTOL
    assert rel(f, g["forces_out"]) < 1e-7
    assert abs(np.abs(u).max() - 3.050405508343811e-04) < 1e-11
    assert abs(f[nodes[:, 2] == 0][:, 1].sum() + 1402.158792651803) < 1e-5


def test_k6_tube_shipped(mods, golden):
    g = golden("fea_tube.npz")
    nodes, elements, cons, forces = mods["fea"].shipped_case()
    u, f = mods["fea"].solve(nodes, elements, cons, forces)
    assert rel(u, g["displacements"]) < U_RTOL
    assert rel(f, g["forces_out"]) < 1e-7


def test_k7_euler_bernoulli(mods, golden):
    g = golden("euler_bernoulli.npz")
    eb = mods["eb"]
    res = eb.run()
    assert rel(res["global_stiffness_matrix"], g["global_stiffness_matrix"]) < KE_RTOL
    assert np.array_equal(res["load_vector"], g["load_vector"])
    assert res["fixed_dofs"] == list(g["fixed_dofs"]) and res["free_dofs"] == list(g["free_dofs"])
    assert rel(res["displacement_vector"], g["displacement_vector"]) < U_RTOL  # direct (cyclic reduction) solve
    m, v = eb.moment_shear(g["displacement_vector"], np.full(100, eb.E * eb.I), np.full(100, eb.element_length))
    assert rel(m, g["moment_vector"]) < 1e-12 and rel(v, g["shear_vector"]) < 1e-12
    assert eb.displacement_vector is res["displacement_vector"]  # lazy module attribute
    # cantilever, n = 100: against the oracle's dense solve (np.linalg.solve, the reference's own solver)
    # and the analytic deflection; FP64 elimination, double-double elimination and Jacobi-PCG
    el, EI, Ls, cons, loads = eb.cantilever_case(100)
    uo, _, _ = fo.solve_beam(el, EI, Ls, cons, loads, method="dense")
    x = np.linspace(0, 1, 101)
    w = -1000.0 * x**2 * (3 - x) / (6 * 210e9 * 1e-6)
    for kw in (dict(extended=False), dict(extended=True)):
        u = eb.solve_beam(el, EI, Ls, cons, loads, **kw)
        assert rel(u, uo) < U_RTOL and rel(u[:, 0], w) < U_RTOL
        assert u[0, 0] == 0.0 and u[0, 1] == 0.0
    u = eb.solve_beam(el, EI, Ls, cons, loads, method="pcg")
    assert rel(u, uo) < 1e-6  # cond(K) ~ 5e8: what Jacobi-PCG at a 1e-12 residual leaves


def uo_K(fo_, el, EI, Ls):
    """The oracle's own assembled beam matrix (what the reference's dense solve sees)."""
    return fo_.assemble_csr(el, fo_.beam_ke_batched(EI, Ls), el.shape[0] + 1, 2)


def test_beam_chain_solver_at_size(mods):
    """fea_chain_solve (block-tridiagonal parallel cyclic reduction) on the Euler-Bernoulli cantilever up
    to BASELINE config 2's 100,000 elements.  cond(K) ~ 5 n^4: 5e12 at n = 1000, 5e20 at 100 k (SURVEY.md
    H3).  Two separate error sources are checked separately:
      * the SOLVER: the double-double elimination reproduces the exact (60-digit decimal) solution of the
        very FP64 matrix it was given to 1e-12, where FP64 elimination -- ours or the reference's LAPACK --
        is off by cond x eps;
      * the MATRIX: what is left against the analytic deflection P x^2 (3L - x) / 6EI (Hermite elements are
        nodally exact for a tip load) is the FP64 rounding of the assembled entries, ~1e-5 at 100 k elements,
        where FP64 LU (LAPACK, SuperLU) returns noise."""
    eb, core = mods["eb"], mods["core"]
    errs = {}
    for n in (1000, 4000, 100_000):
        el, EI, Ls, cons, loads = eb.cantilever_case(n)
        x = np.linspace(0, 1, n + 1)
        w = -1000.0 * x**2 * (3 - x) / (6 * 210e9 * 1e-6)
        u64 = eb.solve_beam(el, EI, Ls, cons, loads, extended=False)
        udd, K, info = eb.solve_beam(el, EI, Ls, cons, loads, extended=True, return_matrix=True)
        errs[n] = (rel(u64[:, 0], w), rel(udd[:, 0], w))
        assert udd[0, 0] == 0.0 and udd[0, 1] == 0.0 and np.all(np.isfinite(udd))
        # the assembled matrix equals the oracle's BIT FOR BIT (EI / L**3 with a once-rounded cube, like the
        # reference's pow): at cond ~ 5 n^4 the last bit of the entries is visible in the solution
        assert np.array_equal(K.values.cpu().numpy(), uo_K(fo, el, EI, Ls).data)
        if n <= 4000:
            exact = fo.chain_solve_exact(K.to_scipy(), cons, loads)  # of the GPU-assembled FP64 matrix itself
            assert rel(udd, exact) < 1e-12
            e64 = rel(u64, exact)
            assert e64 < 1e-16 * 5 * float(n) ** 4  # FP64 elimination: within cond x eps of it
            if n == 1000:
                uo, _, _ = fo.solve_beam(el, EI, Ls, cons, loads, method="dense")  # the reference's solver
                assert e64 < 10 * max(rel(uo, fo.chain_solve_exact(uo_K(fo, el, EI, Ls), cons, loads)), 1e-9)
    assert errs[1000][1] < 1e-6 and errs[4000][1] < 1e-5 and errs[100_000][1] < 1e-3, errs
    # not a chain -> ValueError (an unconstrained beam behaves like LAPACK on a numerically singular K: no
    # exactly zero pivot, huge displacements, no exception)
    nodes, elements, hc, hf = fo.cantilever_case(3, 2)
    Kh = core.assemble_hex8(core.to_device(nodes, torch.float64), core.to_device(elements, torch.int32), 1.0, 0.3)
    with pytest.raises(ValueError):
        core.chain_solve(Kh, torch.zeros(Kh.n_dof, dtype=torch.float64, device="cuda"))


def test_k8_truss(mods, golden):
    g = golden("truss.npz")
    T = mods["truss"]
    n64 = g["nodes"].astype(np.float64)
    f = np.zeros_like(n64)
    assert T.compute_forces(n64, g["members"], g["probe"], f) is None
    assert rel(f, g["f_probe"]) < 1e-13
    f += 1.0  # accumulates in place
    T.compute_forces(n64, g["members"], g["probe"], f)
    assert rel(f, 2 * g["f_probe"] + 1.0) < 1e-13
    # the script's loop for 40 passes, on the device, in the script's own float32 (truss.py:9-10)
    disp, hist = T.relax(T.nodes, T.members, T.loads, 40)
    assert disp.dtype == np.float32
    # float32 arithmetic like the script's; the operation order inside a norm differs (numpy's sdot vs
    # fused multiply-adds), and below ~3e-4 the script's residual is float32 noise (it "stalls at ~1e-4",
    # SURVEY.md K8)
    assert np.allclose(hist, g["residual_history"], rtol=1e-4, atol=3e-4)
    assert np.allclose(disp, g["displaced_history"][-1], atol=2e-6)
    # ... and in FP64 against the oracle's FP64 relaxation
    d64, h64 = T.relax(n64, g["members"], [[int(g["load_node"]), g["load"].astype(np.float64)]], 40)
    od = n64.copy()
    oh = [fo.truss_relax_step(n64, g["members"], od, [[int(g["load_node"]), g["load"].astype(np.float64)]],
                              float(g["stiffness"])) for _ in range(40)]
    assert d64.dtype == np.float64 and np.allclose(h64, oh, rtol=1e-12, atol=1e-12) and rel(d64, od) < 1e-13
    u, K, info = T.solve_linear(*T.shipped_case(), return_matrix=True)
    assert np.allclose(u[2], [0.0, -0.25, 0.0], atol=1e-14)


def test_truss_relax_device_lattice(mods):
    """The relaxation loop on a lattice with many loaded nodes (one in five), per-member spring rates:
    25 passes on the device against the oracle's loop; the member-force evaluation is deterministic
    (node-owner gather in member order, no atomics): two runs are bit-identical."""
    T = mods["truss"]
    nodes, members, k, cons, _ = T.lattice_truss(7, n_rhs=1)
    rng = np.random.default_rng(5)
    loaded = np.sort(rng.choice(np.arange(49, nodes.shape[0]), nodes.shape[0] // 5, replace=False))
    loads = [[int(i), 20.0 * rng.standard_normal(3)] for i in loaded]
    d1, h1, r1 = T.relax(nodes, members, loads, 25, member_stiffness=k, return_residual=True)
    d2, h2, r2 = T.relax(nodes, members, loads, 25, member_stiffness=k, return_residual=True)
    assert np.array_equal(d1, d2) and np.array_equal(h1, h2) and np.array_equal(r1, r2)
    # the oracle's loop with per-member k in the forces and the module constant in the update
    od = nodes.copy()
    oh = []
    for _ in range(25):
        f = np.zeros_like(nodes)
        fo.truss_compute_forces(nodes, members, od, f, k)
        oh.append(float(np.linalg.norm(loads[0][1] + f[loads[0][0]])))
        for i, load in loads:
            od[i] += (load + f[i]) / T.stiffness
    assert np.allclose(h1, oh, rtol=1e-11) and rel(d1, od) < 1e-12
    f = np.zeros_like(nodes)
    T.compute_forces(nodes, members, d1, f, member_stiffness=k)
    fo_ = np.zeros_like(nodes)
    fo.truss_compute_forces(nodes, members, od, fo_, k)
    assert rel(f, fo_) < 1e-9
    with pytest.raises(ValueError):
        T.relax(nodes, members, [loads[0], loads[0]], 2)


def test_device_mesh_builders(mods):
    """the input generators on the device equal the host builders element for element
    (quad grid incl. np.linspace's rounding, extrusion, tube connectivity, config-5 lattice incl. numpy's
    PCG64 stream); tube coordinates agree to the last bit or two (device cos / sin)."""
    C, F, T, U, core = mods["cubebeam"], mods["fea"], mods["truss"], mods["utils"], mods["core"]
    for nx, ny, w, h in ((4, 4, 0.1, 0.1), (7, 3, 0.3, 0.2), (80, 80, 0.1, 0.1), (1, 1, 1.0, 2.0), (33, 129, 0.7, 1.3)):
        n2, q2 = C.generate_quad_grid(nx, ny, w, h)
        dn, dq = C.generate_quad_grid_device(nx, ny, w, h)
        assert np.array_equal(dn.cpu().numpy(), n2) and np.array_equal(dq.cpu().numpy(), q2)
    nodes, elements, cons, forces = C.cantilever_case(37, 9)
    dn, de, dfix, dload = C.cantilever_case_device(37, 9)
    assert np.array_equal(dn.cpu().numpy(), nodes) and np.array_equal(de.cpu().numpy(), elements)
    assert np.array_equal(dfix.cpu().numpy(), (cons.ravel() != 0).astype(np.uint8))
    assert np.array_equal(dload.cpu().numpy(), forces.ravel())
    for n_seg in (26, 7, 360):
        n2, q2, _ = F.tube_section(n_seg)
        dn, dq = F.tube_section_device(n_seg)
        assert np.array_equal(dq.cpu().numpy(), q2)
        assert np.abs(dn.cpu().numpy() - n2).max() <= 4e-16 * F.outer_radius
    for n in (2, 5, 12):
        pts, mem, k, cons, _ = T.lattice_truss(n, n_rhs=1)
        dp, dm, dk, dc = T.lattice_truss_device(n)
        assert dm.shape[0] == 3 * n * n * (n - 1) + 6 * n * (n - 1) ** 2 + 4 * (n - 1) ** 3
        assert np.array_equal(dp.cpu().numpy(), pts) and np.array_equal(dm.cpu().numpy(), mem)
        assert np.array_equal(dk.cpu().numpy(), k) and np.array_equal(dc.cpu().numpy() != 0, cons != 0)
    pts, mem, k, cons, _ = T.lattice_truss(9, n_rhs=1, h=0.37)
    dp, dm, dk, dc = T.lattice_truss_device(9, h=0.37)
    assert np.array_equal(dp.cpu().numpy(), pts) and np.array_equal(dk.cpu().numpy(), k)


def test_truss_lattice_multi_rhs(mods):
    T = mods["truss"]
    nodes, members, k, cons, loads = fo.lattice_truss_case(7, 64)
    X, K, info = T.solve_linear(nodes, members, k, cons, loads, return_matrix=True)
    Kref = fo.assemble_csr(members, fo.truss_ke_batched(nodes, members, k), nodes.shape[0], 3)
    free = fo.free_dofs(cons)
    import scipy.sparse.linalg as spla

    Xd = spla.splu(Kref[free][:, free].tocsc()).solve(loads[free])
    assert rel(X[free], Xd) < U_RTOL
    assert np.all(X[np.setdiff1d(np.arange(loads.shape[0]), free)] == 0)
    Xo, iters = fo.jacobi_pcg_multi(Kref[free][:, free].tocsr(), loads[free], tol=1e-12)
    assert np.abs(info.history - iters).max() <= 5  # per-column iteration counts track the oracle
    assert info.rel_residual <= 1e-12
    # single RHS through the same entry point
    u1 = T.solve_linear(nodes, members, k, cons, loads[:, 0].reshape(-1, 3))
    assert rel(u1.ravel(), X[:, 0]) < 1e-9


def test_config5_n30_converged_vs_sparse_lu(mods, golden):
    """Config-5-shaped lattice truss at n = 30 (27,000 nodes, 78,300 free DOF, 64 load cases), batched
    PCG run to CONVERGENCE on every column, against scipy's sparse LU of the oracle's reduced matrix
    (recorded by oracle/make_golden_large.py config5: a seeded 4000-row sample of X, all 64 columns, and
    the column norms of the full solution)."""
    T = mods["truss"]
    g = golden("oracle_config5_n30.npz")
    n = int(g["n"])
    nodes, members, k, cons, loads = T.lattice_truss(n, n_rhs=64)
    on, om, ok_, oc, ol = fo.lattice_truss_case(n, n_rhs=64)[:5]  # the product's generator == the oracle's
    assert np.array_equal(nodes, on) and np.array_equal(members, om) and np.array_equal(k, ok_)
    assert np.array_equal(loads, ol)
    X, K, info = T.solve_linear(nodes, members, k, cons, loads, return_matrix=True)
    assert info.status == 0 and info.rel_residual <= 1e-12 and len(info.history) == 64
    X = X if isinstance(X, np.ndarray) else X.cpu().numpy()
    scale = g["col_max"][None, :]
    assert (np.abs(X[g["rows"]] - g["X_rows"]) / scale).max() < U_RTOL
    assert np.abs(np.linalg.norm(X, axis=0) / g["col_norms"] - 1.0).max() < U_RTOL
    assert np.all(X[cons.ravel() != 0] == 0.0)


def test_spmm_and_multi_rhs_wide_rows_and_odd_widths(mods):
    """Hub-and-spoke truss: the hub couples to 70 nodes (> one 32-node staging round of the SpMM),
    leaves couple to ~8.  SpMM and the batched PCG at odd, even, multi-tile RHS counts (scalar and
    16-byte vector paths) against scipy / sparse LU."""
    import scipy.sparse.linalg as spla

    core, T = mods["core"], mods["truss"]
    rng = np.random.default_rng(11)
    n_leaf = 70
    leaf = rng.standard_normal((n_leaf, 3))
    leaf /= np.linalg.norm(leaf, axis=1)[:, None]
    pts = np.concatenate([[[0.0, 0.0, 0.0]], leaf * (1 + 0.2 * rng.random((n_leaf, 1)))])
    dist = np.linalg.norm(leaf[:, None] - leaf[None], axis=2)
    np.fill_diagonal(dist, 9.0)
    near = np.argsort(dist, axis=1)[:, :6]  # shell: every leaf to its 6 nearest leaves (cond(K_ff) ~ 8e2)
    pairs = sorted({(min(i, j) + 1, max(i, j) + 1) for i in range(n_leaf) for j in near[i]})
    members = np.concatenate([np.stack([np.zeros(n_leaf, dtype=np.int64), np.arange(1, n_leaf + 1)], axis=1),
                              np.array(pairs)])
    k = rng.uniform(500.0, 1500.0, members.shape[0])
    cons = np.zeros((n_leaf + 1, 3), dtype=int)
    cons[1:8] = 1
    K = core.assemble_truss(core.to_device(pts, torch.float64), core.to_device(members, torch.int32),
                            core.to_device(k, torch.float64), fixed=core._fixed_mask(cons, pts.size))
    assert K.pattern.max_coupled == n_leaf + 1
    Kref = fo.assemble_csr(members, fo.truss_ke_batched(pts, members, k), pts.shape[0], 3)
    free = fo.free_dofs(cons)
    lu = spla.splu(Kref[free][:, free].tocsc())
    for r in (1, 5, 64, 130):
        X = rng.standard_normal((K.n_dof, r))
        Y = K.matmat(core.to_device(X, torch.float64)).cpu().numpy()
        assert rel(Y, Kref @ X) < 1e-13
        U, info = core.pcg_multi(K, core.to_device(X, torch.float64), tol=1e-12)
        U = U.cpu().numpy()
        assert info.status == 0 and info.rel_residual <= 1e-12
        assert rel(U[free], lu.solve(X[free])) < U_RTOL


def test_spmm_and_multi_rhs_other_block_sizes(mods):
    """The SpMM / batched-PCG kernels at D = 2 (beam blocks) and D = 1 (plain DOF-level CSR), odd and
    even RHS counts: the value-slab layout and the union gather are templated on D."""
    import scipy.sparse.linalg as spla

    from fea_b200 import _lib

    core, EB = mods["core"], mods["eb"]
    rng = np.random.default_rng(5)
    n = 40
    elements, EI, Ls, cons, loads = EB.cantilever_case(n)
    EI = EI * rng.uniform(0.5, 1.5, n)
    fixed = core._fixed_mask(cons, 2 * (n + 1))
    K = core.assemble_beam(core.to_device(EI, torch.float64), core.to_device(Ls, torch.float64),
                           core.to_device(elements, torch.int32), n + 1, fixed=fixed)
    Kref = K.to_scipy()
    free = fo.free_dofs(cons)
    lu = spla.splu(Kref[free][:, free].tocsc())
    for r in (1, 3, 8):
        X = rng.standard_normal((K.n_dof, r))
        Xd = core.to_device(X, torch.float64)
        assert rel(K.matmat(Xd).cpu().numpy(), Kref @ X) < 1e-13
        U, info = core.pcg_multi(K, Xd, tol=1e-12)
        assert info.status == 0 and info.rel_residual <= 1e-12
        assert rel(U.cpu().numpy()[free], lu.solve(X[free])) < 1e-5  # cond(K_ff) ~ 1e7
    # D = 1: the hex8 matrix through its DOF-level arrays
    nodes, els, _, _ = fo.cantilever_case(5, 3)
    Kh = core.assemble_hex8(core.to_device(nodes, torch.float64), core.to_device(els, torch.int32), 1.0, 0.3)
    Khref = Kh.to_scipy()
    rowptr, colidx = Kh.pattern.csr(3)
    for r in (2, 5):
        X = rng.standard_normal((Kh.n_dof, r))
        Xd = core.to_device(X, torch.float64)
        Y = torch.empty_like(Xd)
        rc = _lib.load().fea_spmm(Kh.n_dof, 1, rowptr.data_ptr(), colidx.data_ptr(), Kh.values.data_ptr(),
                                  Xd.data_ptr(), Y.data_ptr(), r, None)
        assert rc == 0 and rel(Y.cpu().numpy(), Khref @ X) < 1e-13
    # D = 1 with short rows (<= 6 entries): the union-gather path instead of the wide-row path
    rowptr, colidx = K.pattern.csr(2)
    for r in (4, 7):
        X = rng.standard_normal((K.n_dof, r))
        Xd = core.to_device(X, torch.float64)
        Y = torch.empty_like(Xd)
        rc = _lib.load().fea_spmm(K.n_dof, 1, rowptr.data_ptr(), colidx.data_ptr(), K.values.data_ptr(),
                                  Xd.data_ptr(), Y.data_ptr(), r, None)
        assert rc == 0 and rel(Y.cpu().numpy(), Kref @ X) < 1e-13


def test_mesh_extrude_device(mods):
    U, C = mods["utils"], mods["cubebeam"]
    n2, q2 = C.generate_quad_grid(5, 3, 0.3, 0.2)
    z = np.linspace(0, 2, 9)
    nh, eh = U.stack_faces_2d(n2, q2, z)
    nd, ed = U.stack_faces_2d_device(n2, q2, z)
    assert np.array_equal(nd.cpu().numpy(), nh) and np.array_equal(ed.cpu().numpy(), eh)


# ------------------------------------------------------- full-size, size-independent properties
def test_config3_full_size_properties(mods):
    """BASELINE config 3 (100x20x20, 133,623 DOF): pattern counts, equilibrium, symmetry of the
    response, true residual -- no oracle run needed at this size."""
    core, C = mods["core"], mods["cubebeam"]
    from fea_b200 import model

    nodes, elements, cons, forces = C.cantilever_case(100, 20)
    u, f, info, K = model.solve_hex8(nodes, elements, cons, forces, return_info=True)
    assert K.nnz == 10_080_189 and K.n_dof == 133_623
    assert info.status == 0 and info.rel_residual <= 1e-12 and 1500 < info.iterations < 3000
    free = cons.ravel() == 0
    # nodal forces on free DOF reproduce the applied loads; reactions balance them (quirk Q2)
    assert np.abs(f.ravel()[free] - forces.ravel()[free]).max() < 1e-8 * np.abs(forces).max()
    applied = forces.ravel()[free].reshape(-1)
    assert abs(f[nodes[:, 2] == 0][:, 1].sum() + forces[nodes[:, 2] != 0][:, 1].sum()) < 1e-6 * abs(applied.sum())
    # mirror symmetry about the plane x = width/2 of a y-load: u_y(x) = u_y(w - x), u_x antisymmetric
    grid = u.reshape(101, 21, 21, 3)
    assert np.abs(grid[..., 1] - grid[:, :, ::-1, 1]).max() < 1e-9 * np.abs(u).max()
    assert np.abs(grid[..., 0] + grid[:, :, ::-1, 0]).max() < 1e-9 * np.abs(u).max()
    # Euler-Bernoulli tip deflection of the equivalent beam within a few percent
    total = forces[:, 1].sum() - forces[nodes[:, 2] == 0][:, 1].sum()
    w_tip = grid[-1, :, :, 1].mean()
    EI = fo.E_HEX * 0.1**4 / 12
    assert 0.9 < w_tip / (total * 1.0**3 / (8 * EI)) < 1.15  # uniformly distributed load along z


def test_config3_full_vs_oracle(mods, golden):
    """BASELINE config 3 IN FULL against the oracle (tests/golden/oracle_config3.npz, written by
    oracle/make_golden_large.py: the C twin's Jacobi-PCG to 1e-12 AND scipy's sparse LU of the same
    reduced system): u within 1e-8 of both, same iteration count to a few, same reaction sum."""
    from fea_b200 import model

    g = golden("oracle_config3.npz")
    C = mods["cubebeam"]
    nodes, elements, cons, forces = C.cantilever_case(int(g["A"]), int(g["b"]))
    u, f, info, K = model.solve_hex8(nodes, elements, cons, forces, return_info=True)
    assert info.status == 0 and info.rel_residual <= 1e-12
    assert rel(u.ravel(), g["u_pcg"]) < U_RTOL
    assert rel(u.ravel(), g["u_direct"]) < U_RTOL
    assert abs(info.iterations - int(g["iterations"])) <= 25  # rounding noise in K moves the count by a few
    ry = f[nodes[:, 2] == 0][:, 1].sum()
    assert abs(ry - float(g["sum_reactions_y"])) < 1e-8 * abs(ry)


def test_config4_slab_vs_oracle_pcg(mods):
    """A config-4-shaped slab (50x80x80: 1,003,833 DOF, nnz 79.3 M) against the oracle's own Jacobi-PCG
    (oracle/fea_oracle_c.c on the host cores, run here): K values 1e-10, u 1e-8, iteration count."""
    from fea_b200 import model
    from oracle import c_oracle as co

    C = mods["cubebeam"]
    nodes, elements, cons, forces = C.cantilever_case(50, 80, length=0.125)
    assert nodes.size > 1_000_000
    u, f, info, K = model.solve_hex8(nodes, elements, cons, forces, return_info=True)
    co.set_threads(len(os.sched_getaffinity(0)))
    Ko = co.assemble_hex8(nodes, elements, fo.E_HEX, fo.NU_HEX)
    rowptr, colidx = (t.cpu().numpy() for t in K.pattern.csr(3))
    assert np.array_equal(rowptr, Ko.indptr) and np.array_equal(colidx, Ko.indices)
    assert rel(K.values.cpu().numpy(), Ko.data) < KE_RTOL
    free = fo.free_dofs(cons)
    uf, it, relres = co.jacobi_pcg(co.reduce_csr(Ko, free), forces.flatten()[free], tol=1e-12)
    uo = np.zeros(nodes.size)
    uo[free] = uf
    assert relres <= 1e-12 and info.rel_residual <= 1e-12
    assert rel(u.ravel(), uo) < U_RTOL
    assert abs(info.iterations - it) <= max(10, it // 50)
    assert rel(f.ravel(), co.spmv(Ko, uo)) < 1e-7


def test_config4_full_vs_oracle(mods, golden):
    """BASELINE config 4 IN FULL (400x80x80, 7.9 M DOF, ~8.5 k iterations) against the oracle's full
    Jacobi-PCG solve, recorded offline (one hour of CPU; tests/golden/oracle_config4.npz keeps a seeded
    20,000-entry sample of u and K_full u, the norms and the iteration count)."""
    from fea_b200 import model

    path = os.path.join(os.path.dirname(__file__), "golden", "oracle_config4.npz")
    if not os.path.isfile(path):
        pytest.skip("oracle_config4.npz not generated (python oracle/make_golden_large.py config4)")
    g = golden("oracle_config4.npz")
    C = mods["cubebeam"]
    nodes, elements, cons, forces = C.cantilever_case(int(g["A"]), int(g["b"]))
    u, f, info, K = model.solve_hex8(nodes, elements, cons, forces, return_info=True)
    assert info.status == 0 and info.rel_residual <= 1e-12
    idx = g["sample_index"]
    umax = float(g["u_max_abs"])
    assert np.abs(u.ravel()[idx] - g["u_sample"]).max() / umax < U_RTOL
    assert abs(np.abs(u).max() / umax - 1.0) < U_RTOL and int(np.abs(u).argmax()) == int(g["u_argmax"])
    assert abs(np.linalg.norm(u) / float(g["u_norm2"]) - 1.0) < U_RTOL
    assert np.abs(f.ravel()[idx] - g["f_sample"]).max() / np.abs(g["f_sample"]).max() < 1e-6
    assert abs(info.iterations - int(g["iterations"])) <= 100
    ry = f[nodes[:, 2] == 0][:, 1].sum()
    assert abs(ry - float(g["sum_reactions_y"])) < 1e-7 * abs(ry)


def test_spmv_beyond_int32_value_offsets(mods):
    """d*d*nnz_blocks > 2^31 (9.7 M nodes, 2.3 G matrix values, 18.6 GB): the bulk-copy SpMV computes its
    value offsets in 64 bits.  Rigid-body translations are in the null space of the unconstrained K,
    the product is symmetric, and the LAST rows (largest offsets) agree with an independent evaluation."""
    core, C, U = mods["core"], mods["cubebeam"], mods["utils"]
    if torch.cuda.mem_get_info()[0] < 60e9:
        pytest.skip("needs ~40 GB of free device memory")
    n2, q2 = C.generate_quad_grid(170, 170, 0.1, 0.1)
    nd, el = U.stack_faces_2d_device(n2, q2, np.linspace(0, 1.0, 331))
    K = core.assemble_hex8(nd, el, fo.E_HEX, fo.NU_HEX)
    assert K.values.numel() > 2**31
    scale = float(K.values.abs().max())
    t = torch.zeros(K.n_dof, dtype=torch.float64, device="cuda")
    t[1::3] = 1.0
    assert float(K.matvec(t).abs().max()) < 1e-10 * scale
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(K.n_dof, dtype=torch.float64, device="cuda", generator=g)
    y = torch.randn(K.n_dof, dtype=torch.float64, device="cuda", generator=g)
    Kx = K.matvec(x)
    assert abs(float(Kx @ y - x @ K.matvec(y))) < 1e-9 * abs(float(Kx @ y))
    # last 3 node rows straight from the stored blocks
    rp = K.pattern.node_rowptr[-4:].cpu().numpy().astype(np.int64)
    ci = K.pattern.node_colidx.cpu().numpy()
    for i in range(3):
        lo, hi = rp[i], rp[i + 1]
        cnt = int(hi - lo)
        blk = K.values[9 * lo:9 * hi].cpu().numpy().reshape(3, 3 * cnt)
        cols = (3 * ci[lo:hi][:, None] + np.arange(3)).ravel()
        want = blk @ x.cpu().numpy()[cols]
        got = Kx[3 * (K.pattern.n_nodes - 3 + i):3 * (K.pattern.n_nodes - 3 + i) + 3].cpu().numpy()
        assert np.abs(got - want).max() < 1e-9 * np.abs(want).max()
    del K, x, y, Kx
    torch.cuda.empty_cache()


def test_config4_assembly_properties(mods):
    """BASELINE config 4 (400x80x80, 7.9 M DOF) assembled on one GPU: structural nnz, null space,
    SpMV linearity and symmetry <Kx, y> = <x, Ky>."""
    core, C, U = mods["core"], mods["cubebeam"], mods["utils"]
    n2, q2 = C.generate_quad_grid(80, 80, 0.1, 0.1)
    nd, el = U.stack_faces_2d_device(n2, q2, np.linspace(0, 1.0, 401))
    K = core.assemble_hex8(nd, el, fo.E_HEX, fo.NU_HEX)
    assert K.n_dof == 7_892_883 and K.nnz == 627_797_529
    scale = float(K.values.abs().max())
    for c in range(3):
        t = torch.zeros(K.n_dof, dtype=torch.float64, device="cuda")
        t[c::3] = 1.0
        assert float(K.matvec(t).abs().max()) < 1e-11 * scale
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(K.n_dof, dtype=torch.float64, device="cuda", generator=g)
    y = torch.randn(K.n_dof, dtype=torch.float64, device="cuda", generator=g)
    Kx, Ky = K.matvec(x), K.matvec(y)
    assert abs(float(Kx @ y - x @ Ky)) < 1e-10 * abs(float(Kx @ y))
    z = K.matvec(2.0 * x - 3.0 * y)
    assert float((z - (2.0 * Kx - 3.0 * Ky)).abs().max()) < 1e-12 * float(Kx.abs().max())
    assert float(x @ Kx) > 0


def test_config2_full_size_properties(mods):
    """BASELINE config 2 (Euler-Bernoulli, 100 k elements) assembled at full size: structural nnz of a
    chain, rigid-body null space (uniform deflection; uniform rotation with the matching linear
    deflection), symmetry, and the analytic tip deflection P L^3 / (3 EI) recovered from a moderately
    sized cantilever (the 100 k solve itself is beyond FP64 conditioning)."""
    core, EB = mods["core"], mods["eb"]
    n = 100_000
    elements, EI, Ls, cons, loads = EB.cantilever_case(n)
    el = core.to_device(elements, torch.int32)
    K = core.assemble_beam(core.to_device(EI, torch.float64), core.to_device(Ls, torch.float64), el, n + 1)
    assert K.n_dof == 200_002 and K.nnz == 4 * (3 * (n + 1) - 2) == 1_200_004
    scale = float(K.values.abs().max())
    w = torch.zeros(K.n_dof, dtype=torch.float64, device="cuda")
    w[0::2] = 1.0  # rigid translation
    assert float(K.matvec(w).abs().max()) < 1e-9 * scale
    xs = torch.arange(n + 1, dtype=torch.float64, device="cuda") * float(Ls[0])
    rot = torch.zeros_like(w)
    rot[0::2], rot[1::2] = xs, 1.0  # rigid rotation: w = x, theta = 1
    assert float(K.matvec(rot).abs().max()) < 1e-9 * scale * float(rot.abs().max())
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(K.n_dof, dtype=torch.float64, device="cuda", generator=g)
    y = torch.randn(K.n_dof, dtype=torch.float64, device="cuda", generator=g)
    assert abs(float(K.matvec(x) @ y - x @ K.matvec(y))) < 1e-10 * abs(float(K.matvec(x) @ y))
    # analytic check where FP64 can carry it (cond ~ n^4): 50 Hermite elements are nodally exact for a tip load
    e2, EI2, L2, c2, l2 = EB.cantilever_case(50)
    u = EB.solve_beam(e2, EI2, L2, c2, l2)
    tip = -1000.0 * 1.0**3 / (3 * 210e9 * 1e-6)
    assert abs(np.asarray(u).reshape(-1, 2)[-1, 0] - tip) < 1e-6 * abs(tip)


def test_config5_full_size_properties(mods):
    """BASELINE config 5 (jittered lattice truss n = 93: 10,224,788 members, 2.4 M DOF) assembled at
    full size: member and structural non-zero counts frozen in SURVEY.md §8(d), rigid translations in
    the null space of the unconstrained K, symmetry and linearity of the 64-column SpMM, one batched
    PCG chunk reducing every column's residual."""
    core, T = mods["core"], mods["truss"]
    nodes, members, k, cons, loads = T.lattice_truss(93, 64)
    assert members.shape[0] == 10_224_788 and nodes.shape[0] == 804_357
    nd, mem = core.to_device(nodes, torch.float64), core.to_device(members, torch.int32)
    kd = core.to_device(k, torch.float64)
    K = core.assemble_truss(nd, mem, kd)
    assert K.n_dof == 2_413_071 and K.nnz == 191_285_397
    scale = float(K.values.abs().max())
    Tm = torch.zeros((K.n_dof, 4), dtype=torch.float64, device="cuda")
    for c in range(3):
        Tm[c::3, c] = 1.0
    assert float(K.matmat(Tm).abs().max()) < 1e-11 * scale  # 4 columns: scalar path of the SpMM
    B = core.to_device(loads, torch.float64)
    g = torch.Generator(device="cuda").manual_seed(2)
    Y = torch.randn(B.shape, dtype=torch.float64, device="cuda", generator=g)
    KB, KY = K.matmat(B), K.matmat(Y)
    lhs, rhs = (KB * Y).sum(dim=0), (B * KY).sum(dim=0)
    assert float(((lhs - rhs).abs() / lhs.abs()).max()) < 1e-9
    Z = K.matmat(2.0 * B - 3.0 * Y)
    assert float((Z - (2.0 * KB - 3.0 * KY)).abs().max()) < 1e-12 * float(KB.abs().max())
    del KB, KY, Z, Y
    Kc = core.assemble_truss(nd, mem, kd, fixed=core._fixed_mask(cons, nodes.size))
    X, info = core.pcg_multi(Kc, B, tol=1e-12, max_iter=48, raise_on_failure=False)
    R = B - Kc.matmat(X)
    free = core._fixed_mask(cons, nodes.size) == 0
    rel_res = R[free].norm(dim=0) / B[free].norm(dim=0)
    assert info.iterations == 48 and float(rel_res.max()) < 0.5  # every column is converging


def test_p2p_solver_single_rank(mods, monkeypatch):
    """fea_pcg_solve_p2p with world = 1 (no peers): exercises the comm block API, the private
    stream / CUDA-graph driver and the p-in-comm-block layout on one GPU; must reproduce
    fea_pcg_solve bit for bit (same kernels, same order)."""
    import ctypes

    from fea_b200 import _lib

    core = mods["core"]
    lib = _lib.load()
    # "same kernels": the two-kernel path with the pipeline shape the peer-memory solver uses (below 3 M DOF
    # fea_pcg_solve would otherwise take the persistent kernel on a 3-group pipeline)
    monkeypatch.setenv("FEA_PCG_FUSED", "0")
    monkeypatch.setenv("FEA_TMA_CFG", "2,2,3")
    nodes, elements, cons, forces = fo.cantilever_case(24, 5)
    nd, el = core.to_device(nodes, torch.float64), core.to_device(elements, torch.int32)
    K = core.assemble_hex8(nd, el, fo.E_HEX, fo.NU_HEX, fixed=core._fixed_mask(cons, nodes.size))
    b = core.to_device(forces, torch.float64).reshape(-1)
    u_ref, info_ref = core.pcg(K, b)
    n = K.n_dof
    own = ctypes.c_void_p()
    assert lib.fea_comm_alloc(lib.fea_comm_bytes(n), ctypes.byref(own)) == 0
    handle = (ctypes.c_ubyte * 64)()
    assert lib.fea_comm_ipc_export(own.value, ctypes.addressof(handle)) == 0 and any(handle)
    desc = _lib.PeerComm()
    desc.world, desc.rank, desc.lower_peer, desc.upper_peer, desc.epoch = 1, 0, -1, -1, 1
    desc.comm[0] = own.value
    x = torch.empty(n, dtype=torch.float64, device="cuda")
    ws = lib.fea_pcg_workspace(n)
    work = torch.empty(ws, dtype=torch.uint8, device="cuda")
    res = _lib.PcgResult()
    pt = K.pattern
    desc.algo = -1  # recurrence chosen from the size, as fea_pcg_solve does
    _, info_h = core.pcg(K, b, history=True)
    hist = torch.zeros(10 * n, dtype=torch.float64, device="cuda")
    for epoch in (1, 2):  # the block is reusable across solves
        desc.epoch = epoch
        rc = lib.fea_pcg_solve_p2p(pt.n_nodes, 3, pt.node_rowptr.data_ptr(), pt.node_colidx.data_ptr(),
                                   K.values.data_ptr(), pt.max_coupled, K.dinv.data_ptr(), b.data_ptr(), x.data_ptr(),
                                   1e-12, 10 * n, work.data_ptr(), ws, hist.data_ptr(), ctypes.byref(desc),
                                   ctypes.byref(res), None)
        assert rc == 0 and res.status == 0
        assert res.iterations == info_ref.iterations
        assert torch.equal(x, u_ref)
        # the residual history argument records what fea_pcg_solve records
        assert np.array_equal(hist[:res.iterations].cpu().numpy(), info_h.history)
    assert lib.fea_comm_free(own.value) == 0


def test_checked_build():
    """The library built with -DFEA_CHECKED (device-side assertions on ring offsets and capacities, slot
    searches, partial-sum and peer-slot indices: csrc/common.cuh) runs tools/sanitize.py -- every kernel
    family on small cases, compared with the oracle -- without tripping an assertion.  This stands in for
    compute-sanitizer memcheck, which is closed on the GPU pool."""
    import subprocess
    import sys

    from fea_b200 import build

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = build.build_library(checked=True)
    env = dict(os.environ, FEA_LIB_PATH=lib)
    res = subprocess.run([sys.executable, os.path.join(root, "tools", "sanitize.py")], capture_output=True, text=True,
                         timeout=900, cwd=root, env=env)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "SANITIZE single-GPU pass ok" in res.stdout and "Assertion" not in res.stderr


def test_multi_rank_parity(tmp_path):
    """N-rank slab solve == 1-rank solve: owned K rows bit-identical, u within 1e-10,
    same history; peer-memory solver with both recurrences and both cut styles, the NCCL driver, and the
    collective public cubebeam.solve.  Spawns tools/check_dist.py on 2 ranks; needs >= 2 visible GPUs
    (`gpurun --gpus 2 -- python -m pytest tests -m gpu`)."""
    import socket
    import subprocess
    import sys

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    out_json = tmp_path / "dist.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(root, "tools", "check_dist.py"), "48", "12", "--out",
           str(out_json)]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=root)
    assert res.returncode == 0, res.stdout[-4000:] + res.stderr[-4000:]
    assert "DIST CHECK PASS" in res.stdout, res.stdout[-6000:]
