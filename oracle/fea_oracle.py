"""CPU oracle: a numpy/scipy restatement of jjrreett/fea's hot path (Ke -> assembly -> solve).

TEST INFRASTRUCTURE ONLY -- this is the checker, never the product.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import it.  `fea_b200/` must never import anything under `oracle/`.

Parity status: PINNED.  The reference ships no tests, so the pins are outputs of
the reference itself, run in the build container by `oracle/make_golden.py` through
`oracle/ref_loader.py` and committed under `tests/golden/` (K1..K8 of SURVEY.md §4).
`tests/test_oracle_golden.py` checks every function here against those fixtures on any box;
`tests/test_oracle_vs_reference.py` re-checks against the live reference when /root/reference
exists.  Third-party arithmetic on the reference path: numpy (`det`, `inv`, `solve`, `@`),
pinned 2.2.0 in the reference's uv.lock:363, 2.3.5 here -- LAPACK results agree to ~1e-11 rel,
not bit for bit.

Every function cites the reference file:line it follows.  Where the reference cannot scale
(dense K, dense LU: cubebeam.py:80,98) the restatement is sparse (`coo -> csr`) and the solver
is the explicit Jacobi-PCG that `BASELINE.json:north_star` names; both are proved equal to the
dense reference path on the shipped meshes before being trusted at size.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.sparse as sp

# ----------------------------------------------------------------------------------------------
# constants of the shipped scripts (cubebeam.py:9-25, fea.py:10-35, euler_bernoulli.py:5-15)
# ----------------------------------------------------------------------------------------------
PSI = 6894.76
LBF = 4.44822
FT = 0.3048
INCH = 0.0254
E_HEX = 10_000_000 * PSI  # cubebeam.py:84
NU_HEX = 0.3  # cubebeam.py:84

# Local node signs (xi, eta, zeta): bottom face CCW then top face CCW (utils.py:159-197, 351-353)
HEX8_SIGNS = np.array(
    [
        [-1, -1, -1],
        [+1, -1, -1],
        [+1, +1, -1],
        [-1, +1, -1],
        [-1, -1, +1],
        [+1, -1, +1],
        [+1, +1, +1],
        [-1, +1, +1],
    ],
    dtype=np.float64,
)

JACOBIAN_MESSAGE = "Jacobian determinant is non-positive. Check the element shape."


# ----------------------------------------------------------------------------------------------
# H1: hex8 element stiffness (utils.py:127-239)
# ----------------------------------------------------------------------------------------------
def elasticity_matrix(E: float, nu: float) -> np.ndarray:
    """6x6 isotropic C, engineering shear strains (utils.py:144-153)."""
    C = np.zeros((6, 6))
    C[:3, :3] = nu
    C[0, 0] = C[1, 1] = C[2, 2] = 1 - nu
    C[3, 3] = C[4, 4] = C[5, 5] = (1 - 2 * nu) / 2
    return (E / ((1 + nu) * (1 - 2 * nu))) * C


def hex8_shape_derivatives(xi: float, eta: float, zeta: float) -> np.ndarray:
    """dN/d(xi,eta,zeta), shape (3, 8), trilinear, divided by 8 (utils.py:159-197)."""
    sx, sy, sz = HEX8_SIGNS[:, 0], HEX8_SIGNS[:, 1], HEX8_SIGNS[:, 2]
    fx, fy, fz = 1 + sx * xi, 1 + sy * eta, 1 + sz * zeta
    return np.stack([sx * fy * fz, sy * fx * fz, sz * fx * fy]) / 8.0


def gauss_points_2x2x2() -> np.ndarray:
    """The 8 points in the reference's loop order: xi outer, eta, zeta inner; weights are all 1
    (utils.py:140-141, 200-204)."""
    g = np.array([-1 / np.sqrt(3), 1 / np.sqrt(3)])
    return np.array([[a, b, c] for a in g for b in g for c in g])


def strain_displacement(dN_dx: np.ndarray) -> np.ndarray:
    """B (6,24) from dN/dx (3,8): rows exx, eyy, ezz, gxy, gyz, gzx (utils.py:224-234)."""
    B = np.zeros((6, 24))
    B[0, 0::3] = dN_dx[0]
    B[1, 1::3] = dN_dx[1]
    B[2, 2::3] = dN_dx[2]
    B[3, 0::3] = dN_dx[1]
    B[3, 1::3] = dN_dx[0]
    B[4, 1::3] = dN_dx[2]
    B[4, 2::3] = dN_dx[1]
    B[5, 0::3] = dN_dx[2]
    B[5, 2::3] = dN_dx[0]
    return B


def hex8_ke(nodes8: np.ndarray, E: float, nu: float) -> np.ndarray:
    """One element, the reference's own sequence of operations (utils.py:200-237):
    J = dN @ X, det, inv, dN_dx = J^-1 dN, B, Ke += w * (B^T C B) * detJ."""
    nodes8 = np.asarray(nodes8, dtype=np.float64)
    C = elasticity_matrix(E, nu)
    Ke = np.zeros((24, 24))
    for xi, eta, zeta in gauss_points_2x2x2():
        dN = hex8_shape_derivatives(xi, eta, zeta)
        J = dN @ nodes8
        detJ = np.linalg.det(J)
        if detJ <= 0:  # utils.py:212-215
            raise ValueError(JACOBIAN_MESSAGE)
        B = strain_displacement(np.linalg.inv(J) @ dN)
        Ke += 1.0 * (B.T @ C @ B) * detJ
    return Ke


def hex8_trilinear_coefficients(nodes8: np.ndarray) -> np.ndarray:
    """8 x (coefficient vectors of 1, xi, eta, xi.eta, zeta, xi.zeta, eta.zeta, xi.eta.zeta) of the trilinear map
    x(xi, eta, zeta) = sum_a N_a X_a (utils.py:159-197), i.e. the Walsh-Hadamard transform of the corners over
    their three sign bits.  Row m (bits: 1 = xi, 2 = eta, 4 = zeta).  Rows 1, 2, 4 are the rows of 8 J at the centre."""
    X = np.asarray(nodes8, dtype=np.float64)
    # corner whose (x, y, z) sign bits are the bits of m, then the butterfly bit by bit: differences of corners that
    # share coordinate values cancel EXACTLY in this order (a plain signed sum over the 8 corners does not), which
    # is what makes "== 0" a usable test on grid elements
    c = np.array([X[[a for a in range(8) if tuple(HEX8_SIGNS[a] > 0) == (bool(m & 1), bool(m & 2), bool(m & 4))][0]]
                  for m in range(8)])
    for bit in (1, 2, 4):
        for m in range(8):
            if not m & bit:
                lo, hi = c[m].copy(), c[m | bit].copy()
                c[m], c[m | bit] = lo + hi, hi - lo
    return c


def hex8_ke_affine(nodes8: np.ndarray, E: float, nu: float) -> np.ndarray:
    """Closed form of `hex8_ke` for an AFFINE element (a parallelepiped: the xi.eta, xi.zeta, eta.zeta and
    xi.eta.zeta coefficients of the trilinear map vanish, so J is constant).  The 2x2x2 quadrature of
    utils.py:200-237 then reduces to
        S_ab = sum_gp detJ grad N_a grad N_b^T = adj(Jc) M_ab adj(Jc)^T / (8 det Jc),   Jc = 8 J,
        M_ab = sum_gp dN_a dN_b^T  (a constant table),
    and K_ab[r][r] = C11 S_rr + C44 (S_ss + S_tt), K_ab[r][c] = C12 S_rc + C44 S_cr with the entries of the C of
    utils.py:144-153.  This is the checker of the CUDA assembly's fast path (fea_b200/csrc/assemble.cu); it must
    agree with `hex8_ke` to rounding on every affine element."""
    c = hex8_trilinear_coefficients(nodes8)
    Jc = np.array([c[1], c[2], c[4]])
    det = np.linalg.det(Jc)
    if det <= 0:
        raise ValueError(JACOBIAN_MESSAGE)
    adj = np.linalg.inv(Jc) * det
    dN = [hex8_shape_derivatives(*gp) for gp in gauss_points_2x2x2()]
    C = elasticity_matrix(E, nu)
    c11, c12, c44 = C[0, 0], C[0, 1], C[3, 3]
    Ke = np.zeros((24, 24))
    for a in range(8):
        for b in range(8):
            M = sum(np.outer(d[:, a], d[:, b]) for d in dN)
            S = adj @ M @ adj.T / (8.0 * det)
            blk = c12 * S + c44 * S.T
            tr = np.trace(S)
            for r in range(3):
                blk[r, r] = c11 * S[r, r] + c44 * (tr - S[r, r])
            Ke[3 * a:3 * a + 3, 3 * b:3 * b + 3] = blk
    return Ke


def _det_inv_3x3(J: np.ndarray):
    """Closed-form det / inverse of a stack of 3x3 matrices (what utils.py:211,218 get from
    LAPACK; differs from it by rounding only)."""
    a, b, c = J[..., 0, 0], J[..., 0, 1], J[..., 0, 2]
    d, e, f = J[..., 1, 0], J[..., 1, 1], J[..., 1, 2]
    g, h, i = J[..., 2, 0], J[..., 2, 1], J[..., 2, 2]
    c00, c01, c02 = e * i - f * h, f * g - d * i, d * h - e * g
    det = a * c00 + b * c01 + c * c02
    inv = np.empty_like(J)
    inv[..., 0, 0], inv[..., 0, 1], inv[..., 0, 2] = c00, c * h - b * i, b * f - c * e
    inv[..., 1, 0], inv[..., 1, 1], inv[..., 1, 2] = c01, a * i - c * g, c * d - a * f
    inv[..., 2, 0], inv[..., 2, 1], inv[..., 2, 2] = c02, b * g - a * h, a * e - b * d
    inv /= det[..., None, None]
    return det, inv


def hex8_ke_batched(nodes: np.ndarray, elements: np.ndarray, E: float, nu: float,
                    chunk: int = 20000) -> np.ndarray:
    """All elements at once, (M,24,24).  Same formula as `hex8_ke` (utils.py:200-237), einsum
    instead of the Python loops.  Raises the reference's ValueError if any detJ <= 0."""
    nodes = np.asarray(nodes, dtype=np.float64)
    elements = np.asarray(elements)
    M = elements.shape[0]
    C = elasticity_matrix(E, nu)
    gps = gauss_points_2x2x2()
    dNs = np.stack([hex8_shape_derivatives(*gp) for gp in gps])  # (8 gp, 3, 8)
    out = np.empty((M, 24, 24))
    for lo in range(0, M, chunk):
        X = nodes[elements[lo:lo + chunk]]  # (m, 8, 3)
        m = X.shape[0]
        Ke = np.zeros((m, 24, 24))
        for dN in dNs:
            J = np.einsum("an,mnc->mac", dN, X)
            detJ, Jinv = _det_inv_3x3(J)
            if np.any(detJ <= 0):
                raise ValueError(JACOBIAN_MESSAGE)
            dNdx = np.einsum("mab,bn->man", Jinv, dN)  # (m, 3, 8)
            B = np.zeros((m, 6, 24))
            B[:, 0, 0::3] = dNdx[:, 0]
            B[:, 1, 1::3] = dNdx[:, 1]
            B[:, 2, 2::3] = dNdx[:, 2]
            B[:, 3, 0::3] = dNdx[:, 1]
            B[:, 3, 1::3] = dNdx[:, 0]
            B[:, 4, 1::3] = dNdx[:, 2]
            B[:, 4, 2::3] = dNdx[:, 1]
            B[:, 5, 0::3] = dNdx[:, 2]
            B[:, 5, 2::3] = dNdx[:, 0]
            CB = np.einsum("ij,mjk->mik", C, B)
            Ke += np.einsum("mji,mjk->mik", B, CB) * detJ[:, None, None]
        out[lo:lo + chunk] = Ke
    return out


# ----------------------------------------------------------------------------------------------
# H2/H3: DOF map and assembly (cubebeam.py:80-90 == fea.py:87-97), sparse instead of dense
# ----------------------------------------------------------------------------------------------
def element_dofs(elements: np.ndarray, dof_per_node: int) -> np.ndarray:
    """(M, npe*d) global DOF = d*node + c, node-major (cubebeam.py:86; euler_bernoulli.py:44)."""
    elements = np.asarray(elements, dtype=np.int64)
    d = dof_per_node
    return (elements[:, :, None] * d + np.arange(d)[None, None, :]).reshape(elements.shape[0], -1)


def assemble_csr(elements: np.ndarray, Ke: np.ndarray, n_nodes: int, dof_per_node: int) -> sp.csr_matrix:
    """K = sum_e scatter(Ke_e).  The STRUCTURAL pattern: every (row, col) pair of every element
    is kept even when values cancel to 0.0; duplicates summed; indices sorted.
    `indptr` / `indices` are the bit-exact target for the CUDA symbolic pass."""
    dofs = element_dofs(elements, dof_per_node)
    nd = dofs.shape[1]
    rows = np.repeat(dofs, nd, axis=1).ravel()
    cols = np.tile(dofs, (1, nd)).ravel()
    n = n_nodes * dof_per_node
    K = sp.coo_matrix((np.asarray(Ke, dtype=np.float64).ravel(), (rows, cols)), shape=(n, n)).tocsr()
    K.sum_duplicates()
    K.sort_indices()
    return K


def structural_pattern(elements: np.ndarray, n_nodes: int, dof_per_node: int):
    """(indptr, indices) only, via the same scipy path as `assemble_csr` with all-ones values
    (ones never cancel, so scipy cannot drop anything)."""
    M, npe = np.asarray(elements).shape
    nd = npe * dof_per_node
    K = assemble_csr(elements, np.ones((M, nd, nd)), n_nodes, dof_per_node)
    return K.indptr.astype(np.int64), K.indices.astype(np.int64)


def assemble_dense(elements: np.ndarray, Ke: np.ndarray, n_nodes: int, dof_per_node: int) -> np.ndarray:
    """The reference's own dense scatter in element order (cubebeam.py:80-90); small meshes only."""
    n = n_nodes * dof_per_node
    K = np.zeros((n, n))
    for dofs, ke in zip(element_dofs(elements, dof_per_node), Ke):
        K[np.ix_(dofs, dofs)] += ke
    return K


# ----------------------------------------------------------------------------------------------
# H4-H7: constraint reduction, solve, expansion, reactions (cubebeam.py:92-108)
# ----------------------------------------------------------------------------------------------
def free_dofs(constraints: np.ndarray) -> np.ndarray:
    """Ascending indices of unconstrained DOF (cubebeam.py:92)."""
    return np.where(np.asarray(constraints).flatten() == 0)[0]


def jacobi_pcg(A: sp.csr_matrix, b: np.ndarray, tol: float = 1e-12, maxiter: int | None = None):
    """Explicit Jacobi-preconditioned CG (the solver `north_star` names; the reference only has
    `np.linalg.solve` and a `# TODO iterative solver`, cubebeam.py:98-99).

    x0 = 0; stop when the RECURRENCE residual satisfies ||r||_2 <= tol * ||b||_2.
    Returns (x, iterations, residual history ||r||/||b|| per iteration)."""
    n = A.shape[0]
    b = np.asarray(b, dtype=np.float64)
    if maxiter is None:
        maxiter = 10 * n
    dinv = 1.0 / A.diagonal()
    x = np.zeros(n)
    bnorm = np.linalg.norm(b)
    hist = []
    if bnorm == 0.0:
        return x, 0, hist
    r = b.copy()
    z = dinv * r
    p = z.copy()
    rz = r @ z
    it = 0
    while it < maxiter:
        Ap = A @ p
        pAp = p @ Ap
        if not pAp > 0.0:
            raise np.linalg.LinAlgError("PCG breakdown: p.Ap <= 0 (matrix not positive definite)")
        alpha = rz / pAp
        x += alpha * p
        r -= alpha * Ap
        it += 1
        rel = np.linalg.norm(r) / bnorm
        hist.append(rel)
        if rel <= tol:
            break
        z = dinv * r
        rz_new = r @ z
        p = z + (rz_new / rz) * p
        rz = rz_new
    return x, it, hist


def solve_system(K: sp.csr_matrix, constraints: np.ndarray, forces: np.ndarray, method: str = "direct",
                 tol: float = 1e-12, maxiter: int | None = None):
    """Reduce, solve, expand, reactions -- cubebeam.py:92-108 on a sparse K.

    Returns (u_flat, reactions_flat, info).  `reactions = K_full @ u` uses the UNREDUCED K
    (cubebeam.py:106).  Loads on constrained DOF are dropped by the reduction
    (quirk Q2)."""
    free = free_dofs(constraints)
    Kff = K[free][:, free].tocsr()
    ff = np.asarray(forces, dtype=np.float64).flatten()[free]
    info = {"free": free, "n_free": free.size}
    if method == "direct":
        import scipy.sparse.linalg as spla

        uf = spla.spsolve(Kff.tocsc(), ff)
    elif method == "dense":
        uf = np.linalg.solve(Kff.toarray(), ff)  # cubebeam.py:98 verbatim
    elif method == "pcg":
        uf, it, hist = jacobi_pcg(Kff, ff, tol=tol, maxiter=maxiter)
        info["iterations"] = it
        info["history"] = hist
    else:
        raise ValueError(method)
    u = np.zeros(K.shape[0])
    u[free] = uf
    return u, K @ u, info


def solve_hex8(nodes, elements, constraints, forces, E: float = E_HEX, nu: float = NU_HEX,
               method: str = "direct", tol: float = 1e-12, maxiter: int | None = None):
    """`solve(nodes, elements, constraints, forces) -> (displacements, forces)` of
    cubebeam.py:79-108 == fea.py:86-115, sparse."""
    nodes = np.asarray(nodes, dtype=np.float64)
    Ke = hex8_ke_batched(nodes, elements, E, nu)
    K = assemble_csr(elements, Ke, nodes.shape[0], 3)
    u, f, info = solve_system(K, constraints, forces, method=method, tol=tol, maxiter=maxiter)
    info["K"] = K
    return u.reshape(nodes.shape), f.reshape(nodes.shape), info


# ----------------------------------------------------------------------------------------------
# meshes (cubebeam.py:28-57, utils.py:356-376, fea.py:28-72)
# ----------------------------------------------------------------------------------------------
def generate_quad_grid(nx: int, ny: int, width: float, height: float):
    """(nx+1)(ny+1) nodes, x fastest; quads [n1, n2, n4, n3] CCW (cubebeam.py:28-57)."""
    x = np.linspace(0, width, nx + 1)
    y = np.linspace(0, height, ny + 1)
    nodes = np.empty(((nx + 1) * (ny + 1), 2))
    quads = np.empty((nx * ny, 4), dtype=np.int64)
    for j in range(ny + 1):
        for i in range(nx + 1):
            nodes[j * (nx + 1) + i] = (x[i], y[j])
    for j in range(ny):
        for i in range(nx):
            n1 = j * (nx + 1) + i
            quads[j * nx + i] = (n1, n1 + 1, n1 + nx + 2, n1 + nx + 1)
    return nodes, quads


def stack_faces_2d(nodes2d: np.ndarray, faces2d: np.ndarray, z_heights):
    """Extrude: node = layer*n2d + i; element = [face + lo, face + hi] (utils.py:356-376).
    `z_heights` is the NODE-layer list (quirk Q1)."""
    nodes2d = np.asarray(nodes2d, dtype=np.float64)
    faces2d = np.asarray(faces2d, dtype=np.int64)
    n2d, nl = nodes2d.shape[0], len(z_heights)
    nodes3d = np.zeros((n2d * nl, 3))
    for layer, z in enumerate(z_heights):
        nodes3d[layer * n2d:(layer + 1) * n2d, :2] = nodes2d
        nodes3d[layer * n2d:(layer + 1) * n2d, 2] = z
    elems = [np.concatenate([face + l * n2d, face + (l + 1) * n2d]) for l in range(nl - 1) for face in faces2d]
    return nodes3d, np.array(elems, dtype=np.int64).reshape(-1, 8)


def _beam_bar(n_width, n_node_layers, beam_width, beam_length, f_node):
    n2, q2 = generate_quad_grid(n_width, n_width, beam_width, beam_width)
    nodes, elements = stack_faces_2d(n2, q2, np.linspace(0, beam_length, n_node_layers))
    constraints = np.zeros(nodes.shape, dtype=np.int64)
    constraints[nodes[:, 2] == 0] = 1  # cubebeam.py:112-114 (exact float compare, quirk Q3)
    forces = np.zeros(nodes.shape)
    forces[nodes[:, 1] == 0] += np.array([0.0, f_node, 0.0])  # cubebeam.py:116-118
    return nodes, elements, constraints, forces


def cubebeam_case(n_width: int = 4, n_height: int = 50, beam_width: float = 0.1, beam_length: float = 1.0):
    """Mesh, constraints and loads of the shipped cubebeam.py run (cubebeam.py:9-25, 60-66,
    111-118).  `n_height` is passed to linspace as the NODE-layer count (quirk Q1) while the load
    per node divides by (n_width+1)*(n_height+1) (cubebeam.py:24-25)."""
    total_load = 100.0 * LBF / FT * beam_length
    f_node = total_load / ((n_width + 1) * (n_height + 1))
    return _beam_bar(n_width, n_height, beam_width, beam_length, f_node)


def cantilever_case(A: int, b: int, beam_width: float = 0.1, beam_length: float = 1.0):
    """BASELINE configs 3/4: A element layers along z (linspace(0, L, A+1)),
    b x b elements in the section, load (0, f, 0) on every node with y = 0,
    f = total_load / ((b+1)(A+1)) = total_load / (number of loaded nodes)."""
    total_load = 100.0 * LBF / FT * beam_length
    f_node = total_load / ((b + 1) * (A + 1))
    return _beam_bar(b, A + 1, beam_width, beam_length, f_node)


def tube_case(n_seg: int = 26, n_layers: int = 50):
    """Mesh, constraints and (scrambled, quirk Q5/Q6) loads of the shipped fea.py run
    (fea.py:28-72, 119-122)."""
    outer_radius, inner_radius = 4 * INCH, 3.9 * INCH
    thetas = np.linspace(0, np.pi * 2, n_seg, endpoint=False).reshape(-1, 1)
    unit = np.hstack([np.cos(thetas), np.sin(thetas)])
    nodes2d = np.vstack([unit * inner_radius, unit * outer_radius])
    faces = np.array([[i, i + n_seg, (i + 1) % n_seg + n_seg, (i + 1) % n_seg] for i in range(n_seg)], dtype=np.int64)
    forces2d = np.zeros_like(nodes2d)
    sel = slice(n_seg, (3 * n_seg) // 2)
    forces2d[sel, 1] = -np.cos(np.pi / 2 * nodes2d[sel, 0] / outer_radius) * np.pi / 4 / outer_radius
    nodes, elements = stack_faces_2d(nodes2d, faces, np.linspace(0, 1.0, n_layers))
    forces = np.zeros_like(nodes)
    forces[:, :2] = forces2d.repeat(n_layers, axis=0)  # fea.py:71, reproduced verbatim (Q5)
    constraints = np.zeros(nodes.shape, dtype=np.int64)
    constraints[nodes[:, 2] == 0] = 1
    return nodes, elements, constraints, forces


# ----------------------------------------------------------------------------------------------
# B1-B3: Euler-Bernoulli beam (euler_bernoulli.py:22-73) and its post-processing (:76-102)
# ----------------------------------------------------------------------------------------------
def beam_ke(EI: float, L: float) -> np.ndarray:
    """4x4 Hermite stiffness, DOF (w1, th1, w2, th2) (euler_bernoulli.py:22-39)."""
    return (EI / L**3) * np.array(
        [
            [12, 6 * L, -12, 6 * L],
            [6 * L, 4 * L**2, -6 * L, 2 * L**2],
            [-12, -6 * L, 12, -6 * L],
            [6 * L, 2 * L**2, -6 * L, 4 * L**2],
        ]
    )


def beam_ke_batched(EI: np.ndarray, L: np.ndarray) -> np.ndarray:
    EI = np.asarray(EI, dtype=np.float64)
    L = np.asarray(L, dtype=np.float64)
    return np.stack([beam_ke(a, b) for a, b in zip(EI, L)]) if EI.size < 4096 else _beam_ke_vec(EI, L)


def _beam_ke_vec(EI, L):
    c = EI / L**3
    z = np.zeros_like(L)
    six, four, two = 6 * L, 4 * L**2, 2 * L**2
    t = 12 + z
    rows = [
        [t, six, -t, six],
        [six, four, -six, two],
        [-t, -six, t, -six],
        [six, two, -six, four],
    ]
    return c[:, None, None] * np.stack([np.stack(r, axis=-1) for r in rows], axis=-2)


def beam_elements(n_elements: int) -> np.ndarray:
    """Connectivity [i, i+1]; DOF [2i, 2i+1, 2i+2, 2i+3] (euler_bernoulli.py:44)."""
    i = np.arange(n_elements, dtype=np.int64)
    return np.stack([i, i + 1], axis=1)


def beam_udl_load(q: float, n_elements: int, L_e: float) -> np.ndarray:
    """Consistent load vector for a uniform load, accumulated element by element exactly as
    euler_bernoulli.py:52-57 (note the reference's L/6 moment arm, kept verbatim)."""
    f = np.zeros(2 * (n_elements + 1))
    fe = q * L_e / 2 * np.array([1, L_e / 6, 1, -L_e / 6])
    for i in range(n_elements):
        f[2 * i:2 * i + 4] += fe
    return f


def beam_fixed_fixed_case(n_elements: int = 100, E: float = 210e9, I: float = 1e-6, L: float = 1.0, q: float = 1000):
    """The shipped euler_bernoulli.py problem (:5-19, 52-61): fixed at both ends, UDL."""
    L_e = L / n_elements
    elements = beam_elements(n_elements)
    EI = np.full(n_elements, E * I)
    Ls = np.full(n_elements, L_e)
    constraints = np.zeros((n_elements + 1, 2), dtype=np.int64)
    constraints[0] = 1
    constraints[-1] = 1
    loads = beam_udl_load(q, n_elements, L_e).reshape(-1, 2)
    return elements, EI, Ls, constraints, loads


def beam_cantilever_case(n_elements: int, E: float = 210e9, I: float = 1e-6, L: float = 1.0, P: float = -1000.0):
    """BASELINE config 2: fixed DOF {0,1}, tip load P on DOF 2n."""
    L_e = L / n_elements
    elements = beam_elements(n_elements)
    EI = np.full(n_elements, E * I)
    Ls = np.full(n_elements, L_e)
    constraints = np.zeros((n_elements + 1, 2), dtype=np.int64)
    constraints[0] = 1
    loads = np.zeros((n_elements + 1, 2))
    loads[-1, 0] = P
    return elements, EI, Ls, constraints, loads


def solve_beam(elements, EI, Ls, constraints, loads, method: str = "direct", tol: float = 1e-12,
               maxiter: int | None = None):
    """Assemble + reduce + solve (euler_bernoulli.py:42-73).  Returns (u (n_nodes,2), K, info)."""
    n_nodes = constraints.shape[0]
    K = assemble_csr(elements, beam_ke_batched(EI, Ls), n_nodes, 2)
    u, _, info = solve_system(K, constraints, loads, method=method, tol=tol, maxiter=maxiter)
    return u.reshape(n_nodes, 2), K, info


def chain_solve_exact(K: sp.csr_matrix, constraints: np.ndarray, loads: np.ndarray, d: int = 2, digits: int = 60):
    """The (numerically) EXACT solution of the reduced chain system K_ff u_f = f_f for the FP64 matrix K as
    given: block Thomas elimination in `digits`-digit decimal arithmetic (euler_bernoulli.py:61-73 with
    np.linalg.solve replaced by exact elimination).  Constrained DOF become identity rows/columns.  Used to
    separate a solver's own rounding error from the sensitivity of the solution to the FP64 rounding of
    the matrix entries (cond(K) ~ 5 n^4 for the Hermite beam).  O(n) sequential; n <= ~1e5."""
    from decimal import Decimal, getcontext

    getcontext().prec = digits
    n = K.shape[0] // d
    fixed = np.asarray(constraints).reshape(-1) != 0
    f = np.asarray(loads, dtype=np.float64).reshape(-1)
    Kc = K.tocsr()

    def block(i, j):
        m = [[Decimal(0)] * d for _ in range(d)]
        if not 0 <= j < n:
            return m
        sub = Kc[d * i:d * i + d, d * j:d * j + d].toarray()
        for r in range(d):
            for c in range(d):
                if fixed[d * i + r] or fixed[d * j + c]:
                    m[r][c] = Decimal(1) if (i == j and r == c) else Decimal(0)
                else:
                    m[r][c] = Decimal(float(sub[r, c]))
        return m

    def mul(X, Y):
        return [[sum(X[r][k] * Y[k][c] for k in range(d)) for c in range(d)] for r in range(d)]

    def sub_(X, Y):
        return [[X[r][c] - Y[r][c] for c in range(d)] for r in range(d)]

    def mv(X, v):
        return [sum(X[r][k] * v[k] for k in range(d)) for r in range(d)]

    def inv(X):
        if d == 1:
            return [[Decimal(1) / X[0][0]]]
        det = X[0][0] * X[1][1] - X[0][1] * X[1][0]
        return [[X[1][1] / det, -X[0][1] / det], [-X[1][0] / det, X[0][0] / det]]

    rhs = [[Decimal(0) if fixed[d * i + r] else Decimal(float(f[d * i + r])) for r in range(d)] for i in range(n)]
    Bp, dp, Cs = [], [], []
    for i in range(n):
        B, d_i = block(i, i), rhs[i]
        if i > 0:
            m = mul(block(i, i - 1), inv(Bp[-1]))
            B = sub_(B, mul(m, Cs[-1]))
            t = mv(m, dp[-1])
            d_i = [d_i[r] - t[r] for r in range(d)]
        Bp.append(B)
        dp.append(d_i)
        Cs.append(block(i, i + 1))
    u = [None] * n
    u[n - 1] = mv(inv(Bp[n - 1]), dp[n - 1])
    for i in range(n - 2, -1, -1):
        t = mv(Cs[i], u[i + 1])
        u[i] = mv(inv(Bp[i]), [dp[i][r] - t[r] for r in range(d)])
    return np.array([[float(v) for v in row] for row in u])


def beam_moment_shear(u_flat: np.ndarray, EI: np.ndarray, Ls: np.ndarray):
    """The reference's own per-element "moment"/"shear" formulas, verbatim in meaning
    (euler_bernoulli.py:76-102, quirk Q7): n_nodes entries, last one left 0."""
    u_flat = np.asarray(u_flat, dtype=np.float64).ravel()
    n = len(EI)
    M = np.zeros(n + 1)
    V = np.zeros(n + 1)
    for i in range(n):
        u = u_flat[2 * i:2 * i + 4]
        L = Ls[i]
        M[i] = EI[i] / L**2 * (12 * u[0] - 6 * L * u[1] - 12 * u[2] + 6 * L * u[3])
        V[i] = EI[i] / L**3 * (6 * L * u[0] + 2 * L**2 * u[1] - 6 * L * u[2] + 4 * L**2 * u[3])
    return M, V


# ----------------------------------------------------------------------------------------------
# T1, T2, T1': truss (truss.py:78-119) and its linearisation
# ----------------------------------------------------------------------------------------------
def truss_compute_forces(nodes, members, displaced_nodes, forces, stiffness=1000.0) -> None:
    """Member forces accumulated IN PLACE into `forces`, member by member in list order, in the
    dtype of the inputs (truss.py:78-92).  `stiffness` may be a scalar or per-member array."""
    k = np.broadcast_to(np.asarray(stiffness, dtype=forces.dtype), (len(members),))  # python-float k is a weak scalar
    for m, (a, b) in enumerate(members):
        d0 = nodes[b] - nodes[a]
        d1 = displaced_nodes[b] - displaced_nodes[a]
        l1 = np.linalg.norm(d1)
        dl = np.linalg.norm(d0) - l1
        fvec = (-k[m] * dl) * d1 / l1
        forces[a] += fvec
        forces[b] -= fvec


def truss_relax_step(nodes, members, displaced_nodes, loads, stiffness=1000.0):
    """One pass of the endless loop (truss.py:97-119): returns the residual norm at the first
    loaded node (what the script prints, truss.py:101-103) and updates `displaced_nodes` in place
    for loaded nodes only."""
    forces = np.zeros_like(nodes)
    truss_compute_forces(nodes, members, displaced_nodes, forces, stiffness)
    residual = float(np.linalg.norm(loads[0][1] + forces[loads[0][0]]))
    for i, load in loads:
        displaced_nodes[i] += (load + forces[i, :]) / displaced_nodes.dtype.type(stiffness)
    return residual


def truss_ke(xa: np.ndarray, xb: np.ndarray, k: float) -> np.ndarray:
    """T1' (SURVEY.md §8(a)): Jacobian of truss.py:78-92 at the undeformed state.
    Ke = k [[cc^T, -cc^T], [-cc^T, cc^T]], c = unit member axis; 3 DOF per node."""
    c = (np.asarray(xb, dtype=np.float64) - np.asarray(xa, dtype=np.float64))
    c = c / np.linalg.norm(c)
    cc = k * np.outer(c, c)
    return np.block([[cc, -cc], [-cc, cc]])


def truss_ke_batched(nodes: np.ndarray, members: np.ndarray, k: np.ndarray) -> np.ndarray:
    nodes = np.asarray(nodes, dtype=np.float64)
    members = np.asarray(members, dtype=np.int64)
    c = nodes[members[:, 1]] - nodes[members[:, 0]]
    c /= np.linalg.norm(c, axis=1)[:, None]
    cc = np.asarray(k, dtype=np.float64)[:, None, None] * c[:, :, None] * c[:, None, :]
    Ke = np.empty((members.shape[0], 6, 6))
    Ke[:, :3, :3] = cc
    Ke[:, 3:, 3:] = cc
    Ke[:, :3, 3:] = -cc
    Ke[:, 3:, :3] = -cc
    return Ke


def truss_shipped_case():
    """truss.py:6-24 geometry in float64: 3 nodes, 2 members, k=1000, load (0,-100,0) on node 2.
    Nodes without loads never move in the script, i.e. they are pinned (truss.py:112-119); z=0
    for every node so the z DOF of node 2 is constrained as well (2-D truss in 3-D storage)."""
    nodes = np.array([[0.0, 0.0, 0.0], [0.0, 1.0, 0.0], [1.0, 0.5, 0.0]])
    members = np.array([[0, 2], [1, 2]], dtype=np.int64)
    k = np.array([1000.0, 1000.0])
    constraints = np.ones((3, 3), dtype=np.int64)
    constraints[2, :2] = 0
    loads = np.zeros((3, 3))
    loads[2] = (0.0, -100.0, 0.0)
    return nodes, members, k, constraints, loads


LATTICE_DIRECTIONS = np.array(
    [
        [1, 0, 0], [0, 1, 0], [0, 0, 1],
        [1, 1, 0], [1, -1, 0], [1, 0, 1], [1, 0, -1], [0, 1, 1], [0, 1, -1],
        [1, 1, 1], [1, 1, -1], [1, -1, 1], [1, -1, -1],
    ],
    dtype=np.int64,
)


def lattice_truss_case(n: int, n_rhs: int = 64, h: float = 1.0):
    """BASELINE config 5, frozen in SURVEY.md §8(d): jittered cubic lattice of n^3 nodes, node id
    = (iz*n + iy)*n + ix, members along the 13 half-space neighbour directions (listed direction
    by direction, start nodes in id order) => rigid; jitter U(-0.1h, 0.1h) then k ~ U(500,1500)
    from default_rng(0); nodes with iz = 0 pinned; loads = default_rng(1).standard_normal((3N, n_rhs))."""
    rng = np.random.default_rng(0)
    iz, iy, ix = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    grid = np.stack([ix.ravel(), iy.ravel(), iz.ravel()], axis=1)
    nodes = grid * h + rng.uniform(-0.1 * h, 0.1 * h, size=(n**3, 3))
    members = []
    ids = np.arange(n**3, dtype=np.int64)
    for d in LATTICE_DIRECTIONS:
        tgt = grid + d
        ok = np.all((tgt >= 0) & (tgt < n), axis=1)
        members.append(np.stack([ids[ok], (tgt[ok, 2] * n + tgt[ok, 1]) * n + tgt[ok, 0]], axis=1))
    members = np.concatenate(members).astype(np.int64)
    k = rng.uniform(500.0, 1500.0, size=members.shape[0])
    constraints = np.zeros((n**3, 3), dtype=np.int64)
    constraints[grid[:, 2] == 0] = 1
    loads = np.random.default_rng(1).standard_normal((3 * n**3, n_rhs))
    return nodes, members, k, constraints, loads


def jacobi_pcg_multi(A: sp.csr_matrix, B: np.ndarray, tol: float = 1e-12, maxiter: int | None = None):
    """Batched multi-RHS Jacobi-PCG: the same recurrence as `jacobi_pcg`, independently per
    column (each column has its own alpha/beta and freezes once its own residual <= tol*||b||).
    Returns (X, iterations per column)."""
    n, k = B.shape
    if maxiter is None:
        maxiter = 10 * n
    dinv = 1.0 / A.diagonal()
    X = np.zeros((n, k))
    bnorm = np.linalg.norm(B, axis=0)
    R = B.astype(np.float64).copy()
    Z = dinv[:, None] * R
    P = Z.copy()
    rz = np.einsum("ij,ij->j", R, Z)
    active = bnorm > 0
    iters = np.zeros(k, dtype=np.int64)
    it = 0
    while it < maxiter and active.any():
        AP = A @ P
        pAp = np.einsum("ij,ij->j", P, AP)
        alpha = np.where(active, rz / np.where(active, pAp, 1.0), 0.0)
        X += alpha * P
        R -= alpha * AP
        it += 1
        iters[active] = it
        rel = np.linalg.norm(R, axis=0) / np.where(bnorm > 0, bnorm, 1.0)
        active = active & (rel > tol)
        Z = dinv[:, None] * R
        rz_new = np.einsum("ij,ij->j", R, Z)
        beta = np.where(active, rz_new / np.where(rz != 0, rz, 1.0), 0.0)
        P = Z + beta * P
        rz = rz_new
    return X, iters
