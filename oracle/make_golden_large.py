"""Oracle fixtures at BASELINE sizes: tests/golden/oracle_config3.npz, oracle_config4.npz.

    python oracle/make_golden_large.py config3        # ~1 min  (100x20x20, 133,623 DOF)
    python oracle/make_golden_large.py config4        # ~1 h on 8 cores, ~25 GB (400x80x80, 7.9 M DOF)

TEST INFRASTRUCTURE.  The reference's dense solve() cannot hold these sizes (cubebeam.py:80:
143 GB / 498 TB), so what is recorded here is the output of the ORACLE -- oracle/fea_oracle_c.c,
the C twin of the reference algorithm that tests/test_oracle_golden.py pins to the
reference-generated K1..K8 fixtures -- run offline because it takes minutes to an hour on CPU:
Ke in the reference's dense order (utils.py:127-239), scatter (cubebeam.py:82-90), reduction
(cubebeam.py:92-96), Jacobi-PCG to a 1e-12 recurrence residual, expansion and K_full @ u
(cubebeam.py:102-106).  Config 3 also records scipy's sparse LU solution of the same reduced
system as an independent check.  Config 4 keeps a seeded sample of u (the full vector is 63 MB).
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import c_oracle as co  # noqa: E402
from oracle import fea_oracle as fo  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
SAMPLE = 20000


def solve(A: int, b: int):
    t0 = time.perf_counter()
    nodes, elements, cons, forces = fo.cantilever_case(A, b)
    pattern = co.dof_pattern(elements, nodes.shape[0], 3)
    K = co.assemble_hex8(nodes, elements, fo.E_HEX, fo.NU_HEX, pattern=pattern)
    free = fo.free_dofs(cons)
    Kff = co.reduce_csr(K, free)
    ff = forces.flatten()[free]
    t1 = time.perf_counter()
    print(f"{A}x{b}x{b}: assembled nnz {K.nnz} in {t1 - t0:.1f}s on {co.threads()} threads", flush=True)
    uf, it, rel = co.jacobi_pcg(Kff, ff, tol=1e-12)
    t2 = time.perf_counter()
    print(f"PCG: {it} iterations, rel residual {rel:.3e}, {t2 - t1:.1f}s", flush=True)
    u = np.zeros(K.shape[0])
    u[free] = uf
    f_out = co.spmv(K, u)
    true_rel = float(np.linalg.norm(ff - co.spmv(Kff, uf)) / np.linalg.norm(ff))
    return nodes, cons, free, K, Kff, ff, u, f_out, it, rel, true_rel, (t1 - t0, t2 - t1)


def config3():
    nodes, cons, free, K, Kff, ff, u, f_out, it, rel, true_rel, secs = solve(100, 20)
    import scipy.sparse.linalg as spla

    t0 = time.perf_counter()
    ud = np.zeros(K.shape[0])
    ud[free] = spla.spsolve(Kff.tocsc(), ff)
    print(f"sparse LU: {time.perf_counter() - t0:.1f}s, |u_pcg - u_lu|/|u|max = "
          f"{np.abs(u - ud).max() / np.abs(ud).max():.3e}", flush=True)
    base = nodes[:, 2] == 0
    np.savez_compressed(
        os.path.join(OUT, "oracle_config3.npz"), A=100, b=20, u_pcg=u, u_direct=ud, iterations=it, rel_residual=rel,
        true_rel_residual=true_rel, sum_reactions_y=f_out.reshape(-1, 3)[base, 1].sum(), threads=co.threads(),
        seconds=np.array(secs))


def config4():
    nodes, cons, free, K, Kff, ff, u, f_out, it, rel, true_rel, secs = solve(400, 80)
    idx = np.sort(np.random.default_rng(0).choice(u.size, SAMPLE, replace=False))
    base = nodes[:, 2] == 0
    np.savez_compressed(
        os.path.join(OUT, "oracle_config4.npz"), A=400, b=80, sample_index=idx, u_sample=u[idx],
        f_sample=f_out[idx], u_norm2=np.linalg.norm(u), u_max_abs=np.abs(u).max(), u_argmax=int(np.abs(u).argmax()),
        iterations=it, rel_residual=rel, true_rel_residual=true_rel,
        sum_reactions_y=f_out.reshape(-1, 3)[base, 1].sum(), threads=co.threads(), seconds=np.array(secs))


def config5(n: int = 30):
    """Config-5-shaped lattice truss at n = 30 (27,000 nodes, 78,300 free DOF, 64 load
    cases): scipy's sparse LU of the reduced oracle matrix, every column.  Keeps a seeded row sample."""
    import scipy.sparse.linalg as spla

    nodes, members, k, cons, B = fo.lattice_truss_case(n, n_rhs=64)[:5]
    K = fo.assemble_csr(members, fo.truss_ke_batched(nodes, members, k), nodes.shape[0], 3)
    free = fo.free_dofs(cons)
    t0 = time.perf_counter()
    lu = spla.splu(K[free][:, free].tocsc())
    X = np.zeros_like(B)
    X[free] = lu.solve(B[free])
    print(f"sparse LU of {free.size} DOF x 64 RHS: {time.perf_counter() - t0:.1f}s", flush=True)
    rows = np.sort(np.random.default_rng(0).choice(B.shape[0], 4000, replace=False))
    np.savez_compressed(os.path.join(OUT, "oracle_config5_n30.npz"), n=n, rows=rows, X_rows=X[rows],
                        col_norms=np.linalg.norm(X, axis=0), col_max=np.abs(X).max(axis=0))


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "config3"
    {"config3": config3, "config4": config4, "config5": config5}[which]()
    print("wrote", which)
