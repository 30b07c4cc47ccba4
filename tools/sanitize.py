"""Small end-to-end pass over every kernel family for compute-sanitizer (memcheck / racecheck /
synccheck):  symbolic pass, hex8 / beam / truss assembly, Ke kernels, TMA SpMV, both PCG recurrences
(last-block reductions), multi-RHS SpMM + batched PCG, the device relaxation loop, mesh builders.

    compute-sanitizer --tool racecheck python tools/sanitize.py
    torchrun --nproc-per-node 2 tools/sanitize.py --dist        (under compute-sanitizer per rank)

Sizes are tiny (the tools slow kernels down 10-100x); results are still checked against the oracle.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fea_b200 import core, cubebeam, euler_bernoulli, truss, utils  # noqa: E402
from oracle import fea_oracle as fo  # noqa: E402

E, NU = fo.E_HEX, fo.NU_HEX


def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300))


def single():
    nodes, elements, cons, forces = fo.cantilever_case(10, 3)
    uo, fo_, io = fo.solve_hex8(nodes, elements, cons, forces, method="direct")
    for algo in ("0", "1"):
        os.environ["FEA_PCG_ALGO"] = algo
        u, f = cubebeam.solve(nodes, elements, cons, forces)
        assert rel(u, uo) < 1e-8, rel(u, uo)
    del os.environ["FEA_PCG_ALGO"]
    ke = utils.hexahedral_stiffness_matrix(nodes[elements[0]], E, NU)
    assert rel(ke, fo.hex8_ke(nodes[elements[0]], E, NU)) < 1e-10
    # SpMV / SpMM on the assembled matrix
    nd, el = core.to_device(nodes, torch.float64), core.to_device(elements, torch.int32)
    K = core.assemble_hex8(nd, el, E, NU)
    x = np.random.default_rng(0).standard_normal((K.n_dof, 5))
    assert rel(K.matvec(core.to_device(x[:, 0].copy(), torch.float64)).cpu().numpy(), io["K"] @ x[:, 0]) < 1e-12
    assert rel(K.matmat(core.to_device(x, torch.float64)).cpu().numpy(), io["K"] @ x) < 1e-12
    # truss: linear multi-RHS solve and the relaxation loop
    tn, tm, tk, tc, tl = truss.lattice_truss(5, n_rhs=6)
    X = truss.solve_linear(tn, tm, tk, tc, tl)
    Kt = fo.assemble_csr(tm, fo.truss_ke_batched(tn, tm, tk), tn.shape[0], 3)
    free = fo.free_dofs(tc)
    import scipy.sparse.linalg as spla

    Xo = np.zeros_like(tl)
    Xo[free] = spla.spsolve(Kt[free][:, free].tocsc(), tl[free])
    assert rel(X, Xo) < 1e-8
    if hasattr(truss, "relax_device"):
        hist, _ = truss.relax_device(steps=10)
        assert abs(hist[1] - 60.08) < 0.05, hist[:3]
    # beam
    if hasattr(euler_bernoulli, "run"):
        euler_bernoulli.run()
    print("SANITIZE single-GPU pass ok")


def distributed():
    import torch.distributed as dist

    from fea_b200 import dist as fdist

    rank = int(os.environ["RANK"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    nodes, elements, cons, forces = fo.cantilever_case(16, 3)
    uo, _, _ = fo.solve_hex8(nodes, elements, cons, forces, method="direct")
    for algo in ("0", "1"):
        os.environ["FEA_PCG_ALGO"] = algo
        u, f = cubebeam.solve(nodes, elements, cons, forces)
        if rank == 0:
            assert rel(u, uo) < 1e-8, rel(u, uo)
    assert fdist.SOLVER_USED["kind"] == os.environ.get("FEA_DIST_COMM", "p2p"), fdist.SOLVER_USED
    print(f"SANITIZE 2-rank pass ok (rank {rank}, solver {fdist.SOLVER_USED['kind']})")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    torch.cuda.init()
    if "--dist" in sys.argv:
        distributed()
    else:
        single()
