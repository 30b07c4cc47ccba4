"""CPU, world_size 2 (gloo): partition / halo exchange / all-reduce / convergence logic of
fea_b200.dist, with the step kernels replaced by numpy stand-ins (test infrastructure) that follow
the semantics documented in include/fea_b200.h.  The GPU kernels themselves are covered by the
`-m gpu` tests; `gpurun --gpus N` covers the NCCL path."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fea_b200 import _lib
from fea_b200 import dist as fdist
from oracle import fea_oracle as fo


class NumpyOps:
    """Stand-in for GpuOps on CPU tensors: same state layout, same step semantics."""

    def __init__(self, K_local, plan, d):
        lo = d * plan.offset
        self.K_owned = K_local.tocsr()[lo:lo + d * plan.n_owned]
        self.plan, self.d = plan, d

    @staticmethod
    def _ints(state):
        return state.view(torch.int32)

    def init(self, b, dinv, x, r, p_own, tol, max_iter, state):
        free = dinv != 0
        x.zero_()
        r.copy_(torch.where(free, b, torch.zeros_like(b)))
        p_own.copy_(dinv * r)
        state.zero_()
        state[_lib.PCG_RZ] = float(r @ p_own)
        state[_lib.PCG_BNORM2] = float(r @ r)
        state[_lib.PCG_RR] = float(r @ r)
        state[_lib.PCG_TOL2] = tol * tol
        self._ints(state)[_lib.PCG_MAXITER_I32] = max_iter

    def step_spmv(self, p_ext, ap, state):
        if self._ints(state)[_lib.PCG_DONE_I32]:
            return
        ap.copy_(torch.from_numpy(self.K_owned @ p_ext.numpy()))
        lo = self.d * self.plan.offset
        state[_lib.PCG_PAP] = float(ap @ p_ext[lo:lo + ap.numel()])

    def step_update(self, dinv, p_own, ap, x, r, state):
        i = self._ints(state)
        if i[_lib.PCG_DONE_I32]:
            return
        pap, rz, bn2 = float(state[_lib.PCG_PAP]), float(state[_lib.PCG_RZ]), float(state[_lib.PCG_BNORM2])
        if bn2 == 0.0 or not pap > 0.0:
            i[_lib.PCG_DONE_I32] = 1
            i[_lib.PCG_STATUS_I32] = _lib.FEA_OK if bn2 == 0.0 else _lib.FEA_ERR_BREAKDOWN
            state[_lib.PCG_RR_FINAL] = state[_lib.PCG_RR]
            return
        alpha = rz / pap
        x += alpha * p_own
        r -= alpha * ap
        free = dinv != 0
        state[_lib.PCG_RZ_NEW] = float((r * dinv) @ r)
        state[_lib.PCG_RR] = float(r[free] @ r[free])
        i[_lib.PCG_ITER_I32] += 1

    def step_direction(self, dinv, r, p_own, state):
        i = self._ints(state)
        if i[_lib.PCG_DONE_I32]:
            return
        if float(state[_lib.PCG_RR]) <= float(state[_lib.PCG_TOL2]) * float(state[_lib.PCG_BNORM2]):
            i[_lib.PCG_DONE_I32] = 1
            state[_lib.PCG_RR_FINAL] = state[_lib.PCG_RR]
            return
        beta = float(state[_lib.PCG_RZ_NEW]) / float(state[_lib.PCG_RZ])
        p_own.copy_(dinv * r + beta * p_own)
        state[_lib.PCG_RZ] = state[_lib.PCG_RZ_NEW]
        if int(i[_lib.PCG_ITER_I32]) >= int(i[_lib.PCG_MAXITER_I32]):
            i[_lib.PCG_DONE_I32] = 1
            i[_lib.PCG_STATUS_I32] = _lib.FEA_ERR_MAXITER
            state[_lib.PCG_RR_FINAL] = state[_lib.PCG_RR]


def local_problem(case, plan):
    nodes, elements, cons, forces = case
    g_lo, g_hi = plan.g_lo, plan.g_hi
    el = elements[plan.element_ids] - g_lo
    ln = nodes[g_lo:g_hi]
    K = fo.assemble_csr(el, fo.hex8_ke_batched(ln, el, fo.E_HEX, fo.NU_HEX), g_hi - g_lo, 3)
    lo, hi = 3 * plan.offset, 3 * (plan.offset + plan.n_owned)
    diag = K.diagonal()[lo:hi]
    fixed = cons[plan.own_lo:plan.own_hi].ravel() != 0
    dinv = np.where(fixed, 0.0, 1.0 / diag)
    b = forces[plan.own_lo:plan.own_hi].ravel().copy()
    return K, torch.from_numpy(b), torch.from_numpy(dinv)


def test_plan_invariants():
    nodes, elements, cons, forces = fo.cantilever_case(12, 3)
    layer = 16
    for world in (1, 2, 3, 4):
        cuts = fdist.node_cuts(nodes.shape[0], world, layer=layer)
        assert cuts[0] == 0 and cuts[-1] == nodes.shape[0] and np.all(np.diff(cuts) > 0)
        assert np.all(cuts % layer == 0)
        plans = [fdist.plan_slab(elements, cuts, r) for r in range(world)]
        covered = np.zeros(elements.shape[0], dtype=int)
        Kg = fo.assemble_csr(elements, fo.hex8_ke_batched(nodes, elements, 1.0, 0.3), nodes.shape[0], 3)
        for p in plans:
            covered[p.element_ids] += 1
            assert p.g_lo <= p.own_lo < p.own_hi <= p.g_hi
            assert p.own_lo - p.g_lo in (0, layer) and p.g_hi - p.own_hi in (0, layer)
            # matching send/recv ranges between neighbours
            if p.recv_up is not None:
                q = plans[p.rank + 1]
                assert q.send_down == (p.rank, p.recv_up[1], p.recv_up[2])
            if p.recv_down is not None:
                q = plans[p.rank - 1]
                assert q.send_up == (p.rank, p.recv_down[1], p.recv_down[2])
            # owned rows of the slab matrix are the global rows, bit for bit
            el = elements[p.element_ids] - p.g_lo
            Kl = fo.assemble_csr(el, fo.hex8_ke_batched(nodes[p.g_lo:p.g_hi], el, 1.0, 0.3), p.g_hi - p.g_lo, 3)
            rows = Kl[3 * p.offset:3 * (p.offset + p.n_owned)]
            ref = Kg[3 * p.own_lo:3 * p.own_hi, 3 * p.g_lo:3 * p.g_hi]
            assert np.array_equal(rows.indptr, ref.indptr) and np.array_equal(rows.indices, ref.indices)
            assert np.array_equal(rows.data, ref.data)
        assert covered.min() >= 1 and covered.max() <= 2
        # elements straddling a cut are assembled twice: b*b per interface
        assert covered.sum() - elements.shape[0] == 9 * (world - 1)


def test_default_cuts_are_balanced_and_line_aligned():
    """default_cuts: node-balanced, multiples of 16 nodes; plan_slab then starts the local range early so
    that the owned rows of a local vector begin and end on 128-byte boundaries (16 nodes x 3 doubles), the
    halo ranges stay exactly what the elements need, and neighbours' send / recv ranges still match."""
    nodes, elements, cons, forces = fo.cantilever_case(40, 8)  # 3321 nodes, 81 per layer
    Kg = fo.assemble_csr(elements, fo.hex8_ke_batched(nodes, elements, 1.0, 0.3), nodes.shape[0], 3)
    for world in (2, 3):
        cuts = fdist.default_cuts(nodes.shape[0], world)
        assert cuts[0] == 0 and cuts[-1] == nodes.shape[0] and np.all(cuts[1:-1] % 16 == 0)
        assert np.diff(cuts).max() - np.diff(cuts).min() <= 32
        assert np.any(cuts[1:-1] % 81 != 0)  # not layer-aligned: the cut runs through a node layer
        plans = [fdist.plan_slab(elements, cuts, r) for r in range(world)]
        for p in plans:
            assert p.offset % 16 == 0
            if p.rank < world - 1:
                assert (p.offset + p.n_owned) % 16 == 0
            if p.recv_up is not None:
                assert plans[p.rank + 1].send_down == (p.rank, p.recv_up[1], p.recv_up[2])
            if p.recv_down is not None:
                assert plans[p.rank - 1].send_up == (p.rank, p.recv_down[1], p.recv_down[2])
                assert p.g_lo <= p.recv_down[1] and p.recv_down[1] - p.g_lo < 16  # only the padding in front
            el = elements[p.element_ids] - p.g_lo
            assert el.min() >= 0 and el.max() < p.n_local
            Kl = fo.assemble_csr(el, fo.hex8_ke_batched(nodes[p.g_lo:p.g_hi], el, 1.0, 0.3), p.n_local, 3)
            rows = Kl[3 * p.offset:3 * (p.offset + p.n_owned)]
            ref = Kg[3 * p.own_lo:3 * p.own_hi, 3 * p.g_lo:3 * p.g_hi]
            assert np.array_equal(rows.indptr, ref.indptr) and np.array_equal(rows.indices, ref.indices)
            assert np.array_equal(rows.data, ref.data)
    # the multi-threaded host scan and the numpy path agree (Fortran-ordered input takes the numpy path)
    cuts = fdist.default_cuts(nodes.shape[0], 3)
    for r in range(3):
        a, b = fdist.plan_slab(elements, cuts, r), fdist.plan_slab(np.asfortranarray(elements), cuts, r)
        assert (a.g_lo, a.g_hi, a.send_down, a.send_up, a.recv_down, a.recv_up) == \
               (b.g_lo, b.g_hi, b.send_down, b.send_up, b.recv_down, b.recv_up)
        assert np.array_equal(a.element_ids, b.element_ids)
    # small meshes are not forced onto 16-node boundaries
    small = fdist.default_cuts(96, 4)
    assert list(small) == [0, 24, 48, 72, 96]


def test_thin_slab_rejected():
    nodes, elements, cons, forces = fo.cantilever_case(4, 2)
    # cuts inside a node layer: rank 0's halo would reach past rank 1 into rank 2
    with pytest.raises(ValueError):
        fdist.plan_slab(elements, np.array([0, 3, 6, nodes.shape[0]]), 0)


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        case = fo.cantilever_case(10, 3)
        nodes = case[0]
        cuts = fdist.node_cuts(nodes.shape[0], world, layer=16)
        plan = fdist.plan_slab(case[1], cuts, rank)
        K, b, dinv = local_problem(case, plan)
        ops = NumpyOps(K, plan, 3)
        x, info = fdist.distributed_pcg(ops, plan, 3, b, dinv, tol=1e-12, max_iter=5000, chunk=16)
        gathered = [None] * world
        dist.all_gather_object(gathered, (plan.own_lo, x.numpy(), info.iterations, info.rel_residual, info.status))
        if rank == 0:
            ret["parts"] = gathered
        # zero right-hand side terminates immediately with x = 0
        x0, info0 = fdist.distributed_pcg(ops, plan, 3, torch.zeros_like(b), dinv, tol=1e-12, max_iter=100, chunk=4)
        assert info0.status == 0 and float(x0.abs().max()) == 0.0 and info0.iterations == 0
        # halo exchange moves exactly the neighbour's owned values
        v = torch.zeros(3 * plan.n_local, dtype=torch.float64)
        v[3 * plan.offset:3 * (plan.offset + plan.n_owned)] = torch.arange(3 * plan.own_lo, 3 * plan.own_hi,
                                                                           dtype=torch.float64)
        fdist.HaloExchange(plan, 3)(v)
        assert torch.equal(v, torch.arange(3 * plan.g_lo, 3 * plan.g_hi, dtype=torch.float64))
        # gather of the owned rows on rank 0 (uneven slabs), as the collective solve() does for (u, K u)
        own = torch.arange(plan.own_lo, plan.own_hi, dtype=torch.float64)
        both = torch.stack([own, -own]).reshape(2, plan.n_owned, 1).repeat(1, 1, 3).contiguous()
        full = fdist.gather_rows(plan, cuts, both)
        if rank == 0:
            want = torch.arange(0, nodes.shape[0], dtype=torch.float64)
            assert full.shape == (2, nodes.shape[0], 3)
            assert torch.equal(full[0, :, 2], want) and torch.equal(full[1, :, 0], -want)
        else:
            assert full is None
        # the same solve on node-balanced, 16-node aligned cuts of a larger mesh (halo narrower than the padding)
        case2 = fo.cantilever_case(24, 6)
        cuts2 = fdist.default_cuts(case2[0].shape[0], world) if world == 2 else fdist.node_cuts(case2[0].shape[0], world)
        plan2 = fdist.plan_slab(case2[1], cuts2, rank)
        K2, b2, dinv2 = local_problem(case2, plan2)
        x2, info2 = fdist.distributed_pcg(NumpyOps(K2, plan2, 3), plan2, 3, b2, dinv2, tol=1e-12, max_iter=5000, chunk=16)
        parts2 = [None] * world
        dist.all_gather_object(parts2, (plan2.own_lo, x2.numpy(), info2.status))
        if rank == 0:
            ret["parts2"] = parts2
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_distributed_pcg_gloo(world):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    parts = sorted(ret["parts"], key=lambda t: t[0])
    u = np.concatenate([p[1] for p in parts])
    iters = {p[2] for p in parts}
    assert len(iters) == 1  # every rank stops at the same iteration
    assert all(p[4] == 0 and p[3] <= 1e-12 for p in parts)
    nodes, elements, cons, forces = fo.cantilever_case(10, 3)
    uo, _, io = fo.solve_hex8(nodes, elements, cons, forces, method="pcg", tol=1e-12)
    ud, _, _ = fo.solve_hex8(nodes, elements, cons, forces, method="direct")
    assert abs(iters.pop() - io["iterations"]) <= 3
    assert np.abs(u - uo.ravel()).max() <= 1e-10 * np.abs(uo).max()
    assert np.abs(u - ud.ravel()).max() <= 1e-8 * np.abs(ud).max()
    parts2 = sorted(ret["parts2"], key=lambda t: t[0])
    assert all(p[2] == 0 for p in parts2)
    n2, e2, c2, f2 = fo.cantilever_case(24, 6)
    ud2, _, _ = fo.solve_hex8(n2, e2, c2, f2, method="direct")
    u2 = np.concatenate([p[1] for p in parts2])
    assert np.abs(u2 - ud2.ravel()).max() <= 1e-8 * np.abs(ud2).max()
