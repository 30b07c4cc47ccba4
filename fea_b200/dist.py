"""Multi-GPU solve: 1-D slab (row-block) partition, one process per GPU, torch.distributed.

The reference is single-process; this is the partitioning `north_star` asks for.
`stack_faces_2d` numbers nodes and elements layer-major (utils.py:363-374), so a contiguous range
of nodes is a contiguous block of DOF rows and (nearly) a contiguous block of elements: no
renumbering, DOF numbering stays the reference's.

  * rank r owns the node range [own_lo, own_hi); it assembles every element that touches an owned
    node (elements straddling a cut are evaluated by both neighbours), so assembly needs no
    communication and every owned row is complete and bit-identical to the single-GPU row;
  * its local node range [g_lo, g_hi) adds the halo nodes those elements reference;
  * per PCG iteration: one halo exchange of p with the two neighbours (grouped send/recv), then
    the three step kernels of include/fea_b200.h with the dot products all-reduced in between:
    [PAP] after step 1, [RZ_NEW, RR] after step 2 (two adjacent doubles, one collective).

The driver below is backend-agnostic: `ops` supplies the step kernels (GpuOps = the C ABI), the
collectives go through torch.distributed (NCCL on GPUs; gloo in the CPU tests, where the test
suite plugs in numpy ops to check the partition / halo / reduction logic against the oracle).
"""
from __future__ import annotations

import json
import os
import time
from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist

from . import _lib

STATE_DOUBLES = _lib.PCG_STATE_BYTES // 8


# ------------------------------------------------------------------------------------------------
# partition
# ------------------------------------------------------------------------------------------------
@dataclass
class SlabPlan:
    rank: int
    world: int
    own_lo: int          # owned node range (global ids)
    own_hi: int
    g_lo: int            # local node range = owned + halos (global ids)
    g_hi: int
    element_ids: np.ndarray  # global ids of the elements this rank assembles (ascending)
    # (peer rank, first global node, one-past-last global node)
    send_down: tuple | None  # my lowest owned nodes that rank-1 needs
    send_up: tuple | None    # my highest owned nodes that rank+1 needs
    recv_down: tuple | None  # my lower halo, owned by rank-1
    recv_up: tuple | None    # my upper halo, owned by rank+1

    @property
    def n_owned(self) -> int:
        return self.own_hi - self.own_lo

    @property
    def n_local(self) -> int:
        return self.g_hi - self.g_lo

    @property
    def offset(self) -> int:
        return self.own_lo - self.g_lo


def node_cuts(n_nodes: int, world: int, layer: int | None = None) -> np.ndarray:
    """world+1 cut points of [0, n_nodes); multiples of `layer` (nodes per mesh layer) if given."""
    if layer:
        n_layers = n_nodes // layer
        assert n_layers * layer == n_nodes, "n_nodes must be a multiple of the layer size"
        cuts = (np.arange(world + 1) * n_layers) // world * layer
    else:
        cuts = (np.arange(world + 1) * n_nodes) // world
    return cuts.astype(np.int64)


def _halo_extent(elements: np.ndarray, lo: int, hi: int):
    """Elements touching [lo, hi) and the node range they span."""
    touch = ((elements >= lo) & (elements < hi)).any(axis=1)
    ids = np.nonzero(touch)[0]
    if ids.size == 0:
        return ids, lo, hi
    sub = elements[ids]
    return ids, min(int(sub.min()), lo), max(int(sub.max()) + 1, hi)


def plan_slab(elements: np.ndarray, cuts: np.ndarray, rank: int) -> SlabPlan:
    """Everything rank `rank` needs to know, computed locally and identically on every rank."""
    world = len(cuts) - 1
    own_lo, own_hi = int(cuts[rank]), int(cuts[rank + 1])
    ids, g_lo, g_hi = _halo_extent(elements, own_lo, own_hi)
    recv_down = recv_up = send_down = send_up = None
    if g_lo < own_lo:
        if rank == 0 or g_lo < cuts[rank - 1]:
            raise ValueError("slab thinner than the mesh bandwidth: halo spans more than one neighbour")
        recv_down = (rank - 1, g_lo, own_lo)
    if g_hi > own_hi:
        if rank == world - 1 or g_hi > cuts[rank + 2]:
            raise ValueError("slab thinner than the mesh bandwidth: halo spans more than one neighbour")
        recv_up = (rank + 1, own_hi, g_hi)
    # what the neighbours need from me = their halo ranges
    if rank > 0:
        _, _, nb_hi = _halo_extent(elements, int(cuts[rank - 1]), own_lo)
        if nb_hi > own_lo:
            send_down = (rank - 1, own_lo, nb_hi)
    if rank < world - 1:
        _, nb_lo, _ = _halo_extent(elements, own_hi, int(cuts[rank + 2]))
        if nb_lo < own_hi:
            send_up = (rank + 1, nb_lo, own_hi)
    return SlabPlan(rank, world, own_lo, own_hi, g_lo, g_hi, ids, send_down, send_up, recv_down, recv_up)


# ------------------------------------------------------------------------------------------------
# halo exchange + distributed PCG driver (backend-agnostic)
# ------------------------------------------------------------------------------------------------
class HaloExchange:
    """Grouped send/recv of the boundary node values of a local vector (d values per node)."""

    def __init__(self, plan: SlabPlan, d: int, group=None):
        self.plan, self.d, self.group = plan, d, group

    def _slice(self, vec: torch.Tensor, lo: int, hi: int) -> torch.Tensor:
        g = self.plan.g_lo
        return vec[(lo - g) * self.d:(hi - g) * self.d]

    def __call__(self, vec_ext: torch.Tensor) -> None:
        pl = self.plan
        if pl.world == 1:
            return
        ops = []
        # fixed global order (all "up" transfers, then all "down") keeps send/recv pairs matched
        if pl.send_up is not None:
            ops.append(dist.P2POp(dist.isend, self._slice(vec_ext, pl.send_up[1], pl.send_up[2]), pl.send_up[0],
                                  group=self.group))
        if pl.recv_down is not None:
            ops.append(dist.P2POp(dist.irecv, self._slice(vec_ext, pl.recv_down[1], pl.recv_down[2]), pl.recv_down[0],
                                  group=self.group))
        if pl.send_down is not None:
            ops.append(dist.P2POp(dist.isend, self._slice(vec_ext, pl.send_down[1], pl.send_down[2]), pl.send_down[0],
                                  group=self.group))
        if pl.recv_up is not None:
            ops.append(dist.P2POp(dist.irecv, self._slice(vec_ext, pl.recv_up[1], pl.recv_up[2]), pl.recv_up[0],
                                  group=self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()


@dataclass
class DistInfo:
    iterations: int
    rel_residual: float
    status: int
    bnorm: float


def distributed_pcg(ops, plan: SlabPlan, d: int, b_owned: torch.Tensor, dinv_owned: torch.Tensor, tol: float = 1e-12,
                    max_iter: int = 100000, chunk: int = 32, group=None):
    """Jacobi-PCG over all ranks.  `ops` provides init / step_spmv / step_update / step_direction
    with the semantics of include/fea_b200.h; vectors live wherever `b_owned` lives.
    Returns (x_owned, DistInfo)."""
    dev = b_owned.device
    n_own = plan.n_owned * d
    x = torch.empty(n_own, dtype=torch.float64, device=dev)
    r = torch.empty(n_own, dtype=torch.float64, device=dev)
    ap = torch.empty(n_own, dtype=torch.float64, device=dev)
    p_ext = torch.zeros(plan.n_local * d, dtype=torch.float64, device=dev)
    p_own = p_ext[plan.offset * d:plan.offset * d + n_own]
    state = torch.zeros(STATE_DOUBLES, dtype=torch.float64, device=dev)
    state_i = state.view(torch.int32)
    halo = HaloExchange(plan, d, group)
    multi = plan.world > 1

    ops.init(b_owned, dinv_owned, x, r, p_own, tol, max_iter, state)
    if multi:
        dist.all_reduce(state[_lib.PCG_RZ:_lib.PCG_BNORM2 + 1], group=group)
    snap = [torch.empty(STATE_DOUBLES, dtype=torch.float64).pin_memory() if dev.type == "cuda"
            else torch.empty(STATE_DOUBLES, dtype=torch.float64) for _ in range(2)]
    events = [torch.cuda.Event() for _ in range(2)] if dev.type == "cuda" else None
    pending = [False, False]
    done_iter, slot, finished = 0, 0, False

    def run_iterations(count: int) -> None:
        for _ in range(count):
            halo(p_ext)
            ops.step_spmv(p_ext, ap, state)
            if multi:
                dist.all_reduce(state[_lib.PCG_PAP:_lib.PCG_PAP + 1], group=group)
            ops.step_update(dinv_owned, p_own, ap, x, r, state)
            if multi:
                dist.all_reduce(state[_lib.PCG_RZ_NEW:_lib.PCG_RR + 1], group=group)
            ops.step_direction(dinv_owned, r, p_own, state)

    # On GPUs a chunk of iterations (kernels + NCCL send/recv + all-reduces, all with fixed
    # arguments) is captured once into a CUDA graph and replayed: at 8 ranks a rank's kernels take
    # ~0.13 ms per iteration, less than the host needs to issue 3 launches and 3 collectives.
    graph = None
    # Off by default: measured at 2 ranks the captured NCCL ops cost ~90 us more per iteration than
    # plain launches (5.5 s vs 4.8 s per solve); it only pays when a rank is host-bound.
    if dev.type == "cuda" and max_iter > chunk and os.environ.get("FEA_DIST_GRAPH", "0") == "1":
        run_iterations(1)  # warm-up outside capture (lazy NCCL / kernel-attribute initialisation)
        done_iter += 1
        torch.cuda.synchronize()
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                run_iterations(chunk)
            graph = g
        except Exception as exc:  # capture unsupported in this setup: plain launches
            if plan.rank == 0:
                print(f"fea_b200.dist: CUDA-graph capture failed ({type(exc).__name__}: {exc}); using plain launches",
                      flush=True)
            graph = None
            torch.cuda.synchronize()

    while not finished:
        todo = min(chunk, max_iter - done_iter)
        if graph is not None and todo == chunk:
            graph.replay()
        else:
            run_iterations(todo)
        done_iter += todo
        snap[slot].copy_(state, non_blocking=True)
        if events is not None:
            events[slot].record()
        pending[slot] = True
        prev = slot ^ 1
        if pending[prev]:
            if events is not None:
                events[prev].synchronize()
            pending[prev] = False
            if int(snap[prev].view(torch.int32)[_lib.PCG_DONE_I32]) != 0:
                finished = True
        if events is None and int(snap[slot].view(torch.int32)[_lib.PCG_DONE_I32]) != 0:
            finished = True
        if done_iter >= max_iter:
            finished = True
        slot ^= 1
    if dev.type == "cuda":
        torch.cuda.current_stream().synchronize()
    final = state.cpu()
    fi = final.view(torch.int32)
    bn2 = float(final[_lib.PCG_BNORM2])
    status = int(fi[_lib.PCG_STATUS_I32])
    if int(fi[_lib.PCG_DONE_I32]) == 0 and status == _lib.FEA_OK:
        status = _lib.FEA_ERR_MAXITER
    done = int(fi[_lib.PCG_DONE_I32]) != 0
    rr = float(final[_lib.PCG_RR_FINAL if done else _lib.PCG_RR])
    info = DistInfo(int(fi[_lib.PCG_ITER_I32]), float(np.sqrt(rr / bn2)) if bn2 > 0 else 0.0,
                    status, float(np.sqrt(bn2)))
    return x, info


# ------------------------------------------------------------------------------------------------
# GPU backend: the C ABI
# ------------------------------------------------------------------------------------------------
class GpuOps:
    """Step kernels of libfea_b200.so on this rank's slab matrix (a core.BlockCSR over the local
    node range; only the owned rows are used)."""

    def __init__(self, K, plan: SlabPlan):
        from . import core

        self.core = core
        self.lib = _lib.load()
        self.K, self.plan, self.d = K, plan, K.dof_per_node
        self.partials = torch.empty(2 * _lib.PCG_PARTIALS, dtype=torch.float64, device=K.values.device)
        # owned rows = node_rowptr shifted by the lower-halo node count (entries stay absolute)
        self.rowptr_owned = K.pattern.node_rowptr[plan.offset:]

    def init(self, b, dinv, x, r, p_own, tol, max_iter, state):
        _lib.check(self.lib.fea_pcg_init(b.numel(), b.data_ptr(), dinv.data_ptr(), x.data_ptr(), r.data_ptr(),
                                         p_own.data_ptr(), float(tol), int(max_iter), state.data_ptr(),
                                         self.partials.data_ptr(), self.core._stream()), "fea_pcg_init")

    def step_spmv(self, p_ext, ap, state):
        pt = self.K.pattern
        _lib.check(self.lib.fea_pcg_step_spmv(self.plan.n_owned, self.d, self.rowptr_owned.data_ptr(),
                                              pt.node_colidx.data_ptr(), self.K.values.data_ptr(), pt.max_coupled,
                                              p_ext.data_ptr(), ap.data_ptr(), self.plan.offset, state.data_ptr(),
                                              self.partials.data_ptr(), self.core._stream()), "fea_pcg_step_spmv")

    def step_update(self, dinv, p_own, ap, x, r, state):
        _lib.check(self.lib.fea_pcg_step_update(x.numel(), dinv.data_ptr(), p_own.data_ptr(), ap.data_ptr(),
                                                x.data_ptr(), r.data_ptr(), state.data_ptr(),
                                                self.partials.data_ptr(), self.core._stream()), "fea_pcg_step_update")

    def step_direction(self, dinv, r, p_own, state):
        _lib.check(self.lib.fea_pcg_step_direction(r.numel(), dinv.data_ptr(), r.data_ptr(), p_own.data_ptr(),
                                                   state.data_ptr(), None, self.core._stream()),
                   "fea_pcg_step_direction")

    def matvec_owned(self, x_ext, y_owned):
        pt = self.K.pattern
        _lib.check(self.lib.fea_spmv(self.plan.n_owned, self.d, self.rowptr_owned.data_ptr(),
                                     pt.node_colidx.data_ptr(), self.K.values.data_ptr(), pt.max_coupled,
                                     x_ext.data_ptr(), y_owned.data_ptr(), self.core._stream()), "fea_spmv")


class P2PComm:
    """This rank's communication block (header + halo-extended p vector) and the IPC mappings of
    every peer's block, for fea_pcg_solve_p2p.  Cached per slab plan: allocation and the handle
    exchange happen once, later solves only bump the epoch."""

    _cache: dict = {}
    _warned = False

    @classmethod
    def get(cls, plan: SlabPlan, d: int, group=None) -> "P2PComm":
        key = (plan.rank, plan.world, plan.own_lo, plan.own_hi, plan.g_lo, plan.g_hi, d)
        if key not in cls._cache:
            cls._cache[key] = cls(plan, d, group)
        return cls._cache[key]

    def __init__(self, plan: SlabPlan, d: int, group=None):
        import ctypes

        lib = _lib.load()
        self.lib, self.plan, self.d, self.epoch = lib, plan, d, 0
        self.available, self.why = True, ""
        self.own, self.ptrs, self.g_los = None, [], []

        def agree(ok: bool, why: str) -> bool:
            """Every rank must take the same path: all-reduce the local outcome."""
            flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if int(flag) == 0:
                self.available, self.why = False, why if not ok else "a peer rank could not set up CUDA IPC"
            return self.available

        handle = (ctypes.c_ubyte * 64)()
        ok, why = True, ""
        try:
            nbytes = lib.fea_comm_bytes(plan.n_local * d)
            own = ctypes.c_void_p()
            _lib.check(lib.fea_comm_alloc(nbytes, ctypes.byref(own)), "fea_comm_alloc")
            self.own = own.value
            _lib.check(lib.fea_comm_ipc_export(self.own, ctypes.addressof(handle)), "fea_comm_ipc_export")
        except Exception as exc:  # noqa: BLE001 -- reported, then every rank falls back together
            ok, why = False, f"{type(exc).__name__}: {exc}"
        if not agree(ok, why):
            return
        infos = [None] * plan.world
        dist.all_gather_object(infos, (bytes(handle), plan.g_lo), group=group)
        self.g_los = [g for _, g in infos]
        try:
            for r, (h, _) in enumerate(infos):
                if r == plan.rank:
                    self.ptrs.append(self.own)
                    continue
                buf = (ctypes.c_ubyte * 64).from_buffer_copy(h)
                p = ctypes.c_void_p()
                _lib.check(lib.fea_comm_ipc_open(ctypes.addressof(buf), ctypes.byref(p)), "fea_comm_ipc_open")
                self.ptrs.append(p.value)
        except Exception as exc:  # noqa: BLE001
            ok, why = False, f"{type(exc).__name__}: {exc}"
        if not agree(ok, why):
            return
        dist.barrier(group=group)  # every block is mapped everywhere before anyone writes

    def descriptor(self) -> "_lib.PeerComm":
        pl = self.plan
        self.epoch += 1
        c = _lib.PeerComm()
        c.world, c.rank = pl.world, pl.rank
        c.lower_peer = pl.send_down[0] if pl.send_down is not None else -1
        c.upper_peer = pl.send_up[0] if pl.send_up is not None else -1
        for r, p in enumerate(self.ptrs):
            c.comm[r] = p
        c.own_offset_nodes = pl.offset
        if pl.send_down is not None:
            peer, lo, hi = pl.send_down
            c.send_lower_first, c.send_lower_count, c.send_lower_dst = lo - pl.own_lo, hi - lo, lo - self.g_los[peer]
        if pl.send_up is not None:
            peer, lo, hi = pl.send_up
            c.send_upper_first, c.send_upper_count, c.send_upper_dst = lo - pl.own_lo, hi - lo, lo - self.g_los[peer]
        c.epoch = self.epoch
        return c


def p2p_pcg(K, plan: SlabPlan, b_owned: torch.Tensor, dinv_owned: torch.Tensor, tol: float, max_iter: int, group=None):
    """Distributed Jacobi-PCG through fea_pcg_solve_p2p (NVLink peer memory, no NCCL per iteration)."""
    import ctypes

    from . import core

    lib = _lib.load()
    d = K.dof_per_node
    comm = P2PComm.get(plan, d, group)
    desc = comm.descriptor()
    n = plan.n_owned * d
    x = torch.empty(n, dtype=torch.float64, device=b_owned.device)
    ws_bytes = lib.fea_pcg_workspace(n)
    work = torch.empty(ws_bytes, dtype=torch.uint8, device=b_owned.device)
    rowptr_owned = K.pattern.node_rowptr[plan.offset:]
    res = _lib.PcgResult()
    pt = K.pattern
    _lib.check(lib.fea_pcg_solve_p2p(plan.n_owned, d, rowptr_owned.data_ptr(), pt.node_colidx.data_ptr(),
                                     K.values.data_ptr(), pt.max_coupled, dinv_owned.data_ptr(), b_owned.data_ptr(),
                                     x.data_ptr(), float(tol), int(max_iter), work.data_ptr(), ws_bytes,
                                     ctypes.byref(desc), ctypes.byref(res), core._stream()), "fea_pcg_solve_p2p")
    if res.status == _lib.FEA_ERR_PEER:
        raise _lib.FeaLibraryError("fea_pcg_solve_p2p: a peer rank never delivered its halo / partial sum")
    return x, DistInfo(res.iterations, res.rel_residual, res.status, res.bnorm)


def solve_hex8_slab(nodes, elements, constraints, forces, E, nu, plan: SlabPlan, tol=1e-12, max_iter=None,
                    group=None):
    """This rank's share of solve(nodes, elements, constraints, forces) (cubebeam.py:79-108).
    All arguments are the GLOBAL host arrays (each rank slices its slab).  Returns
    (u_owned (n_owned,3) device, reactions_owned device, DistInfo, BlockCSR)."""
    from . import core

    prof = STAGE_PROFILE if STAGE_PROFILE.get("enabled") else None

    def mark(name, t0):
        if prof is None:
            return t0
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        prof[name] = prof.get(name, 0.0) + (t1 - t0) * 1e3
        return t1

    t = time.perf_counter()
    g_lo, g_hi = plan.g_lo, plan.g_hi
    nodes_d = core.to_device(np.ascontiguousarray(nodes[g_lo:g_hi]), torch.float64)
    ids = plan.element_ids
    if ids.size and int(ids[-1]) - int(ids[0]) + 1 == ids.size:
        # slabs of a layer-major mesh own a contiguous element range: a view, shifted on the device
        el_view = np.ascontiguousarray(np.asarray(elements)[int(ids[0]):int(ids[-1]) + 1])
        elements_d = (core.to_device(el_view, torch.int64) - g_lo).to(torch.int32)
    else:
        elements_d = core.to_device(np.ascontiguousarray(np.asarray(elements)[ids] - g_lo), torch.int32)
    fixed = core._fixed_mask(np.ascontiguousarray(np.asarray(constraints)[g_lo:g_hi]), 3 * plan.n_local)
    t = mark("slice_h2d", t)
    K = core.assemble_hex8(nodes_d, elements_d, E, nu, fixed=fixed)
    lo, hi = 3 * plan.offset, 3 * (plan.offset + plan.n_owned)
    dinv_owned = K.dinv[lo:hi].contiguous()
    b_owned = core.to_device(np.ascontiguousarray(np.asarray(forces)[plan.own_lo:plan.own_hi]),
                             torch.float64).reshape(-1)
    t = mark("symbolic_assembly", t)
    ops = GpuOps(K, plan)
    if max_iter is None:
        max_iter = 10 * 3 * int(np.asarray(nodes).shape[0])
    # NVLink peer-memory solver by default; FEA_DIST_COMM=nccl selects the torch.distributed loop, which
    # is also what every rank falls back to (together) when CUDA IPC cannot be set up between the ranks
    use_p2p = plan.world > 1 and os.environ.get("FEA_DIST_COMM", "p2p") == "p2p"
    if use_p2p and not P2PComm.get(plan, 3, group).available:
        if plan.rank == 0 and not P2PComm._warned:
            print(f"fea_b200.dist: peer-memory solver unavailable ({P2PComm.get(plan, 3, group).why}); "
                  "using the NCCL driver", flush=True)
            P2PComm._warned = True
        use_p2p = False
    SOLVER_USED["kind"] = "p2p" if use_p2p else ("nccl" if plan.world > 1 else "single")
    if use_p2p:
        x, info = p2p_pcg(K, plan, b_owned, dinv_owned, tol, max_iter, group=group)
    else:
        x, info = distributed_pcg(ops, plan, 3, b_owned, dinv_owned, tol=tol, max_iter=max_iter, group=group)
    t = mark("pcg", t)
    # reactions: K_full u on the owned rows needs u on the halo
    u_ext = torch.zeros(3 * plan.n_local, dtype=torch.float64, device=x.device)
    u_ext[lo:hi] = x
    HaloExchange(plan, 3, group)(u_ext)
    reactions = torch.empty_like(x)
    ops.matvec_owned(u_ext, reactions)
    mark("reactions", t)
    return x.reshape(-1, 3), reactions.reshape(-1, 3), info, K


# which distributed solver the last solve_hex8_slab call used ("p2p" | "nccl" | "single")
SOLVER_USED: dict = {"kind": None}

# stage timings of solve_hex8_slab (ms, accumulated; each stage ends with a device synchronise) when
# STAGE_PROFILE["enabled"] is set -- bench.py switches it on for ONE extra untimed step.
STAGE_PROFILE: dict = {"enabled": False}


# ------------------------------------------------------------------------------------------------
# bench.py entry for N > 1 (launched by torchrun, one rank per GPU)
# ------------------------------------------------------------------------------------------------
def bench_entry(args, A, b, tol, E, nu, measured_peaks, ClockSampler, cpu_sample):
    from . import core, cubebeam

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _lib.load()
    nodes, elements, constraints, forces = cubebeam.cantilever_case(A, b)
    n_free = int((constraints == 0).sum())
    cuts = node_cuts(nodes.shape[0], world, layer=(b + 1) ** 2)
    plan = plan_slab(elements, cuts, rank)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    out = {}

    def step():
        u, react, info, K = solve_hex8_slab(nodes, elements, constraints, forces, E, nu, plan, tol=tol)
        out.update(info=info, K=K)

    for _ in range(args.warmup):
        step()
    barrier()
    lib.fea_profile_enable(1)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([ev0.elapsed_time(ev1) / args.steps], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms)
    prof = (4 * __import__("ctypes").c_double)()
    lib.fea_profile_read(prof)
    # one more, untimed, step with a synchronise after every stage: where a step's time goes
    STAGE_PROFILE.clear()
    STAGE_PROFILE["enabled"] = True
    barrier()
    step()
    STAGE_PROFILE["enabled"] = False
    stages = {k: round(v, 3) for k, v in STAGE_PROFILE.items() if k != "enabled"}
    info, K = out["info"], out["K"]
    # SpMV kernel alone on this rank's slab (CUDA events, after the timed region)
    p_ext = torch.randn(3 * plan.n_local, dtype=torch.float64, device="cuda")
    y = torch.empty(3 * plan.n_owned, dtype=torch.float64, device="cuda")
    ops = GpuOps(K, plan)
    for _ in range(3):
        ops.matvec_owned(p_ext, y)
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    reps = 20
    for _ in range(reps):
        ops.matvec_owned(p_ext, y)
    c.record()
    torch.cuda.synchronize()
    spmv_ms = a.elapsed_time(c) / reps
    own_nnz = 9 * int(K.pattern.node_rowptr[plan.offset + plan.n_owned] - K.pattern.node_rowptr[plan.offset])
    alg = 12 * own_nnz + 20 * 3 * plan.n_owned
    stats = torch.tensor([spmv_ms, alg / (spmv_ms / 1e3) / 1e9], dtype=torch.float64, device="cuda")
    dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    agg = torch.tensor([alg / (spmv_ms / 1e3) / 1e9], dtype=torch.float64, device="cuda")
    dist.all_reduce(agg)
    if rank == 0:
        hbm_peak, peak_src = measured_peaks()
        per_gpu = float(agg) / world
        line = {
            "metric": "hex8 beam solved DOF/s", "value": n_free / (ms_per_step / 1e3), "unit": "solved DOF/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"cubebeam hex8 cantilever {A}x{b}x{b}", "dof": int(nodes.size),
                       "free_dof": n_free, "elements": int(elements.shape[0]), "tol": tol,
                       "pcg_iterations": info.iterations, "rel_residual": info.rel_residual,
                       "preconditioner": "jacobi",
                       "parallelism": (f"{world} z-slabs, NVLink peer-memory exchange fused into the three PCG "
                                       "kernels of the solver's CUDA graph (no NCCL per iteration)"
                                       if SOLVER_USED["kind"] == "p2p" else
                                       f"{world} z-slabs, NCCL halo send/recv + all-reduced dots"),
                       "l2": "per-rank CSR slab larger than L2 for N <= 8; no flush"},
            "roofline": {"kernel": "spmv_tma_kernel<3,2> on the rank's slab (timed alone after the steps)", "bound": "hbm",
                         "achieved": per_gpu, "peak": hbm_peak, "unit": "GB/s", "frac": per_gpu / hbm_peak,
                         "traffic": None, "peak_source": peak_src, "aggregate_gb_per_s": float(agg),
                         "slowest_rank_ms": float(stats[0])},
            "stages_ms_rank0": stages,
            "cpu_baseline": None,
            "e2e": {"value": n_free / (ms_per_step / 1e3), "unit": "solved DOF/s",
                    "h2d_bytes_per_step": int(nodes[plan.g_lo:plan.g_hi].nbytes + plan.element_ids.size * 64
                                              + 2 * forces[plan.own_lo:plan.own_hi].nbytes),
                    "d2h_bytes_per_step": 256,
                    "note": "every step starts from the host mesh arrays: slab slicing, H2D, symbolic, assembly, "
                            "distributed PCG, reactions are all inside the timed region; results stay sharded on the GPUs"},
            "clocks": clocks, "gpu_launches": int(prof[0]) // max(args.steps, 1),
        }
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()
