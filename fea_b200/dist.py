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


def _scan_ranges(elements: np.ndarray, ranges):
    """For every node range [lo, hi): (element count, first id, last id, min node, max node) of the
    elements with at least one node inside.  One multi-threaded pass through fea_slab_scan for
    contiguous int64 / int32 connectivity; numpy otherwise."""
    elements = np.asarray(elements)
    if elements.flags.c_contiguous and elements.dtype in (np.int64, np.int32) and elements.ndim == 2:
        lib = _lib.load()
        rg = np.ascontiguousarray(np.asarray(ranges, dtype=np.int64).reshape(-1))
        out = np.empty(5 * len(ranges), dtype=np.int64)
        _lib.check(lib.fea_slab_scan(elements.ctypes.data, elements.dtype.itemsize, elements.shape[0],
                                     elements.shape[1], rg.ctypes.data, len(ranges), out.ctypes.data), "fea_slab_scan")
        return [tuple(int(v) for v in out[5 * i:5 * i + 5]) for i in range(len(ranges))]
    res = []
    for lo, hi in ranges:
        ids = np.nonzero(((elements >= lo) & (elements < hi)).any(axis=1))[0]
        if ids.size == 0:
            res.append((0, -1, -1, 0, 0))
        else:
            sub = elements[ids]
            res.append((int(ids.size), int(ids[0]), int(ids[-1]), int(sub.min()), int(sub.max())))
    return res


def plan_slab(elements: np.ndarray, cuts: np.ndarray, rank: int) -> SlabPlan:
    """Everything rank `rank` needs to know, computed locally and identically on every rank."""
    elements = np.asarray(elements)
    world = len(cuts) - 1
    own_lo, own_hi = int(cuts[rank]), int(cuts[rank + 1])
    ranges = [(own_lo, own_hi)]
    if rank > 0:
        ranges.append((int(cuts[rank - 1]), own_lo))
    if rank < world - 1:
        ranges.append((own_hi, int(cuts[rank + 2])))
    stats = _scan_ranges(elements, ranges)
    count, first, last, mn, mx = stats[0]
    if count == 0:
        ids, g_lo, g_hi = np.empty(0, dtype=np.int64), own_lo, own_hi
    else:
        g_lo, g_hi = min(mn, own_lo), max(mx + 1, own_hi)
        # Start the local range a few nodes early so that the owned rows begin on a 16-node (= 128-byte for
        # 3 doubles per node, or any smaller block) boundary of the local vectors: together with cuts at
        # multiples of 16 (default_cuts) no cache line then holds both owned and halo rows, and the SpMV's
        # face tiles may gather halo rows through L1 (fea_pcg_solve_p2p).  The extra nodes belong to the
        # lower neighbour, carry no local element and are never referenced.
        pad = (-(own_lo - g_lo)) % SLAB_ALIGN_NODES
        if rank > 0 and g_lo < own_lo and g_lo - pad >= int(cuts[rank - 1]):
            g_lo -= pad
        if last - first + 1 == count:  # layer-major meshes: a contiguous element range
            ids = np.arange(first, last + 1, dtype=np.int64)
        else:
            ids = np.nonzero(((elements >= own_lo) & (elements < own_hi)).any(axis=1))[0]
    recv_down = recv_up = send_down = send_up = None
    if count and mn < own_lo:
        if rank == 0 or g_lo < cuts[rank - 1]:
            raise ValueError("slab thinner than the mesh bandwidth: halo spans more than one neighbour")
        recv_down = (rank - 1, int(mn), own_lo)
    if g_hi > own_hi:
        if rank == world - 1 or g_hi > cuts[rank + 2]:
            raise ValueError("slab thinner than the mesh bandwidth: halo spans more than one neighbour")
        recv_up = (rank + 1, own_hi, g_hi)
    # what the neighbours need from me = their halo ranges
    k = 1
    if rank > 0:
        c, _, _, _, nb_max = stats[k]
        k += 1
        if c and nb_max + 1 > own_lo:
            send_down = (rank - 1, own_lo, nb_max + 1)
    if rank < world - 1:
        c, _, _, nb_min, _ = stats[k]
        if c and nb_min < own_hi:
            send_up = (rank + 1, nb_min, own_hi)
    return SlabPlan(rank, world, own_lo, own_hi, g_lo, g_hi, ids, send_down, send_up, recv_down, recv_up)


# ------------------------------------------------------------------------------------------------
# halo exchange + distributed PCG driver (backend-agnostic)
# ------------------------------------------------------------------------------------------------
class _DeviceMemory:
    """Raw device memory as a __cuda_array_interface__ provider (zero-copy torch view)."""

    def __init__(self, ptr: int, n_doubles: int):
        self.__cuda_array_interface__ = {"shape": (n_doubles,), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 2}


def _device_view(ptr: int, n_doubles: int, device) -> torch.Tensor:
    return torch.as_tensor(_DeviceMemory(ptr, n_doubles), device=device)


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
    history: np.ndarray | None = None
    seconds: float = 0.0  # wall time of the solver proper (synchronised), where the driver measures it


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
    exchange happen once, later solves only bump the epoch.  A different plan evicts (and frees) the
    cached blocks; `close_all()` runs at interpreter exit."""

    _cache: dict = {}
    _warned = False
    MAX_CACHED = 2

    @classmethod
    def get(cls, plan: SlabPlan, d: int, group=None) -> "P2PComm":
        key = (plan.rank, plan.world, plan.own_lo, plan.own_hi, plan.g_lo, plan.g_hi, d)
        if key not in cls._cache:
            # eviction is collective by construction: every rank builds the same sequence of plans
            while len(cls._cache) >= cls.MAX_CACHED:
                cls._cache.pop(next(iter(cls._cache))).close(group)
            cls._cache[key] = cls(plan, d, group)
        return cls._cache[key]

    @classmethod
    def close_all(cls) -> None:
        for comm in list(cls._cache.values()):
            comm.close(None, collective=False)
        cls._cache.clear()

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

    def close(self, group=None, collective: bool = True) -> None:
        """Unmap the peers' blocks and free the own one.  Collective (barrier first: nobody may still be
        writing into a block that is about to disappear) unless called at interpreter exit."""
        if collective and dist.is_available() and dist.is_initialized():
            torch.cuda.synchronize()
            dist.barrier(group=group)
        for r, p in enumerate(self.ptrs):
            if p is not None and r != self.plan.rank:
                self.lib.fea_comm_ipc_close(p)
        self.ptrs = []
        if self.own is not None:
            self.lib.fea_comm_free(self.own)
            self.own = None
        self.available = False

    def descriptor(self, boundary=(0, 0), algo: int = -1, max_rank_dof: int = 0) -> "_lib.PeerComm":
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
        c.boundary_lower_nodes, c.boundary_upper_nodes = int(boundary[0]), int(boundary[1])
        c.algo, c.max_rank_dof = int(algo), int(max_rank_dof)
        return c


import atexit  # noqa: E402

atexit.register(P2PComm.close_all)


def boundary_nodes(node_rowptr: torch.Tensor, node_colidx: torch.Tensor, offset: int, n_owned: int):
    """(lower, upper): owned nodes [0, lower) couple to a lower-halo node (local id < offset), the
    last `upper` owned nodes to an upper-halo node (local id >= offset + n_owned).  Column lists are
    sorted, so the first / last entry of a node's list decides.  The SpMV sweeps the rows in between
    first and waits for the neighbours' halo only then (fea_peer_comm, include/fea_b200.h)."""
    rp = node_rowptr[offset:offset + n_owned + 1].long()
    first = node_colidx[rp[:-1]]
    last = node_colidx[rp[1:] - 1]
    lo_idx = torch.nonzero(first < offset).flatten()
    up_idx = torch.nonzero(last >= offset + n_owned).flatten()
    lower = int(lo_idx.max()) + 1 if lo_idx.numel() else 0
    upper = n_owned - int(up_idx.min()) if up_idx.numel() else 0
    return lower, upper


def p2p_pcg(K, plan: SlabPlan, b_owned: torch.Tensor, dinv_owned: torch.Tensor, tol: float, max_iter: int, group=None,
            history: bool = False, max_rank_dof: int = 0, algo: int = -1):
    """Distributed Jacobi-PCG through fea_pcg_solve_p2p (NVLink peer memory, no NCCL per iteration).
    Returns (x_owned, DistInfo, history or None)."""
    import ctypes

    from . import core

    lib = _lib.load()
    d = K.dof_per_node
    comm = P2PComm.get(plan, d, group)
    pt = K.pattern
    desc = comm.descriptor(boundary_nodes(pt.node_rowptr, pt.node_colidx, plan.offset, plan.n_owned), algo=algo,
                           max_rank_dof=max_rank_dof)
    n = plan.n_owned * d
    x = torch.empty(n, dtype=torch.float64, device=b_owned.device)
    ws_bytes = lib.fea_pcg_workspace(n)
    work = torch.empty(ws_bytes, dtype=torch.uint8, device=b_owned.device)
    hist = torch.zeros(max_iter, dtype=torch.float64, device=b_owned.device) if history else None
    rowptr_owned = pt.node_rowptr[plan.offset:]
    res = _lib.PcgResult()
    # The first exchange of the solve is the only point where the ranks meet with host-side skew (each
    # one sliced, copied and assembled its slab on its own); meet on the host first, so that the
    # in-kernel spins only ever see device-side skew.
    torch.cuda.synchronize()
    dist.barrier(group=group)
    _lib.check(lib.fea_pcg_solve_p2p(plan.n_owned, d, rowptr_owned.data_ptr(), pt.node_colidx.data_ptr(),
                                     K.values.data_ptr(), pt.max_coupled, dinv_owned.data_ptr(), b_owned.data_ptr(),
                                     x.data_ptr(), float(tol), int(max_iter), work.data_ptr(), ws_bytes,
                                     None if hist is None else hist.data_ptr(),
                                     ctypes.byref(desc), ctypes.byref(res), core._stream()), "fea_pcg_solve_p2p")
    return (x, DistInfo(res.iterations, res.rel_residual, res.status, res.bnorm),
            hist[:res.iterations].cpu().numpy() if history else None)


# ------------------------------------------------------------------------------------------------
# slab inputs on the device, slab solve, gather
# ------------------------------------------------------------------------------------------------
@dataclass
class SlabInputs:
    """This rank's slab of the global mesh, resident on its GPU."""
    plan: SlabPlan
    nodes: torch.Tensor        # (n_local, 3) f64, local node range [g_lo, g_hi)
    elements: torch.Tensor     # (m, npe) int32, local node ids, ascending global element order
    fixed: torch.Tensor        # (3 n_local,) uint8
    loads_owned: torch.Tensor  # (3 n_owned,) f64
    n_nodes_global: int
    h2d_bytes: int


def upload_slab(nodes, elements, constraints, forces, plan: SlabPlan) -> SlabInputs:
    """Slice this rank's slab out of the GLOBAL host arrays and copy it to the device."""
    from . import core

    g_lo, g_hi = plan.g_lo, plan.g_hi
    nodes = np.asarray(nodes)
    elements = np.asarray(elements)
    nodes_h = np.ascontiguousarray(nodes[g_lo:g_hi], dtype=np.float64)
    nodes_d = core.to_device(nodes_h, torch.float64)
    ids = plan.element_ids
    if ids.size and int(ids[-1]) - int(ids[0]) + 1 == ids.size:
        # slabs of a layer-major mesh own a contiguous element range: a view, shifted on the device
        el_h = np.ascontiguousarray(elements[int(ids[0]):int(ids[-1]) + 1])
        elements_d = (core.to_device(el_h, torch.int64) - g_lo).to(torch.int32)
    else:
        el_h = np.ascontiguousarray(elements[ids] - g_lo)
        elements_d = core.to_device(el_h, torch.int32)
    cons_h = np.ascontiguousarray(np.asarray(constraints)[g_lo:g_hi])
    fixed = core._fixed_mask(cons_h, 3 * plan.n_local)
    loads_h = np.ascontiguousarray(np.asarray(forces)[plan.own_lo:plan.own_hi], dtype=np.float64)
    loads = core.to_device(loads_h, torch.float64).reshape(-1)
    h2d = int(nodes_h.nbytes + el_h.nbytes + cons_h.size + loads_h.nbytes)  # constraints travel as one byte per DOF
    return SlabInputs(plan, nodes_d, elements_d, fixed, loads, int(nodes.shape[0]), h2d)


def solve_slab(inp: SlabInputs, E: float, nu: float, tol: float = 1e-12, max_iter: int | None = None, group=None,
               history: bool = False, raise_on_failure: bool = True, max_rank_dof: int = 0):
    """Symbolic pass + assembly + distributed Jacobi-PCG + reactions on device-resident slab inputs.
    Returns (u_owned (n_owned, 3), reactions_owned (n_owned, 3), DistInfo, BlockCSR); info.history is
    the world residual history when `history`."""
    from . import core

    plan = inp.plan
    prof = STAGE_PROFILE if STAGE_PROFILE.get("enabled") else None

    def mark(name, t0):
        if prof is None:
            return t0
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        prof[name] = prof.get(name, 0.0) + (t1 - t0) * 1e3
        return t1

    t = time.perf_counter()
    K = core.assemble_hex8(inp.nodes, inp.elements, E, nu, fixed=inp.fixed)
    lo, hi = 3 * plan.offset, 3 * (plan.offset + plan.n_owned)
    dinv_owned = K.dinv[lo:hi].contiguous()
    t = mark("symbolic_assembly", t)
    ops = GpuOps(K, plan)
    if max_iter is None:
        max_iter = 10 * 3 * inp.n_nodes_global
    max_iter = int(min(max_iter, 2**31 - 1))
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
    hist = None
    if use_p2p:
        x, info, hist = p2p_pcg(K, plan, inp.loads_owned, dinv_owned, tol, max_iter, group=group, history=history,
                                max_rank_dof=max_rank_dof)
    else:
        x, info = distributed_pcg(ops, plan, 3, inp.loads_owned, dinv_owned, tol=tol, max_iter=max_iter, group=group)
    info.history = hist
    t = mark("pcg", t)
    if raise_on_failure and info.status != _lib.FEA_OK:
        # every rank holds the same status (the convergence decisions are bitwise identical): all raise
        _lib.raise_for_status(np.array([info.status, 0x7FFFFFFF - info.iterations]))
    # reactions: K_full u on the owned rows needs u on the halo
    u_ext = torch.zeros(3 * plan.n_local, dtype=torch.float64, device=x.device)
    u_ext[lo:hi] = x
    HaloExchange(plan, 3, group)(u_ext)
    reactions = torch.empty_like(x)
    ops.matvec_owned(u_ext, reactions)
    mark("reactions", t)
    return x.reshape(-1, 3), reactions.reshape(-1, 3), info, K


def solve_hex8_slab(nodes, elements, constraints, forces, E, nu, plan: SlabPlan, tol=1e-12, max_iter=None,
                    group=None, **kw):
    """This rank's share of solve(nodes, elements, constraints, forces) (cubebeam.py:79-108) from the
    GLOBAL host arrays: slice + H2D (upload_slab), then solve_slab.  Results stay on the device."""
    t = time.perf_counter()
    inp = upload_slab(nodes, elements, constraints, forces, plan)
    if STAGE_PROFILE.get("enabled"):
        torch.cuda.synchronize()
        STAGE_PROFILE["slice_h2d"] = STAGE_PROFILE.get("slice_h2d", 0.0) + (time.perf_counter() - t) * 1e3
    return solve_slab(inp, E, nu, tol=tol, max_iter=max_iter, group=group, **kw)


def gather_rows(plan: SlabPlan, cuts, owned: torch.Tensor, group=None, dst: int = 0):
    """`owned` is (k, n_owned, ...): concatenate every rank's owned nodes along axis 1 on rank `dst`
    (None elsewhere).  Slab sizes are uneven, hence grouped send / recv instead of a gather."""
    world, rank = plan.world, plan.rank
    if world == 1:
        return owned
    if rank == dst:
        n_total = int(cuts[-1])
        full = torch.empty((owned.shape[0], n_total) + tuple(owned.shape[2:]), dtype=owned.dtype, device=owned.device)
        full[:, plan.own_lo:plan.own_hi] = owned
        # one contiguous staging buffer per peer (a column slice of `full` is strided)
        bufs, ops = {}, []
        for r in range(world):
            if r == dst:
                continue
            lo, hi = int(cuts[r]), int(cuts[r + 1])
            bufs[r] = torch.empty((owned.shape[0], hi - lo) + tuple(owned.shape[2:]), dtype=owned.dtype,
                                  device=owned.device)
            ops.append(dist.P2POp(dist.irecv, bufs[r], r, group=group))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        for r, buf in bufs.items():
            full[:, int(cuts[r]):int(cuts[r + 1])] = buf
        return full
    for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, owned.contiguous(), dst, group=group)]):
        req.wait()
    return None


SLAB_ALIGN_NODES = 16  # 16 nodes x 3 doubles = 3 x 128 bytes: slab borders fall on cache-line boundaries


# ------------------------------------------------------------------------------------------------
# multi-RHS (BASELINE config 5) on slabs: the batched solver's step kernels + NCCL in between
# ------------------------------------------------------------------------------------------------
def distributed_pcg_multi(K, plan: SlabPlan, B_owned: torch.Tensor, dinv_owned: torch.Tensor, tol: float = 1e-12,
                          max_iter: int = 100000, chunk: int = 16, group=None):
    """Batched Jacobi-PCG (one recurrence per column, as fea_pcg_solve_multi) over all ranks.
    B_owned (n_owned_dof, R).  Per iteration: halo exchange of the (rows, R) search directions with the two
    slab neighbours (grouped NCCL send / recv), SpMM on the owned rows with the per-column p.Ap fused,
    all-reduce of R sums, update, all-reduce of 2R sums, direction.  Returns (X_owned, DistInfo) with
    info.history = iterations per column."""
    import ctypes

    from . import core

    lib = _lib.load()
    d = K.dof_per_node
    n_own, R = plan.n_owned * d, int(B_owned.shape[1])
    dev = B_owned.device
    B_owned = B_owned.contiguous()
    X = torch.empty_like(B_owned)
    multi = plan.world > 1
    # Halo exchange: NVLink peer memory when it can be set up (the search directions then live in this rank's
    # IPC-shared communication block and the neighbours store their face rows straight into it:
    # fea_peer_push / fea_peer_wait), else grouped NCCL / gloo send / recv.  FEA_DIST_COMM=nccl forces the latter.
    peer = None
    if (multi and dev.type == "cuda" and dist.get_backend(group) == "nccl"
            and os.environ.get("FEA_DIST_COMM", "p2p") != "nccl"):
        comm = P2PComm.get(plan, d * R, group)
        if comm.available:
            peer = comm
    if peer is not None:
        header = int(lib.fea_comm_bytes(0))
        P_ext = _device_view(peer.own + header, plan.n_local * d * R, dev).view(plan.n_local * d, R)
        P_ext.zero_()
    else:
        P_ext = torch.zeros((plan.n_local * d, R), dtype=torch.float64, device=dev)
    P_own = P_ext[plan.offset * d:plan.offset * d + n_own]
    SOLVER_USED["multi_halo"] = "p2p" if peer is not None else ("nccl" if multi else "none")
    ws_bytes = lib.fea_pcg_multi_workspace(n_own, R)
    work = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    off = (ctypes.c_int64 * 5)()
    _lib.check(lib.fea_pcg_multi_layout(R, ctypes.addressof(off)), "fea_pcg_multi_layout")
    state = work[off[0]:off[0] + 64].view(torch.int32)
    scal = work[off[1]:off[1] + 8 * 5 * R].view(torch.float64)    # world sums (what the kernels read)
    local = work[off[4]:off[4] + 8 * 5 * R].view(torch.float64)   # this rank's sums (what they write)

    def world_sum(lo: int, hi: int) -> None:
        scal[lo:hi].copy_(local[lo:hi])
        if multi:
            dist.all_reduce(scal[lo:hi], group=group)
    iters = work[off[3]:off[3] + 4 * R].view(torch.int32)
    pt = K.pattern
    rowptr_owned = pt.node_rowptr[plan.offset:]
    s = core._stream
    if peer is not None:
        peer.epoch += 1
        epoch, row_bytes = peer.epoch, 8 * d * R
        lower_block = upper_block = src_lo = dst_lo = src_hi = dst_hi = None
        cnt_lo = cnt_hi = 0
        if plan.send_down is not None:
            r, lo, hi = plan.send_down
            lower_block, cnt_lo = peer.ptrs[r], (hi - lo) * d * R
            src_lo = P_ext.data_ptr() + (lo - plan.g_lo) * row_bytes
            dst_lo = peer.ptrs[r] + header + (lo - peer.g_los[r]) * row_bytes
        if plan.send_up is not None:
            r, lo, hi = plan.send_up
            upper_block, cnt_hi = peer.ptrs[r], (hi - lo) * d * R
            src_hi = P_ext.data_ptr() + (lo - plan.g_lo) * row_bytes
            dst_hi = peer.ptrs[r] + header + (lo - peer.g_los[r]) * row_bytes
        has_lo, has_hi = int(plan.recv_down is not None), int(plan.recv_up is not None)

        def halo(_vec) -> None:
            _lib.check(lib.fea_peer_push(peer.own, lower_block, src_lo, dst_lo, cnt_lo, upper_block, src_hi, dst_hi,
                                         cnt_hi, state.data_ptr(), epoch, s()), "fea_peer_push")
            _lib.check(lib.fea_peer_wait(peer.own, has_lo, has_hi, state.data_ptr(), epoch, s()), "fea_peer_wait")
    else:
        halo = HaloExchange(plan, d, group)
    _lib.check(lib.fea_pcg_multi_init(n_own, R, B_owned.data_ptr(), dinv_owned.data_ptr(), X.data_ptr(),
                                      P_own.data_ptr(), float(tol), int(max_iter), work.data_ptr(), ws_bytes, s()),
               "fea_pcg_multi_init")
    world_sum(0, 2 * R)
    _lib.check(lib.fea_pcg_multi_activate(n_own, R, work.data_ptr(), s()), "fea_pcg_multi_activate")
    def iteration() -> None:
        halo(P_ext)
        _lib.check(lib.fea_pcg_multi_step_spmm(plan.n_owned, d, rowptr_owned.data_ptr(), pt.node_colidx.data_ptr(),
                                               K.values.data_ptr(), P_ext.data_ptr(), plan.offset, R,
                                               work.data_ptr(), s()), "fea_pcg_multi_step_spmm")
        world_sum(4 * R, 5 * R)
        _lib.check(lib.fea_pcg_multi_step_update(n_own, R, dinv_owned.data_ptr(), P_own.data_ptr(), X.data_ptr(),
                                                 work.data_ptr(), s()), "fea_pcg_multi_step_update")
        world_sum(2 * R, 4 * R)
        _lib.check(lib.fea_pcg_multi_step_direction(n_own, R, dinv_owned.data_ptr(), P_own.data_ptr(),
                                                    work.data_ptr(), s()), "fea_pcg_multi_step_direction")

    # One CUDA graph per chunk of iterations (kernels, the NCCL halo send/recv and the per-column all-reduces
    # captured together): the eager loop issues ~20 host operations per iteration and was host-bound beyond 2 ranks.
    # Every step kernel starts with `if (done) return` and the exchanges are idempotent after convergence, so
    # replaying a whole chunk past the last iteration is harmless.  The first chunk runs eagerly (NCCL sets up its
    # point-to-point channels on first use, which cannot happen inside a capture).
    use_graph = (multi and dev.type == "cuda" and dist.get_backend(group) == "nccl"
                 and os.environ.get("FEA_MULTI_GRAPH", "1") != "0")
    graph = None
    done_iter, finished = 0, False
    while not finished:
        if use_graph and done_iter > 0:
            if graph is None:
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                    for _ in range(chunk):
                        iteration()
            graph.replay()
            done_iter += chunk
        else:
            todo = min(chunk, max_iter - done_iter)
            for _ in range(todo):
                iteration()
            done_iter += todo
        st = state.cpu()  # one synchronisation per chunk; the decisions are identical on every rank
        finished = bool(int(st[1]) != 0) or done_iter >= max_iter
    if peer is not None:
        err = ctypes.c_int32(0)
        _lib.check(lib.fea_comm_error(peer.own, ctypes.byref(err), s()), "fea_comm_error")
        if err.value != 0:
            _lib.raise_for_status(np.array([_lib.FEA_ERR_PEER, 0x7FFFFFFF]))
    st = state.cpu()
    sc = scal.cpu().numpy()
    bn2, rr = sc[R:2 * R], sc[3 * R:4 * R]
    worst = float(np.sqrt((rr[bn2 > 0] / bn2[bn2 > 0]).max())) if (bn2 > 0).any() else 0.0
    status = int(st[2])
    if int(st[1]) == 0 and status == _lib.FEA_OK:
        status = _lib.FEA_ERR_MAXITER
    return X, DistInfo(int(st[0]), worst, status, float(np.sqrt(bn2.max())), iters.cpu().numpy().astype(np.int64))


def solve_truss_multi(nodes, members, k, constraints, loads, tol: float = 1e-12, max_iter: int | None = None,
                      group=None, cuts=None, gather: bool = True):
    """Linear truss K U = F for R load cases (BASELINE config 5) on every GPU of the process group:
    COLLECTIVE, every rank passes the same global host arrays (loads (3N, R)).  The lattice generator
    numbers nodes z-major, so a contiguous node range is a slab.  Returns
    (U (3N, R) host array on rank 0 / None elsewhere [device shards if not `gather`], DistInfo)."""
    from . import core

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    nodes = np.asarray(nodes, dtype=np.float64)
    members = np.asarray(members).reshape(-1, 2)
    n_nodes = nodes.shape[0]
    if cuts is None:
        cuts = default_cuts(n_nodes, world)
    plan = plan_slab(members, cuts, rank)
    g_lo, g_hi = plan.g_lo, plan.g_hi
    ids = plan.element_ids
    nodes_d = core.to_device(np.ascontiguousarray(nodes[g_lo:g_hi]), torch.float64)
    members_d = core.to_device(np.ascontiguousarray(members[ids] - g_lo), torch.int32)
    k_d = core.to_device(np.ascontiguousarray(np.asarray(k, dtype=np.float64)[ids]), torch.float64)
    fixed = core._fixed_mask(np.ascontiguousarray(np.asarray(constraints)[g_lo:g_hi]), 3 * plan.n_local)
    K = core.assemble_truss(nodes_d, members_d, k_d, fixed=fixed)
    lo, hi = 3 * plan.offset, 3 * (plan.offset + plan.n_owned)
    dinv_owned = K.dinv[lo:hi].contiguous()
    loads = np.asarray(loads, dtype=np.float64)
    B_owned = core.to_device(np.ascontiguousarray(loads[3 * plan.own_lo:3 * plan.own_hi]), torch.float64)
    if max_iter is None:
        max_iter = min(10 * 3 * n_nodes, 2**31 - 1)
    if nodes_d.is_cuda:
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    X, info = distributed_pcg_multi(K, plan, B_owned, dinv_owned, tol=tol, max_iter=max_iter, group=group)
    if nodes_d.is_cuda:
        torch.cuda.synchronize()
    info.seconds = time.perf_counter() - t0
    if info.status != _lib.FEA_OK:
        _lib.raise_for_status(np.array([info.status, 0x7FFFFFFF - info.iterations]))
    if not gather:
        return X, info
    R = X.shape[1]
    full = gather_rows(plan, cuts, X.reshape(1, plan.n_owned, 3 * R), group=group)
    out = None
    if full is not None:
        out = full.reshape(n_nodes * 3, R).cpu().numpy()
    return out, info


def default_cuts(n_nodes: int, world: int) -> np.ndarray:
    """Node-balanced slab cuts (not aligned to mesh layers: 401 layers on 8 ranks would leave one rank
    with 51 layers against 50, and the slowest rank sets the pace of every iteration), rounded to
    multiples of 16 nodes so that no cache line of a local vector holds rows of two ranks (plan_slab)."""
    cuts = node_cuts(n_nodes, world)
    if n_nodes >= 64 * SLAB_ALIGN_NODES * world:
        cuts[1:-1] = cuts[1:-1] // SLAB_ALIGN_NODES * SLAB_ALIGN_NODES
    return cuts


def solve_hex8(nodes, elements, constraints, forces, E: float, nu: float, tol: float = 1e-12,
               max_iter: int | None = None, group=None, cuts=None, plan: SlabPlan | None = None,
               return_info: bool = False, all_ranks: bool = False):
    """The reference's solve(nodes, elements, constraints, forces) -> (displacements, forces)
    (cubebeam.py:79-108) on every GPU of the process group: a COLLECTIVE call, every rank passes the
    same global host arrays.  Each rank copies only its slab of node layers to its GPU, assembles it
    and takes part in one distributed Jacobi-PCG; the owned rows of u and of K_full u are gathered on
    rank 0 over NVLink and copied to the host there.  Rank 0 returns the host arrays (N, 3) like the
    reference; the other ranks return (None, None) unless `all_ranks` (then every rank receives a copy).
    Raises like the single-GPU path, on every rank (ValueError for an inverted element,
    numpy.linalg.LinAlgError for a singular / non-converging reduced system)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    nodes_a = np.asarray(nodes)
    n_nodes = int(nodes_a.shape[0])
    if cuts is None:
        cuts = default_cuts(n_nodes, world)
    if plan is None:
        plan = plan_slab(np.asarray(elements), cuts, rank)
    inp = upload_slab(nodes_a, elements, constraints, forces, plan)
    max_rank_dof = 3 * int(np.diff(cuts).max())
    u, react, info, K = solve_slab(inp, E, nu, tol=tol, max_iter=max_iter, group=group, max_rank_dof=max_rank_dof)
    t = time.perf_counter()
    both = torch.stack([u.reshape(-1), react.reshape(-1)]).reshape(2, plan.n_owned, 3)
    full = gather_rows(plan, cuts, both, group=group)
    if all_ranks:
        if full is None:
            full = torch.empty((2, n_nodes, 3), dtype=torch.float64, device=u.device)
        dist.broadcast(full, src=0, group=group)
    out = (None, None)
    if full is not None:
        host = full.cpu().numpy()
        out = (host[0].reshape(nodes_a.shape), host[1].reshape(nodes_a.shape))
    if STAGE_PROFILE.get("enabled"):
        torch.cuda.synchronize()
        STAGE_PROFILE["gather_d2h"] = STAGE_PROFILE.get("gather_d2h", 0.0) + (time.perf_counter() - t) * 1e3
    if return_info:
        return out[0], out[1], info, K
    return out


def solve_hex8_or_none(nodes, elements, constraints, forces, E, nu, group=None, **kw):
    """solve_hex8, or None on EVERY rank when the mesh cannot be cut into `world` slabs (a slab thinner
    than the mesh bandwidth, or a node numbering that is not layer-major): the caller then solves on one
    GPU per rank.  The decision is collective (all-reduced), so no rank is left waiting."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n_nodes = int(np.asarray(nodes).shape[0])
    cuts = default_cuts(n_nodes, world)
    ok, plan = 1, None
    try:
        if int(np.diff(cuts).min()) < 1:
            raise ValueError("fewer nodes than ranks")
        plan = plan_slab(np.asarray(elements), cuts, rank)
    except ValueError:
        ok = 0
    flag = torch.tensor([ok], dtype=torch.int32, device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    if int(flag) == 0:
        return None
    return solve_hex8(nodes, elements, constraints, forces, E, nu, group=group, cuts=cuts, plan=plan, **kw)


def active_world(group=None) -> int:
    """World size of the initialised NCCL process group (1 if there is none): model.solve_hex8 routes
    through solve_hex8 above when it is > 1."""
    if not (dist.is_available() and dist.is_initialized()):
        return 1
    try:
        if dist.get_backend(group) != "nccl":
            return 1
    except Exception:  # noqa: BLE001
        return 1
    return dist.get_world_size(group)


# which distributed solver the last solve_slab call used ("p2p" | "nccl" | "single")
SOLVER_USED: dict = {"kind": None}

# stage timings of the slab solve (ms, accumulated; each stage ends with a device synchronise) when
# STAGE_PROFILE["enabled"] is set -- bench.py switches it on for ONE extra untimed step.
STAGE_PROFILE: dict = {"enabled": False}
