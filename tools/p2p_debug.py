"""torchrun --nproc-per-node N tools/p2p_debug.py [A b [max_iter]]: one capped multi-GPU solve with
FEA_P2P_DEBUG=1 (per-iteration %globaltimer stamps of SpMV start / face wait / halo push, printed by
fea_pcg_solve_p2p on stderr) and the per-iteration time."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fea_b200 import cubebeam
from fea_b200 import dist as fdist
A, b = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (400, 80)
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 400
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
nodes, elements, cons, forces = cubebeam.cantilever_case(A, b)
cuts = fdist.default_cuts(nodes.shape[0], world)
plan = fdist.plan_slab(elements, cuts, rank)
inp = fdist.upload_slab(nodes, elements, cons, forces, plan)
E, NU = 10_000_000 * 6894.76, 0.3
for rep in range(2):
    if rep == 1:
        os.environ["FEA_P2P_DEBUG"] = "1"
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    u, f, info, K = fdist.solve_slab(inp, E, NU, max_iter=iters, raise_on_failure=False, max_rank_dof=3 * int(np.diff(cuts).max()))
    torch.cuda.synchronize(); t1 = time.perf_counter()
    if rank == 0:
        print(f"rep {rep}: {info.iterations} iterations, {(t1 - t0) / max(info.iterations, 1) * 1e6:.1f} us/iteration (incl. assembly)", flush=True)
# the plain (ungated) SpMV kernel on this rank's slab, both ranks busy at the same time
os.environ.pop("FEA_P2P_DEBUG", None)
ops = fdist.GpuOps(K, plan)
x_ext = torch.randn(3 * plan.n_local, dtype=torch.float64, device="cuda")
y = torch.empty(3 * plan.n_owned, dtype=torch.float64, device="cuda")
for _ in range(3):
    ops.matvec_owned(x_ext, y)
dist.barrier()
a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(50):
    ops.matvec_owned(x_ext, y)
c.record(); torch.cuda.synchronize()
print(f"rank {rank}: plain fea_spmv on the slab, torch-allocated x: {a.elapsed_time(c) / 50 * 1e3:.1f} us", flush=True)
# the same with x inside the (peer-mapped) communication block
comm = fdist.P2PComm.get(plan, 3)
import ctypes
from fea_b200 import _lib
lib = _lib.load()
pt = K.pattern
rowptr_owned = pt.node_rowptr[plan.offset:]
x_ptr = comm.own + 4096
lib.fea_spmv(plan.n_owned, 3, rowptr_owned.data_ptr(), pt.node_colidx.data_ptr(), K.values.data_ptr(), pt.max_coupled, x_ptr, y.data_ptr(), None)
torch.cuda.synchronize(); dist.barrier()
a.record()
for _ in range(50):
    lib.fea_spmv(plan.n_owned, 3, rowptr_owned.data_ptr(), pt.node_colidx.data_ptr(), K.values.data_ptr(), pt.max_coupled, x_ptr, y.data_ptr(), None)
c.record(); torch.cuda.synchronize()
print(f"rank {rank}: plain fea_spmv on the slab, x in the peer-mapped comm block: {a.elapsed_time(c) / 50 * 1e3:.1f} us", flush=True)
dist.barrier(); dist.destroy_process_group()
