"""torchrun --nproc-per-node N tools/check_dist.py [A b]: the N-rank slab solve against the
1-rank solve of the same mesh on rank 0 (values of owned K rows bit-exact, u within 1e-10)."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fea_b200 import core, cubebeam, model
from fea_b200 import dist as fdist

A, b = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (64, 16)
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
E, NU = 10_000_000 * 6894.76, 0.3
nodes, elements, cons, forces = cubebeam.cantilever_case(A, b)
cuts = fdist.node_cuts(nodes.shape[0], world, layer=(b + 1) ** 2)
plan = fdist.plan_slab(elements, cuts, rank)
u, react, info, K = fdist.solve_hex8_slab(nodes, elements, cons, forces, E, NU, plan)
# single-GPU reference on every rank (cheap at this size)
u1, f1, info1, K1 = model.solve_hex8(nodes, elements, cons, forces, E, NU, return_info=True)
lo, hi = plan.own_lo, plan.own_hi
rp1 = K1.pattern.node_rowptr.cpu().numpy()
rpl = K.pattern.node_rowptr.cpu().numpy()
v1 = K1.values[9 * rp1[lo]:9 * rp1[hi]]
vl = K.values[9 * rpl[plan.offset]:9 * rpl[plan.offset + plan.n_owned]]
same_vals = bool(torch.equal(v1, vl))
uerr = np.abs(u.cpu().numpy() - u1[lo:hi]).max() / np.abs(u1).max()
ferr = np.abs(react.cpu().numpy() - f1[lo:hi]).max() / np.abs(f1).max()
res = [None] * world
dist.all_gather_object(res, (rank, same_vals, float(uerr), float(ferr), info.iterations, info1.iterations,
                             info.rel_residual, info.status))
if rank == 0:
    for r in res:
        print("rank %d: owned K rows bit-identical=%s  |u-u1|=%.2e  |f-f1|=%.2e  iters %d (1 GPU: %d)  rel_res %.2e status %d" % r)
    ok = all(r[1] and r[2] < 1e-10 and r[3] < 1e-9 and r[7] == 0 for r in res)
    print("DIST CHECK", "PASS" if ok else "FAIL")
dist.barrier()
dist.destroy_process_group()
