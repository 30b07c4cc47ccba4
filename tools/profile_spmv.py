"""Short driver for ncu: assemble the hex8 cantilever A x b x b on the device, then run a few plain
SpMV launches and a PCG capped at a few iterations (so that `ncu -k regex:spmv -c 3` sees the
bench's dominant kernel on the bench's own matrix without replaying 25k launches).

    python tools/profile_spmv.py [A b [pcg_iterations]]
"""
import sys, os
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fea_b200 import core, cubebeam, utils

A, b = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (400, 80)
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
E, NU = 10_000_000 * 6894.76, 0.3
n2, q2 = cubebeam.generate_quad_grid(b, b, 0.1, 0.1)
nodes, elements = utils.stack_faces_2d_device(n2, q2, np.linspace(0, 1.0, A + 1))
fixed = (nodes[:, 2] == 0).repeat_interleave(3).to(torch.uint8)
loads = torch.zeros(nodes.shape[0] * 3, dtype=torch.float64, device="cuda")
loads[1::3] = (nodes[:, 1] == 0).to(torch.float64)
K = core.assemble_hex8(nodes, elements, E, NU, fixed=fixed)
x = torch.randn(K.n_dof, dtype=torch.float64, device="cuda")
y = torch.empty_like(x)
for _ in range(5):
    K.matvec(x, out=y)
a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
a.record()
for _ in range(20):
    K.matvec(x, out=y)
c.record()
torch.cuda.synchronize()
ms = a.elapsed_time(c) / 20
alg = 12 * K.nnz + 20 * K.n_dof
from fea_b200 import _lib
import ctypes
lib = _lib.load()
lib.fea_profile_enable(1)
a.record()
u, info = core.pcg(K, loads, tol=1e-12, max_iter=iters, raise_on_failure=False)
c.record()
torch.cuda.synchronize()
prof = (4 * ctypes.c_double)()
lib.fea_profile_read(prof)
print("ok", K.n_dof, K.nnz, info.iterations, "spmv %.4f ms = %.0f GB/s algorithmic" % (ms, alg / ms / 1e6),
      "| pcg %.4f ms/iter, sampled pcg-spmv %.4f ms" % (a.elapsed_time(c) / max(info.iterations, 1),
                                                       prof[2] / max(prof[1], 1)))
