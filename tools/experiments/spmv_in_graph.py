"""Floor of the SpMV inside a CUDA graph at a given mesh size: N back-to-back fea_spmv launches captured
once and replayed (what the PCG's graph pays per SpMV, without the vector kernels between them).

    python tools/experiments/spmv_in_graph.py A b [launches]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from fea_b200 import core, cubebeam, utils  # noqa: E402

A, b = int(sys.argv[1]), int(sys.argv[2])
N = int(sys.argv[3]) if len(sys.argv) > 3 else 200
n2, q2 = cubebeam.generate_quad_grid(b, b, 0.1, 0.1)
nodes, elements = utils.stack_faces_2d_device(n2, q2, np.linspace(0, 1.0, A + 1))
fixed = (nodes[:, 2] == 0).repeat_interleave(3).to(torch.uint8)
K = core.assemble_hex8(nodes, elements, 10_000_000 * 6894.76, 0.3, fixed=fixed)
x = torch.randn(K.n_dof, dtype=torch.float64, device="cuda")
y = torch.empty_like(x)
for _ in range(5):
    K.matvec(x, out=y)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    K.matvec(x, out=y)
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=side):
        for _ in range(N):
            K.matvec(x, out=y)
best = 1e9
for _ in range(5):
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    g.replay()
    c.record()
    torch.cuda.synchronize()
    best = min(best, a.elapsed_time(c) / N * 1e3)
fmt = 8.0 * K.nnz + 4.0 * K.nnz / 9 + 4.0 * K.n_dof / 3 + 16.0 * K.n_dof
print({"mesh": [A, b, b], "dof": K.n_dof, "us_per_spmv_in_graph": round(best, 2),
       "format_GB_per_s": round(fmt / best / 1e3, 1), "FEA_TMA_L2": os.environ.get("FEA_TMA_L2", "default")})
