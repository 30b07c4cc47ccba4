"""A/B of the halo-gated PCG SpMV against the plain one on ONE GPU: fea_pcg_solve_p2p at world = 1,
with and without FEA_P2P_FORCE_GATED=1, on the slab a rank owns at N GPUs (A/N x b x b).

    python tools/gated_probe.py [A b [iterations]]
"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fea_b200 import _lib, core, cubebeam, utils  # noqa: E402

A, b = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (400, 80)
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 400
E, NU = 10_000_000 * 6894.76, 0.3
nodes, elements, fixed, loads = cubebeam.cantilever_case_device(A, b)
K = core.assemble_hex8(nodes, elements, E, NU, fixed=fixed)
lib = _lib.load()
n = K.n_dof
own = ctypes.c_void_p()
assert lib.fea_comm_alloc(lib.fea_comm_bytes(n), ctypes.byref(own)) == 0
pt = K.pattern
x = torch.empty(n, dtype=torch.float64, device="cuda")
ws = lib.fea_pcg_workspace(n)
work = torch.empty(ws, dtype=torch.uint8, device="cuda")
res = _lib.PcgResult()
epoch = 0
only = os.environ.get("PROBE_ONLY")  # "algo,gated": a single variant (for ncu)
variants = [(int(only.split(",")[0]), only.split(",")[1])] if only else [(a, f) for a in (0, 1) for f in ("0", "1", "0", "1")]
for algo, forced in variants:
    if True:
        os.environ["FEA_P2P_FORCE_GATED"] = forced
        epoch += 1
        desc = _lib.PeerComm()
        desc.world, desc.rank, desc.lower_peer, desc.upper_peer, desc.epoch = 1, 0, -1, -1, epoch
        desc.comm[0] = own.value
        desc.algo = algo
        if os.environ.get("FEA_P2P_FAKE_TILES"):
            desc.boundary_lower_nodes = desc.boundary_upper_nodes = (b + 1) ** 2 + b + 2
        lib.fea_profile_enable(1)
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rc = lib.fea_pcg_solve_p2p(pt.n_nodes, 3, pt.node_rowptr.data_ptr(), pt.node_colidx.data_ptr(),
                                   K.values.data_ptr(), pt.max_coupled, K.dinv.data_ptr(), loads.data_ptr(), x.data_ptr(),
                                   1e-12, iters, work.data_ptr(), ws, None, ctypes.byref(desc), ctypes.byref(res), None)
        c.record()
        torch.cuda.synchronize()
        prof = (4 * ctypes.c_double)()
        lib.fea_profile_read(prof)
        print(f"{A}x{b}x{b} algo {algo} gated={forced}: rc {rc} status {res.status} iterations {res.iterations} "
              f"{a.elapsed_time(c) / max(res.iterations, 1) * 1e3:.1f} us/iteration, sampled SpMV "
              f"{prof[2] / max(prof[1], 1) * 1e3:.1f} us ({int(prof[1])} samples), |x| {float(x.norm()):.6e}")
