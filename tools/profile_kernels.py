"""Short driver that launches every kernel class of the path a few times, for ncu captures and
CUDA-event timings of the kernels bench.py does not time individually:

  * Ke microbench: ke_hex8_kernel on congruent and on randomly distorted hexes
    (corner jitter U(-0.2h, 0.2h), default_rng(2)) -> elem/s and FP64 TFLOP/s at 21 kflop/elem
  * assemble_hex8_kernel on an A x b x b cantilever (algorithmic bytes 8 nnz + 24 N + 32 M)
  * sustained plain SpMV (many back-to-back launches, timed in groups) vs the PCG's fused SpMV
  * multi-RHS kernels (SpMM, update, direction) on the lattice truss

    python tools/profile_kernels.py [--hex A b] [--ke M] [--lattice n] [--spmv-reps R] [--only ke,asm,spmv,multi]
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fea_b200 import core, cubebeam, truss, utils  # noqa: E402

E_HEX, NU_HEX = 10_000_000 * 6894.76, 0.3


def arg(name, default, n=1):
    if name in sys.argv:
        i = sys.argv.index(name)
        vals = sys.argv[i + 1:i + 1 + n]
        return [type(d)(v) for d, v in zip(default, vals)] if n > 1 else type(default)(vals[0])
    return default


ONCE = "--once" in sys.argv  # single launches, no warm-up: the pass that runs under `ncu --set full`


def timed(fn, reps=1):
    """(result, best ms over `reps` individually timed calls).  Boxes of the pool show 2-4x run-to-run
    noise on ~1 ms kernels (power management), so the minimum is the comparable figure."""
    if ONCE:
        reps = 1
    best = float("inf")
    for _ in range(reps):
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        out = fn()
        c.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(c))
    return out, best


def device_mesh(A, b):
    n2, q2 = cubebeam.generate_quad_grid(b, b, 0.1, 0.1)
    return utils.stack_faces_2d_device(n2, q2, np.linspace(0, 1.0, A + 1))


def ke_bench(M):
    """M hexes cut from a cube grid; the distorted variant jitters every corner independently."""
    side = int(round(M ** (1 / 3))) + 1
    n2, q2 = cubebeam.generate_quad_grid(side, side, 1.0, 1.0)
    nodes, elements = device_mesh_from(n2, q2, side)
    elements = elements[:M].contiguous()
    h = 1.0 / side
    out = {}
    for label, jitter in (("congruent", 0.0), ("distorted", 0.2)):
        nd = nodes.clone()
        if jitter:
            rng = np.random.default_rng(2)
            nd += torch.from_numpy(rng.uniform(-jitter * h, jitter * h, size=tuple(nd.shape))).to(nd.device)
        for _ in range(0 if ONCE else 2):
            utils.hexahedral_stiffness_matrices(nd, elements, E_HEX, NU_HEX)
        _, ms = timed(lambda: utils.hexahedral_stiffness_matrices(nd, elements, E_HEX, NU_HEX), 20)
        out[label] = {"ms": ms, "elem_per_s": M / ms * 1e3, "tflops_at_21kflop": 21e3 * M / ms / 1e9,
                      "write_gb_per_s": 4608 * M / ms / 1e6}
    print(json.dumps({"bench": "ke_hex8", "elements": M, **out}), flush=True)


def device_mesh_from(n2, q2, layers):
    return utils.stack_faces_2d_device(n2, q2, np.linspace(0, 1.0, layers + 1))


def asm_bench(A, b):
    nodes, elements = device_mesh(A, b)
    fixed = (nodes[:, 2] == 0).repeat_interleave(3).to(torch.uint8)
    for _ in range(0 if ONCE else 2):
        pat = core.symbolic(elements, nodes.shape[0])
        K = core.assemble_hex8(nodes, elements, E_HEX, NU_HEX, pattern=pat, fixed=fixed)
    pat, ms_sym = timed(lambda: core.symbolic(elements, nodes.shape[0]), 3)
    K, ms_num = timed(lambda: core.assemble_hex8(nodes, elements, E_HEX, NU_HEX, pattern=pat, fixed=fixed), 10)
    M, N = elements.shape[0], nodes.shape[0]
    alg = 8 * K.nnz + 24 * N + 32 * M
    print(json.dumps({"bench": "assemble_hex8", "mesh": [A, b, b], "elements": M, "nnz": K.nnz,
                      "ms": {"symbolic": ms_sym, "numeric": ms_num},
                      "numeric_elem_per_s": M / ms_num * 1e3, "numeric_algorithmic_gb_per_s": alg / ms_num / 1e6,
                      "numeric_tflops_at_43kflop_executed": 43e3 * M / ms_num / 1e9}), flush=True)
    return nodes, K


def spmv_bench(nodes, K, reps):
    import ctypes

    from fea_b200 import _lib

    x = torch.randn(K.n_dof, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    for _ in range(0 if ONCE else 5):
        K.matvec(x, out=y)
    groups = []
    for _ in range(max(1, reps // 200)):
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(1 if ONCE else 200):
            K.matvec(x, out=y)
        c.record()
        torch.cuda.synchronize()
        groups.append(round(a.elapsed_time(c) / (1 if ONCE else 200), 4))
    alg = 12 * K.nnz + 20 * K.n_dof
    lib = _lib.load()
    lib.fea_profile_enable(1)
    loads = torch.zeros(K.n_dof, dtype=torch.float64, device="cuda")
    loads[1::3] = (nodes[:, 1] == 0).to(torch.float64)
    (u, info), ms_pcg = timed(lambda: core.pcg(K, loads, tol=1e-12, max_iter=2 if ONCE else 600, raise_on_failure=False))
    prof = (4 * ctypes.c_double)()
    lib.fea_profile_read(prof)
    print(json.dumps({"bench": "spmv_sustained", "dof": K.n_dof, "nnz": K.nnz, "ms_per_spmv_by_group_of_200": groups,
                      "algorithmic_gb_per_s_last_group": alg / groups[-1] / 1e6,
                      "pcg_ms_per_iteration": ms_pcg / max(info.iterations, 1),
                      "pcg_spmv_sampled_ms": prof[2] / max(prof[1], 1)}), flush=True)


def multi_bench(n, n_rhs, iters):
    nodes, members, k, cons, loads = truss.lattice_truss(n, n_rhs)
    nd, mem = core.to_device(nodes, torch.float64), core.to_device(members, torch.int32)
    kd = core.to_device(k, torch.float64)
    fixed = core._fixed_mask(cons, nodes.size)
    B = core.to_device(loads, torch.float64)
    K = core.assemble_truss(nd, mem, kd, fixed=fixed)
    if not ONCE:
        K.matmat(B)
    Yout = torch.empty_like(B)
    spmm_rounds = [timed(lambda: K.matmat(B, out=Yout), 4)[1] for _ in range(1 if ONCE else 5)]
    ms_spmm = min(spmm_rounds)
    if not ONCE:
        core.pcg_multi(K, B, tol=1e-12, max_iter=16, raise_on_failure=False)
    # per-iteration time = slope between a short and a long capped solve (removes the fixed set-up cost)
    slopes = []
    for _ in range(1 if ONCE else 3):
        (_, info0), ms0 = timed(lambda: core.pcg_multi(K, B, tol=1e-12, max_iter=16, raise_on_failure=False))
        (X, info), ms1 = timed(lambda: core.pcg_multi(K, B, tol=1e-12, max_iter=16 + iters, raise_on_failure=False))
        slopes.append((ms1 - ms0) if not ONCE else ms1)
    ms = min(slopes)
    extra_iters = max(info.iterations - info0.iterations, 1) if not ONCE else max(info.iterations, 1)
    info.iterations = extra_iters
    n_dof = K.n_dof
    alg_spmm = 12 * K.nnz + 4 * n_dof + 16 * n_dof * n_rhs
    alg_iter = alg_spmm + 9 * 8 * n_dof * n_rhs
    print(json.dumps({"bench": "multi_rhs", "lattice": n, "dof": n_dof, "nnz": K.nnz, "n_rhs": n_rhs,
                      "variant": os.environ.get("FEA_SPMM_VARIANT", "0"),
                      "ms": {"spmm": ms_spmm, "pcg_iteration": ms / max(info.iterations, 1), "fixed_overhead": ms0,
                             "spmm_rounds": [round(v, 3) for v in spmm_rounds],
                             "pcg_iteration_rounds": [round(v / extra_iters, 3) for v in slopes]},
                      "iterations": info.iterations,
                      "spmm_algorithmic_gb_per_s": alg_spmm / ms_spmm / 1e6,
                      "iteration_algorithmic_gb_per_s": alg_iter / (ms / max(info.iterations, 1)) / 1e6}), flush=True)


if __name__ == "__main__":
    only = arg("--only", "ke,asm,spmv,multi").split(",")
    A, b = arg("--hex", [400, 80], 2)
    if "ke" in only:
        ke_bench(arg("--ke", 500_000))
    if "asm" in only or "spmv" in only:
        nodes, K = asm_bench(A, b)
        if "spmv" in only:
            spmv_bench(nodes, K, arg("--spmv-reps", 2000))
        del nodes, K
    if "multi" in only:
        multi_bench(arg("--lattice", 93), 64, 2 if ONCE else arg("--multi-iters", 96))
