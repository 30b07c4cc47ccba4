#!/usr/bin/env python
"""Benchmark of the hot path: hex8 cantilever assembled and solved to a 1e-12 recurrence residual.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload A,b]

One "step" = one full pass of the hot path over the workload: symbolic pass + fused Ke/assembly +
Jacobi-PCG to ||r|| <= 1e-12 ||b|| + reaction product.  Metric (BASELINE.json): solved DOF/s =
free DOF / step time.  Workload (all N): BASELINE config "cubebeam hex8 cantilever 400x80x80"
(7,892,883 DOF, 2,560,000 elements, structural nnz 627,797,529); N > 1 partitions the SAME
mesh into z-slabs (strong scaling) with a halo exchange per SpMV and all-reduced dot products.

Prints ONE JSON line (rank 0).  `value`: inputs resident in HBM, CUDA-event timed.  `e2e`: the
reference-facing call `cubebeam.solve(nodes, elements, constraints, forces)` on pinned HOST arrays,
host<->device copies inside the timed region.  `roofline`: the dominant kernel (PCG SpMV), timed
live with CUDA events inside the timed steps.  `cpu_baseline`: the oracle (numpy/scipy port of the
reference path) timed on this box's host cores on a bounded slab of the same mesh.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

TOL = 1e-12
E_HEX, NU_HEX = 10_000_000 * 6894.76, 0.3


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("FEA_BENCH_WORKLOAD", "400,80"),
                    help="A,b: A element layers along z, b x b elements in the section")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def committed_traffic(workload: str):
    """dram bytes per SpMV launch from the committed `ncu --set full` capture, if there is one."""
    try:
        with open(os.path.join(ROOT, "profiles", "spmv_traffic.json")) as fh:
            return json.load(fh).get(workload)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle on a bounded slab of the same mesh, extrapolated to the full workload
# ----------------------------------------------------------------------------------------------
def cpu_sample(A: int, b: int, iterations_full: int | None, slab_layers: int = 16, pcg_iters: int = 20):
    """CPU arm on `slab_layers` element layers of the b x b section, scaled to the workload
    (assembly linearly in elements, PCG linearly in nnz x iterations).

    Headline figure: oracle/fea_oracle_c.c -- OUR OWN C/OpenMP port of the reference algorithm
    (utils.py:127-239 + cubebeam.py:79-108 with the sparse Jacobi-PCG) on all host threads.  It is
    a port written for this bench with straightforward parallel loops (no cache blocking, no NUMA
    pinning), not the reference's code: the reference is single-threaded Python and cannot hold
    this size.  The single-core numpy/scipy oracle, which is closest to what the reference itself
    executes, is timed beside it (`numpy_1core_*`) so that both comparisons are on the line."""
    from oracle import c_oracle as co
    from oracle import fea_oracle as fo

    layers = min(slab_layers, A)
    nodes, elements, cons, forces = fo.cantilever_case(layers, b, beam_length=layers / A)
    n_elem_full = A * b * b
    n1, n3 = b + 1, A + 1
    nnz_full = 9 * (3 * n1 - 2) ** 2 * (3 * n3 - 2)
    free_full = 3 * n1 * n1 * A
    if iterations_full is None:
        iterations_full = int(21.3 * A)  # ~linear in the long dimension
    free = fo.free_dofs(cons)
    ff = forces.flatten()[free]

    # --- C / OpenMP port, all host threads
    cores = co.threads()
    t0 = time.perf_counter()
    pattern = co.dof_pattern(elements, nodes.shape[0], 3)
    t1 = time.perf_counter()
    K = co.assemble_hex8(nodes, elements, E_HEX, NU_HEX, pattern=pattern)
    t2 = time.perf_counter()
    Kff = K[free][:, free].tocsr()
    t3 = time.perf_counter()
    co.jacobi_pcg(Kff, ff, tol=0.0, maxiter=5)  # thread pool / page warm-up
    t4 = time.perf_counter()
    c_iters = 10 * pcg_iters
    co.jacobi_pcg(Kff, ff, tol=0.0, maxiter=c_iters)
    t5 = time.perf_counter()
    c_asm = ((t2 - t0) + (t3 - t2)) * n_elem_full / elements.shape[0]
    c_it = (t5 - t4) / c_iters * nnz_full / Kff.nnz
    c_total = c_asm + c_it * iterations_full

    # --- numpy / scipy oracle, one core (closest to the reference's own execution)
    s0 = time.perf_counter()
    Ke = fo.hex8_ke_batched(nodes, elements, E_HEX, NU_HEX)
    s1 = time.perf_counter()
    Kn = fo.assemble_csr(elements, Ke, nodes.shape[0], 3)
    Knff = Kn[free][:, free].tocsr()
    s2 = time.perf_counter()
    del Ke
    fo.jacobi_pcg(Knff, ff, tol=0.0, maxiter=pcg_iters)
    s3 = time.perf_counter()
    n_asm = (s2 - s0) * n_elem_full / elements.shape[0]
    n_it = (s3 - s2) / pcg_iters * nnz_full / Knff.nnz
    n_total = n_asm + n_it * iterations_full

    return {
        "value": free_full / c_total,
        "unit": "solved DOF/s",
        "cores": cores,
        "kind": "port",
        "sample": (f"OUR C/OpenMP port of the reference algorithm (oracle/fea_oracle_c.c; plain parallel loops, not "
                   f"cache-blocked or NUMA-pinned; the reference itself is single-threaded Python and cannot hold this "
                   f"size) on {cores} threads, {layers}x{b}x{b} slab of the workload ({elements.shape[0]} elements, "
                   f"nnz {Kff.nnz}): pattern {t1 - t0:.2f}s + Ke/assembly {t2 - t1:.2f}s + reduce {t3 - t2:.2f}s + "
                   f"{c_iters} Jacobi-PCG iterations {t5 - t4:.2f}s; extrapolated linearly to {n_elem_full} elements "
                   f"and {iterations_full} iterations x nnz {nnz_full} (= {c_total:.0f}s).  Beside it, the numpy/scipy "
                   f"oracle on 1 core (what the reference's own code path amounts to): {n_total:.0f}s"),
        "assembly_elem_per_s": elements.shape[0] / (t3 - t0),
        "spmv_gb_per_s": (12 * Kff.nnz + 20 * Kff.shape[0]) / ((t5 - t4) / c_iters) / 1e9,
        "numpy_1core_value": free_full / n_total,
        "numpy_1core_ke_elem_per_s": elements.shape[0] / (s1 - s0),
        "numpy_1core_assembly_elem_per_s": elements.shape[0] / (s2 - s0),
        "numpy_1core_spmv_gb_per_s": (12 * Knff.nnz + 20 * Knff.shape[0]) / ((s3 - s2) / pcg_iters) / 1e9,
        "seconds": (t5 - t0) + (s3 - s0),
    }


def run_reference(args, A, b):
    """--impl reference: the CPU arm.  The reference is pure Python and its dense solve() cannot hold
    this workload (cubebeam.py:80: 498 TB), so the arm times OUR C/OpenMP port of its algorithm on all
    host threads (kind "port", see cpu_sample); the 1-core numpy figure rides along in cpu_baseline."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n1 = b + 1
    free_full = 3 * n1 * n1 * A
    vals = []
    sample = None
    for i in range(args.warmup + args.steps):
        sample = cpu_sample(A, b, None, slab_layers=8 if i < args.warmup else 16)
        if i >= args.warmup:
            vals.append(sample["value"])
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": "hex8 beam solved DOF/s", "value": value, "unit": "solved DOF/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": free_full / value * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"cubebeam hex8 cantilever {A}x{b}x{b}", "tol": TOL,
                   "note": "our C/OpenMP port of the reference algorithm on a bounded slab, all host threads, "
                           "extrapolated (see cpu_baseline.sample); not the reference's own code"},
        "cpu_baseline": {k: sample[k] for k in ("value", "unit", "cores", "kind", "sample", "numpy_1core_value")},
        "e2e": {"value": value, "unit": "solved DOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    line["cpu_baseline"]["value"] = value
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def run_ours(args, A, b):
    import torch

    from fea_b200 import _lib, core, cubebeam, model

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    lib = _lib.load()
    if world > 1:
        from fea_b200 import dist as fdist

        return fdist.bench_entry(args, A, b, TOL, E_HEX, NU_HEX, measured_peaks, ClockSampler, cpu_sample)

    workload = f"cubebeam hex8 cantilever {A}x{b}x{b}"
    nodes, elements, constraints, forces = cubebeam.cantilever_case(A, b)
    n_nodes, n_dof = nodes.shape[0], nodes.size
    n_free = int((constraints == 0).sum())

    # pinned host copies (e2e inputs) and device-resident copies (kernel-path inputs)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    nodes_h, elements_h, cons_h, forces_h = pin(nodes), pin(elements), pin(constraints), pin(forces)
    nodes_d = nodes_h.cuda()
    elements_d = elements_h.cuda().to(torch.int32)
    fixed_d = (cons_h.cuda().reshape(-1) != 0).to(torch.uint8)
    loads_d = forces_h.cuda().reshape(-1)
    torch.cuda.synchronize()

    state = {}

    def step_device():
        pat = core.symbolic(elements_d, n_nodes)
        K = core.assemble_hex8(nodes_d, elements_d, E_HEX, NU_HEX, pattern=pat, fixed=fixed_d)
        u, reactions, info = core.solve_system(K, loads_d, tol=TOL)
        state.update(K=K, info=info, u=u, reactions=reactions)

    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()
    lib.fea_profile_enable(1)
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(args.steps):
        step_device()
    ev1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    prof = (4 * __import__("ctypes").c_double)()
    lib.fea_profile_read(prof)
    lib.fea_profile_enable(0)
    ms_per_step = ev0.elapsed_time(ev1) / args.steps
    value = n_free / (ms_per_step / 1e3)
    K, info = state["K"], state["info"]
    nnz = K.nnz
    launches = int(prof[0]) // max(args.steps, 1)
    spmv_ms = prof[2] / max(prof[1], 1.0)

    # stage timings (single extra pass, CUDA events): symbolic / numeric assembly / solve
    def timed(fn):
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        out = fn()
        c.record()
        torch.cuda.synchronize()
        return out, a.elapsed_time(c)

    pat, ms_sym = timed(lambda: core.symbolic(elements_d, n_nodes))
    Kt, ms_num = timed(lambda: core.assemble_hex8(nodes_d, elements_d, E_HEX, NU_HEX, pattern=pat, fixed=fixed_d))
    _, ms_scatter = timed(lambda: _scatter(lib, nodes_d, elements_d, pat, torch))
    del Kt

    hbm_peak, peak_src = measured_peaks()
    alg_bytes = 12 * nnz + 20 * n_dof
    moved_bytes = 8 * nnz + 4 * (nnz // 9) + 4 * (n_nodes + 1) + 16 * n_dof
    achieved = alg_bytes / (spmv_ms / 1e3) / 1e9 if spmv_ms > 0 else None
    roofline = {
        "kernel": "pcg_spmv_tma_kernel<3,2> (ap = K p fused with p.ap)", "bound": "hbm",
        "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak if achieved else None,
        "traffic": committed_traffic(workload), "peak_source": peak_src,
        "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": spmv_ms, "launches_timed": int(prof[1]),
        "format_bytes_per_launch": moved_bytes,
        "frac_of_format_bytes": moved_bytes / (spmv_ms / 1e3) / 1e9 / hbm_peak if spmv_ms > 0 else None,
        "spmv_share_of_step": spmv_ms * info.iterations / ms_per_step if ms_per_step > 0 else None,
    }

    # end to end through the reference-facing API with pinned host buffers
    e2e = None
    if not args.no_e2e:
        arrays = (nodes_h.numpy(), elements_h.numpy(), cons_h.numpy(), forces_h.numpy())
        cubebeam.solve(*arrays)  # warm-up
        torch.cuda.synchronize()
        k_e2e = min(args.steps, 3)
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            u_h, f_h = cubebeam.solve(*arrays)
        torch.cuda.synchronize()
        t_e2e = (time.perf_counter() - t0) / k_e2e
        e2e = {"value": n_free / t_e2e, "unit": "solved DOF/s",
               "h2d_bytes_per_step": int(sum(a.nbytes for a in arrays)),
               "d2h_bytes_per_step": int(u_h.nbytes + f_h.nbytes), "ms_per_step": t_e2e * 1e3, "steps": k_e2e,
               "api": "fea_b200.cubebeam.solve(nodes, elements, constraints, forces) on pinned numpy arrays"}

    cpu = None
    if not args.no_cpu_baseline:
        cpu = cpu_sample(A, b, info.iterations)

    line = {
        "metric": "hex8 beam solved DOF/s", "value": value, "unit": "solved DOF/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "dof": n_dof, "free_dof": n_free, "elements": int(elements.shape[0]),
                   "nnz": nnz, "tol": TOL, "pcg_iterations": info.iterations, "rel_residual": info.rel_residual,
                   "preconditioner": "jacobi", "parallelism": "single GPU",
                   "l2": "inputs larger than L2 (CSR values 5.0 GB >> 126 MB), no flush needed"},
        "stages_ms": {"symbolic": ms_sym, "numeric_assembly_gather": ms_num, "numeric_assembly_scatter_atomics": ms_scatter,
                      "pcg_iteration_avg": (ms_per_step - ms_sym - ms_num) / max(info.iterations, 1)},
        "assembly_elem_per_s": {"symbolic": elements.shape[0] / (ms_sym / 1e3),
                                "numeric_incl_ke": elements.shape[0] / (ms_num / 1e3)},
        "spmv_gb_per_s": achieved,
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks, "gpu_launches": launches,
    }
    print(json.dumps(line))


def _scatter(lib, nodes_d, elements_d, pat, torch):
    vals = torch.zeros(9 * pat.nnz_blocks, dtype=torch.float64, device=nodes_d.device)
    st = torch.zeros(2, dtype=torch.int32, device=nodes_d.device)
    lib.fea_assemble_hex8_scatter(nodes_d.data_ptr(), elements_d.data_ptr(), elements_d.shape[0], E_HEX, NU_HEX,
                                  pat.node_rowptr.data_ptr(), pat.node_colidx.data_ptr(), vals.data_ptr(),
                                  st.data_ptr(), torch.cuda.current_stream().cuda_stream)
    return vals


def main():
    args = parse_args()
    A, b = (int(x) for x in args.workload.split(","))
    if args.impl == "reference":
        run_reference(args, A, b)
    else:
        run_ours(args, A, b)


if __name__ == "__main__":
    main()
