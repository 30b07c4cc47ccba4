#!/usr/bin/env python
"""Benchmark of the hot path: hex8 cantilever assembled and solved to a 1e-12 recurrence residual.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload A,b]
    torchrun --nproc-per-node N bench.py --gpus N ...                      (N > 1, one rank per GPU)

One "step" = one full pass of the hot path over the workload: symbolic pass + fused Ke/assembly +
Jacobi-PCG to ||r|| <= 1e-12 ||b|| + reaction product.  Metric (BASELINE.json): solved DOF/s =
free DOF / step time.  Workload (all N): BASELINE config "cubebeam hex8 cantilever 400x80x80"
(7,892,883 DOF, 2,560,000 elements, structural nnz 627,797,529); N > 1 partitions the SAME mesh
into z-slabs (strong scaling): a halo exchange per SpMV and world sums of the dot products.

Prints ONE JSON line (rank 0).
  value      inputs resident in HBM (each rank's slab at N > 1), CUDA-event timed, max over ranks.
  e2e        the reference-facing call cubebeam.solve(nodes, elements, constraints, forces) on HOST
             arrays, host->device copies, gather and device->host copy of (displacements, forces)
             inside the timed region; at N > 1 the same call is collective and rank 0 receives the
             full host arrays (fea_b200/dist.py:solve_hex8).
  roofline   the dominant kernel (PCG SpMV), timed live with CUDA events inside the timed solves;
             `achieved` counts the bytes the node-block format has to move, the
             scalar-CSR figure of SURVEY.md §8(d) is reported beside it under its own keys.
  cpu_baseline / --impl reference
             the CPU arm: oracle/fea_oracle_c.c (our C/OpenMP restatement of the reference
             algorithm, the reference itself being single-threaded Python that cannot hold this
             size) on ALL host cores, on the FULL mesh: pattern, Ke + assembly and reduction are
             timed once at full size, the Jacobi-PCG is sampled (>= 500 iterations on the full
             matrix with the driver's K = 20) and scaled to the iteration count of the full solve.
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
METRIC, UNIT = "hex8 beam solved DOF/s", "solved DOF/s"


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
    ap.add_argument("--cpu-iters", type=int, default=0,
                    help="Jacobi-PCG iterations per CPU step (0: sized so that a step takes about 1.5 s, >= 25)")
    ap.add_argument("--cpu-full-solve", action="store_true",
                    help="--impl reference: also run the CPU Jacobi-PCG to convergence once (5-10 min at 400,80)")
    return ap.parse_args()


def mesh_counts(A: int, b: int):
    n1, n3 = b + 1, A + 1
    return {"nodes": n1 * n1 * n3, "dof": 3 * n1 * n1 * n3, "free_dof": 3 * n1 * n1 * A, "elements": A * b * b,
            "nnz": 9 * (3 * n1 - 2) ** 2 * (3 * n3 - 2)}


def workload_config(A: int, b: int, world: int):
    """The static description of the workload: identical in both arms (`--impl ours|reference`)."""
    c = mesh_counts(A, b)
    return {"workload": f"cubebeam hex8 cantilever {A}x{b}x{b}", "dof": c["dof"], "free_dof": c["free_dof"],
            "elements": c["elements"], "nnz": c["nnz"], "tol": TOL, "preconditioner": "jacobi",
            "parallelism": "single GPU" if world == 1 else f"{world} z-slabs of node layers, one rank per GPU",
            "l2": "inputs larger than L2 (CSR values 5.0 GB over all ranks, >= 0.63 GB per rank at N <= 8, "
                  "L2 126 MB); no flush"}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def committed_traffic(key: str):
    """dram bytes per SpMV launch from the committed `ncu --set full` capture, if there is one."""
    try:
        with open(os.path.join(ROOT, "profiles", "spmv_traffic.json")) as fh:
            return json.load(fh).get(key)
    except Exception:
        return None


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


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
# CPU arm: the C/OpenMP oracle on all host cores, full-size mesh, sampled PCG
# ----------------------------------------------------------------------------------------------
def recorded_iterations(A: int, b: int):
    """Iterations the ORACLE's Jacobi-PCG needs on this workload, from the committed fixtures
    (oracle/make_golden_large.py); None for a workload without a fixture."""
    name = {(400, 80): "oracle_config4.npz", (100, 20): "oracle_config3.npz"}.get((A, b))
    if name is None:
        return None
    try:
        return int(np.load(os.path.join(ROOT, "tests", "golden", name))["iterations"])
    except Exception:
        return None


class CpuArm:
    """oracle/fea_oracle_c.c on the host: our C/OpenMP restatement of utils.py:127-239 (dense
    B^T C B Ke in the reference's order) + cubebeam.py:79-108 (scatter, reduction, Jacobi-PCG in place
    of np.linalg.solve).  Plain parallel loops, not cache-blocked or NUMA-pinned.  The thread count is
    pinned to the host's cores explicitly: torchrun exports OMP_NUM_THREADS=1 to every rank."""

    def __init__(self, A: int, b: int):
        from oracle import c_oracle as co
        from oracle import fea_oracle as fo

        self.A, self.b, self.co, self.fo = A, b, co, fo
        self.cores = co.set_threads(host_cores())
        self.counts = mesh_counts(A, b)
        self.setup_s = {}

    def setup(self):
        """Pattern, Ke + assembly and constraint reduction on the FULL mesh, each timed once."""
        co, fo = self.co, self.fo
        nodes, elements, cons, forces = fo.cantilever_case(self.A, self.b)  # host mesh, like the GPU arm's inputs
        t0 = time.perf_counter()
        pattern = co.dof_pattern(elements, nodes.shape[0], 3)
        t1 = time.perf_counter()
        K = co.assemble_hex8(nodes, elements, E_HEX, NU_HEX, pattern=pattern)
        t2 = time.perf_counter()
        free = fo.free_dofs(cons)
        self.Kff = co.reduce_csr(K, free)
        self.ff = forces.flatten()[free]
        t3 = time.perf_counter()
        del K, pattern
        self.setup_s = {"pattern": t1 - t0, "ke_assembly": t2 - t1, "reduce": t3 - t2}
        self.assembly_s = t3 - t0
        co.jacobi_pcg(self.Kff, self.ff, tol=0.0, maxiter=3)  # thread pool / page warm-up
        t4 = time.perf_counter()
        co.jacobi_pcg(self.Kff, self.ff, tol=0.0, maxiter=5)
        self.probe_iter_s = (time.perf_counter() - t4) / 5

    def iters_per_step(self, requested: int) -> int:
        if requested > 0:
            return requested
        return int(min(200, max(25, round(1.5 / max(self.probe_iter_s, 1e-6)))))

    def step(self, iters: int) -> float:
        """`iters` Jacobi-PCG iterations on the full reduced matrix (every iteration costs the same:
        one SpMV + the vector updates), seconds."""
        t0 = time.perf_counter()
        self.co.jacobi_pcg(self.Kff, self.ff, tol=0.0, maxiter=iters)
        return time.perf_counter() - t0

    def full_solve(self):
        t0 = time.perf_counter()
        _, it, rel = self.co.jacobi_pcg(self.Kff, self.ff, tol=TOL)
        return time.perf_counter() - t0, it, rel

    def numpy_1core(self, iterations_full: int, layers: int = 8, pcg_iters: int = 20):
        """The numpy/scipy oracle on one core -- closest to what the reference itself executes --
        on a `layers`-layer slab, scaled linearly (assembly in elements, PCG in nnz x iterations)."""
        fo = self.fo
        layers = min(layers, self.A)
        nodes, elements, cons, forces = fo.cantilever_case(layers, self.b, beam_length=layers / self.A)
        free = fo.free_dofs(cons)
        s0 = time.perf_counter()
        Ke = fo.hex8_ke_batched(nodes, elements, E_HEX, NU_HEX)
        Kn = fo.assemble_csr(elements, Ke, nodes.shape[0], 3)
        Knff = Kn[free][:, free].tocsr()
        s1 = time.perf_counter()
        del Ke
        fo.jacobi_pcg(Knff, forces.flatten()[free], tol=0.0, maxiter=pcg_iters)
        s2 = time.perf_counter()
        total = ((s1 - s0) * self.counts["elements"] / elements.shape[0]
                 + (s2 - s1) / pcg_iters * self.counts["nnz"] / Knff.nnz * iterations_full)
        return self.counts["free_dof"] / total

    def summary(self, step_s: float, iters: int, steps_timed: int, iterations_full: int, iterations_source: str):
        iter_s = step_s / iters
        total = self.assembly_s + iter_s * iterations_full
        s = self.setup_s
        return {
            "value": self.counts["free_dof"] / total, "unit": UNIT, "cores": self.cores, "kind": "port",
            "sample": (f"oracle/fea_oracle_c.c (OUR C/OpenMP port of the reference algorithm; the reference itself is "
                       f"single-threaded Python whose dense solve() cannot hold this size) on {self.cores} host threads, "
                       f"FULL {self.A}x{self.b}x{self.b} mesh: pattern {s['pattern']:.2f}s + Ke/assembly "
                       f"{s['ke_assembly']:.2f}s + reduction {s['reduce']:.2f}s measured once at full size; Jacobi-PCG "
                       f"sampled on the full reduced matrix, {steps_timed} x {iters} iterations at "
                       f"{iter_s * 1e3:.1f} ms/iteration, scaled to the {iterations_full} iterations of the full solve "
                       f"({iterations_source}) = {total:.0f}s per solve"),
            "assembly_elem_per_s": self.counts["elements"] / self.assembly_s,
            "spmv_gb_per_s_csr": (12 * self.Kff.nnz + 20 * self.Kff.shape[0]) / iter_s / 1e9,
            "ms_per_iteration": iter_s * 1e3, "iterations_full": iterations_full, "seconds_per_solve": total,
        }


def run_reference(args, A, b):
    """--impl reference: the CPU arm on rank 0 (other ranks exit).  `ms_per_step` is the measured time of
    one step = one bounded sample of the solve (cpu_iters PCG iterations on the full matrix); `value` is
    free DOF / (full-size assembly, measured once + the full solve's iteration count x measured time per
    iteration)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    arm = CpuArm(A, b)
    arm.setup()
    iters = arm.iters_per_step(args.cpu_iters)
    for _ in range(args.warmup):
        arm.step(min(iters, 5))
    secs = [arm.step(iters) for _ in range(max(args.steps, 1))]
    step_s = float(np.mean(secs))
    it_full = recorded_iterations(A, b)
    src = "recorded by the oracle's own full solve, tests/golden"
    if it_full is None:
        it_full, src = int(21.3 * A), "estimated 21.3 x A, SURVEY.md H1"
    cpu = arm.summary(step_s, iters, len(secs), it_full, src)
    if args.cpu_full_solve:
        t, it, rel = arm.full_solve()
        cpu["full_solve_measured"] = {"seconds": t + arm.assembly_s, "pcg_seconds": t, "iterations": it, "rel_residual": rel}
    cpu["numpy_1core_value"] = arm.numpy_1core(it_full)
    line = {
        "impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(A, b, args.gpus),
        "step_definition": (f"one step = {iters} Jacobi-PCG iterations on the full reduced matrix (a bounded sample of "
                            f"the solve); value = free DOF / seconds_per_solve, see cpu_baseline.sample"),
        "cpu_baseline": cpu,
        "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
class SingleGpu:
    """Device-resident inputs on one GPU; step = symbolic + assembly + PCG + reactions."""

    def __init__(self, mesh, torch, core):
        nodes, elements, constraints, forces = mesh
        self.torch, self.core = torch, core
        self.nodes_d = core.to_device(nodes, torch.float64)
        self.elements_d = core.to_device(elements, torch.int32)
        self.fixed_d = core._fixed_mask(constraints, nodes.size)
        self.loads_d = core.to_device(forces, torch.float64).reshape(-1)
        self.n_nodes = nodes.shape[0]
        self.K = self.info = None

    def step(self):
        core = self.core
        pat = core.symbolic(self.elements_d, self.n_nodes)
        self.K = core.assemble_hex8(self.nodes_d, self.elements_d, E_HEX, NU_HEX, pattern=pat, fixed=self.fixed_d)
        _, _, self.info = core.solve_system(self.K, self.loads_d, tol=TOL)

    def owned_format_bytes(self):
        nnz, n = self.K.nnz, self.K.n_dof
        return 8 * nnz + 4 * (nnz // 9) + 4 * (self.n_nodes + 1) + 16 * n, 12 * nnz + 20 * n

    def stage_times(self):
        torch, core = self.torch, self.core

        def timed(fn):
            a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            out = fn()
            c.record()
            torch.cuda.synchronize()
            return out, a.elapsed_time(c)

        pat, ms_sym = timed(lambda: core.symbolic(self.elements_d, self.n_nodes))
        _, ms_num = timed(lambda: core.assemble_hex8(self.nodes_d, self.elements_d, E_HEX, NU_HEX, pattern=pat,
                                                     fixed=self.fixed_d))
        return {"symbolic": ms_sym, "numeric_assembly": ms_num}


class Slabs:
    """N > 1: each rank's slab resident on its GPU; step = symbolic + assembly + distributed PCG
    + reactions (fea_b200/dist.py:solve_slab)."""

    def __init__(self, mesh, torch, fdist, rank, world):
        nodes, elements, constraints, forces = mesh
        self.torch, self.fdist = torch, fdist
        self.cuts = fdist.default_cuts(nodes.shape[0], world)
        self.plan = fdist.plan_slab(elements, self.cuts, rank)
        self.inp = fdist.upload_slab(nodes, elements, constraints, forces, self.plan)
        self.max_rank_dof = 3 * int(np.diff(self.cuts).max())
        self.K = self.info = None

    def step(self):
        _, _, self.info, self.K = self.fdist.solve_slab(self.inp, E_HEX, NU_HEX, tol=TOL,
                                                         max_rank_dof=self.max_rank_dof)

    def owned_format_bytes(self):
        pl, rp = self.plan, self.K.pattern.node_rowptr
        blocks = int(rp[pl.offset + pl.n_owned] - rp[pl.offset])
        nnz, n = 9 * blocks, 3 * pl.n_owned
        return 8 * nnz + 4 * blocks + 4 * (pl.n_owned + 1) + 16 * n, 12 * nnz + 20 * n

    def stage_times(self):
        return {}


def run_ours(args, A, b):
    import ctypes

    import torch

    from fea_b200 import _lib, core, cubebeam
    from fea_b200 import dist as fdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    lib = _lib.load()
    if world > 1:
        import torch.distributed as dist

        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def world_max(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    def world_sum(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        return float(t)

    counts = mesh_counts(A, b)
    mesh = cubebeam.cantilever_case(A, b)
    n_free = counts["free_dof"]
    runner = SingleGpu(mesh, torch, core) if world == 1 else Slabs(mesh, torch, fdist, rank, world)

    # ---- device-resident steps: `value`
    for _ in range(args.warmup):
        runner.step()
    barrier()
    lib.fea_profile_enable(1)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        runner.step()
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = world_max(ev0.elapsed_time(ev1) / args.steps)
    prof = (4 * ctypes.c_double)()
    lib.fea_profile_read(prof)
    lib.fea_profile_enable(0)
    value = n_free / (ms_per_step / 1e3)
    info = runner.info
    launches = int(world_sum(float(prof[0]))) // max(args.steps, 1)
    spmv_ms = prof[2] / max(prof[1], 1.0)          # this rank's PCG SpMV, sampled inside the timed solves
    fmt_bytes, csr_bytes = runner.owned_format_bytes()

    # ---- roofline of the dominant kernel: format bytes of this rank's rows / its own launch time
    hbm_peak, peak_src = measured_peaks()
    mine = fmt_bytes / (spmv_ms / 1e3) / 1e9 if spmv_ms > 0 else 0.0
    achieved = world_sum(mine) / world           # mean over ranks of the per-GPU figure
    slowest_ms = world_max(spmv_ms)
    csr_equiv = world_sum(csr_bytes / (spmv_ms / 1e3) / 1e9 if spmv_ms > 0 else 0.0) / world
    traffic_key = f"{A}x{b}x{b}/N{world}"
    roofline = {
        "kernel": ("pcg_spmv_tma_kernel<3,2> (ap = K p fused with p.ap"
                   + (", halo-gated face tiles" if world > 1 else "") + ")"),
        "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
        "traffic": committed_traffic(traffic_key), "peak_source": peak_src,
        "bytes_counted": "node-block CSR format bytes of the rank's rows: 8 B/value + 4 B per 3x3 block + 4 B per "
                         "node rowptr + 16 B per DOF (x read, y written); per GPU, mean over ranks",
        "format_bytes_per_launch": fmt_bytes, "avg_launch_ms": spmv_ms, "slowest_rank_avg_launch_ms": slowest_ms,
        "launches_timed": int(prof[1]),
        "csr_equivalent_gb_per_s": csr_equiv, "csr_equivalent_frac": csr_equiv / hbm_peak,
        "csr_equivalent_bytes_per_launch": csr_bytes,
        "spmv_share_of_step": slowest_ms * info.iterations / ms_per_step if ms_per_step > 0 else None,
    }

    # ---- where a step goes (one extra untimed pass)
    stages = runner.stage_times()
    iter_us = None
    if world == 1:
        iter_us = (ms_per_step - stages["symbolic"] - stages["numeric_assembly"]) / max(info.iterations, 1) * 1e3
        stages["pcg_iteration_avg_us"] = iter_us
        stages["pcg_spmv_avg_us"] = spmv_ms * 1e3
    assembly_rates = None
    if world == 1:
        assembly_rates = {"symbolic": counts["elements"] / (stages["symbolic"] / 1e3),
                          "numeric_incl_ke": counts["elements"] / (stages["numeric_assembly"] / 1e3)}

    # ---- end to end through the reference-facing API with host buffers
    e2e = None
    if not args.no_e2e:
        pin = (lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()) if world == 1 else (lambda a: a)
        arrays = tuple(pin(a) for a in mesh)
        cubebeam.solve(*arrays)  # warm-up (IPC set-up, caches)
        k_e2e = min(max(args.steps, 1), 3)
        barrier()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            u_h, f_h = cubebeam.solve(*arrays)
        barrier()
        t_e2e = world_max((time.perf_counter() - t0) / k_e2e)
        if world == 1:
            h2d = int(sum(a.nbytes for a in arrays))
        else:
            h2d = int(world_sum(float(runner.inp.h2d_bytes)))
        d2h = int(world_sum(float(u_h.nbytes + f_h.nbytes) if u_h is not None else 0.0))
        e2e = {"value": n_free / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": t_e2e * 1e3, "steps": k_e2e,
               "api": "fea_b200.cubebeam.solve(nodes, elements, constraints, forces): host numpy arrays in, host "
                      "(displacements, forces) out" + ("" if world == 1 else
                                                       " on rank 0 (collective call; every rank copies its slab up, "
                                                       "rank 0 gathers over NVLink and copies the result down)")}
        if world > 1:
            fdist.STAGE_PROFILE.clear()
            fdist.STAGE_PROFILE["enabled"] = True
            barrier()
            cubebeam.solve(*arrays)
            fdist.STAGE_PROFILE["enabled"] = False
            stages = {k: round(v, 3) for k, v in fdist.STAGE_PROFILE.items() if k != "enabled"}
            if "pcg" in stages:
                stages["pcg_iteration_avg_us"] = stages["pcg"] / max(info.iterations, 1) * 1e3
            stages["pcg_spmv_avg_us_slowest_rank"] = slowest_ms * 1e3

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        arm = CpuArm(A, b)
        arm.setup()
        iters = args.cpu_iters or int(min(400, max(50, round(8.0 / max(arm.probe_iter_s, 1e-6)))))
        cpu = arm.summary(arm.step(iters), iters, 1, info.iterations, "the count of the GPU solve in this run")
        cpu["numpy_1core_value"] = arm.numpy_1core(info.iterations)

    if rank == 0:
        cfg = workload_config(A, b, world)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "solve": {"pcg_iterations": info.iterations, "rel_residual": info.rel_residual,
                      "solver": ("fea_pcg_solve" if world == 1 else
                                 {"p2p": "fea_pcg_solve_p2p (NVLink peer-memory exchange fused into the solver kernels, "
                                         "no NCCL per iteration)",
                                  "nccl": "torch.distributed driver (NCCL halo send/recv + all-reduced dots)"}
                                 .get(fdist.SOLVER_USED["kind"], "?"))},
            "stages_ms" if world == 1 else "stages_ms_rank0": stages,
            "assembly_elem_per_s": assembly_rates,
            "spmv_gb_per_s": achieved,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks, "gpu_launches": launches,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    A, b = (int(x) for x in args.workload.split(","))
    if args.impl == "reference":
        run_reference(args, A, b)
    else:
        run_ours(args, A, b)


if __name__ == "__main__":
    main()
