"""bench.py contract checks that need no GPU: the CPU arm (`--impl reference`) prints ONE JSON line
with the keys the driver reads, on the same `config` as our arm, with the host's cores whatever
OMP_NUM_THREADS the launcher exported, and with a step time that is really the time of a step."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_reference(extra_env=None, gpus=1, steps=2, warmup=1):
    env = dict(os.environ)
    env.update(extra_env or {})
    t0 = time.perf_counter()
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", str(gpus),
                          "--steps", str(steps), "--warmup", str(warmup), "--workload", "40,8", "--cpu-iters", "30"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    wall = time.perf_counter() - t0
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    return lines, wall


def test_reference_arm_json_line():
    import bench

    lines, wall = run_reference()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "hex8 beam solved DOF/s" and d["unit"] == "solved DOF/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["dtype"] == "f64" and d["vs_baseline"] is None
    # the same static config as our arm prints for this workload and N
    assert d["config"] == bench.workload_config(40, 8, 1)
    # a step is a measured, bounded sample: steps x ms_per_step fits inside the run
    assert d["ms_per_step"] * d["steps"] / 1e3 < wall
    cpu = d["cpu_baseline"]
    assert cpu["kind"] == "port" and cpu["value"] == d["value"]
    assert cpu["cores"] == len(os.sched_getaffinity(0))
    assert cpu["numpy_1core_value"] > 0 and cpu["iterations_full"] > 0
    assert "OUR C/OpenMP port" in cpu["sample"] and "FULL 40x8x8 mesh" in cpu["sample"]
    # value = free DOF / (assembly once + iterations x time per iteration)
    free = bench.mesh_counts(40, 8)["free_dof"]
    assert abs(cpu["seconds_per_solve"] * d["value"] / free - 1.0) < 1e-9
    assert d["e2e"] == {"value": d["value"], "unit": "solved DOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_ignores_launcher_thread_cap():
    """torchrun exports OMP_NUM_THREADS=1 to every rank: the CPU arm still uses the host's cores."""
    lines, _ = run_reference({"OMP_NUM_THREADS": "1", "RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0"}, gpus=2)
    d = json.loads(lines[0])
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert d["n_gpus"] == 2


def test_reference_arm_other_ranks_exit_quietly():
    lines, _ = run_reference({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, gpus=2, steps=1, warmup=0)
    assert lines == []


def test_recorded_iteration_counts():
    import bench

    assert bench.recorded_iterations(100, 20) == 2127  # 2,127 iterations at 100x20x20
    assert bench.recorded_iterations(7, 3) is None
