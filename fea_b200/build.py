"""Build libfea_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m fea_b200.build [--force] [--verbose]

The shared library lands next to the sources (fea_b200/csrc/libfea_b200.so): it is git-ignored
but travels to the GPU box with the repo snapshot.  No JIT cache, no torch extension machinery --
the C ABI (include/fea_b200.h) is loaded with ctypes.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
INCLUDE = os.path.join(ROOT, "include")
LIB_PATH = os.path.join(CSRC, "libfea_b200.so")
SOURCES = ["symbolic.cu", "element_ke.cu", "assemble.cu", "pcg.cu", "p2p.cu", "multi.cu", "misc.cu", "host.cu", "mesh.cu", "truss.cu", "chain.cu"]
HEADERS = ["common.cuh", "hex8.cuh", "spmv.cuh"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--fmad=true",
    "-I", INCLUDE, "-I", CSRC,
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.isfile(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


CHECKED_LIB_PATH = os.path.join(CSRC, "libfea_b200_checked.so")


def build_library(force: bool = False, verbose: bool = False, checked: bool = False) -> str:
    """checked=True: -DFEA_CHECKED (device-side assertions, csrc/common.cuh) into libfea_b200_checked.so,
    objects under csrc/checked/."""
    nvcc = find_nvcc()
    lib_path = CHECKED_LIB_PATH if checked else LIB_PATH
    obj_dir = os.path.join(CSRC, "checked") if checked else CSRC
    os.makedirs(obj_dir, exist_ok=True)
    flags = NVCC_FLAGS + (["-DFEA_CHECKED"] if checked else [])
    sources = [s for s in SOURCES if os.path.isfile(os.path.join(CSRC, s))]
    headers = sorted(f for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    common_deps = [os.path.join(CSRC, h) for h in headers] + [os.path.join(INCLUDE, "fea_b200.h"), __file__]
    objs, jobs = [], []
    for src in sources:
        src_path = os.path.join(CSRC, src)
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [src_path] + common_deps):
            cmd = [nvcc, *flags, "-c", src_path, "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)

    def run(cmd):
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
        return res.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as pool:
            for log in pool.map(run, jobs):
                if verbose and log:
                    print(log, file=sys.stderr)
    if jobs or force or _stale(lib_path, objs):
        run([nvcc, "-shared", "-o", lib_path, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static",
             "-lpthread", "-ldl", "-lrt"])
    return lib_path


if __name__ == "__main__":
    path = build_library(force="--force" in sys.argv, verbose="--verbose" in sys.argv, checked="--checked" in sys.argv)
    print(path)
