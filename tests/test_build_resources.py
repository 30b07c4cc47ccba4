"""Register / local-memory budget of the hot kernels, read from the built objects with cuobjdump (no
GPU needed).  The PCG SpMV is a persistent kernel that needs THREE resident CTAs per SM (352 threads
each): above 56 registers per thread, or with spills, it silently drops to one or two and loses up to
40 % -- which happened twice during development (a second inlined gather loop; a by-value kernel
argument indexed by lane), each time looking like a multi-GPU communication problem."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "fea_b200", "csrc")


def resources(obj):
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.isfile(exe):
        pytest.skip("cuobjdump not available")
    path = os.path.join(CSRC, obj)
    if not os.path.isfile(path):
        from fea_b200 import build

        build.build_library()
    out = subprocess.run([exe, "--dump-resource-usage", path], capture_output=True, text=True, check=True).stdout
    res, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            name = m.group(1)
        m = re.search(r"REG:(\d+) STACK:(\d+)", line)
        if m and name:
            res[name] = (int(m.group(1)), int(m.group(2)))
    return res


def test_pcg_spmv_keeps_three_ctas_per_sm():
    res = resources("pcg.o")
    hot = {k: v for k, v in res.items() if "pcg_spmv_tma_kernelILi3ELi2E" in k}
    assert len(hot) == 2, sorted(res)  # plain and halo-gated variant of <D = 3, G = 2>
    for name, (regs, stack) in hot.items():
        assert regs <= 56 and stack == 0, (name, regs, stack)  # 3 x 352 threads x 56 registers <= 65536
    plain = [v for k, v in res.items() if "15spmv_tma_kernelILi3ELi2E" in k]
    assert plain and plain[0][0] <= 56 and plain[0][1] == 0


def test_vector_kernels_fit_their_launch_bounds():
    res = resources("pcg.o")
    cgcg = [v for k, v in res.items() if "pcg_cgcg_kernel" in k]
    assert cgcg and cgcg[0][0] <= 80 and cgcg[0][1] == 0  # __launch_bounds__(256, 3)
    halo = [v for k, v in resources("p2p.o").items() if "p2p_halo_kernel" in k]
    assert halo and halo[0][0] <= 48  # must fit beside the persistent SpMV (6400 registers free per SM)


def test_persistent_pcg_keeps_two_ctas_per_sm_without_spills():
    """pcg_fused_kernel<3,3> (config 3's shape: 3 consumer groups + producer = 512 threads, 2 CTAs per SM) must stay at
    64 registers with NO local memory: inlining the gather loop into it spilled inside the loop and cost 70 % (36
    against 21 us per iteration) -- the loop lives in a __noinline__ function for that reason."""
    res = resources("pcg.o")
    hot = [v for k, v in res.items() if "pcg_fused_kernelILi3ELi3E" in k]
    assert hot and hot[0][0] <= 64 and hot[0][1] == 0, hot


def test_assembly_kernels_keep_five_ctas_per_sm():
    res = resources("assemble.o")
    for key in ("assemble_hex8_affine_kernel", "20assemble_hex8_kernel"):
        hit = [v for k, v in res.items() if key in k]
        assert hit and hit[0][0] <= 102 and hit[0][1] == 0, (key, hit)  # 5 x 128 threads x 102 registers <= 65536
