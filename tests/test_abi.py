"""CPU: the C-ABI library builds, loads, and exports every symbol include/fea_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from fea_b200 import build

    return build.build_library()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "fea_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fea_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_path():
    names = declared_symbols()
    for must in ("fea_ke_hex8", "fea_ke_beam", "fea_ke_truss", "fea_csr_symbolic_count", "fea_csr_symbolic_fill",
                 "fea_assemble_hex8", "fea_spmv", "fea_spmm", "fea_pcg_solve", "fea_pcg_solve_multi"):
        assert must in names


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, missing


def test_python_prototypes_match_header(lib_path):
    from fea_b200 import _lib

    assert sorted(_lib.PROTOTYPES) == declared_symbols()
    lib = _lib.load()
    assert lib.fea_version().decode().startswith("fea_b200")
    assert "sm_100a" in lib.fea_version().decode()


def test_argument_validation_needs_no_gpu(lib_path):
    from fea_b200 import _lib

    lib = _lib.load()
    # null pointers are rejected before any CUDA call
    assert lib.fea_ke_hex8(None, None, 1, 1.0, 0.3, None, None, None) == _lib.FEA_ERR_INVALID
    assert lib.fea_spmv(0, 3, None, None, None, 27, None, None, None) == _lib.FEA_ERR_INVALID
    assert lib.fea_pcg_workspace(1000) >= 3 * 8 * 1000
    assert lib.fea_csr_symbolic_workspace(1000, 100, 8) >= 4 * 1000


def test_product_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from fea_b200 import _lib, cubebeam

    nodes, elements, constraints, forces = cubebeam.cantilever_case(2, 1)
    with pytest.raises(_lib.FeaLibraryError):
        cubebeam.solve(nodes, elements, constraints, forces)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "fea_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
