"""ctypes binding of libfea_b200.so (C ABI declared in include/fea_b200.h).

There is deliberately NO fallback: if the CUDA library is missing or fails to load, every product
entry point raises.  The CPU oracle under oracle/ is test infrastructure and is never imported here.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_int32, c_int64, c_size_t, c_void_p

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FEA_LIB_PATH") or os.path.join(HERE, "csrc", "libfea_b200.so")

FEA_OK = 0
FEA_ERR_INVALID = 1
FEA_ERR_CUDA = 2
FEA_ERR_JACOBIAN = 3
FEA_ERR_BREAKDOWN = 4
FEA_ERR_MAXITER = 5
FEA_ERR_WORKSPACE = 6
FEA_ERR_DEGENERATE = 7

ASSEMBLE_FULL = 0
ASSEMBLE_ELIMINATED = 1

PCG_STATE_BYTES = 256
PCG_PARTIALS = 2048
# double indices into the state
PCG_RZ, PCG_BNORM2, PCG_RZ_NEW, PCG_RR, PCG_PAP, PCG_TOL2, PCG_RR_FINAL = 0, 1, 2, 3, 4, 5, 6
# int32 indices into the state
PCG_ITER_I32, PCG_DONE_I32, PCG_STATUS_I32, PCG_MAXITER_I32 = 32, 33, 34, 35

JACOBIAN_MESSAGE = "Jacobian determinant is non-positive. Check the element shape."  # utils.py:213-215


class PcgResult(ctypes.Structure):
    _fields_ = [
        ("iterations", c_int32),
        ("status", c_int32),
        ("rel_residual", c_double),
        ("bnorm", c_double),
    ]


FEA_ERR_PEER = 8
FEA_ERR_STAGNATION = 9
MAX_PEERS = 8


class PeerComm(ctypes.Structure):
    """fea_peer_comm of include/fea_b200.h."""

    _fields_ = [
        ("world", c_int32), ("rank", c_int32), ("lower_peer", c_int32), ("upper_peer", c_int32),
        ("comm", c_void_p * MAX_PEERS),
        ("own_offset_nodes", c_int64),
        ("send_lower_first", c_int64), ("send_lower_count", c_int64), ("send_lower_dst", c_int64),
        ("send_upper_first", c_int64), ("send_upper_count", c_int64), ("send_upper_dst", c_int64),
        ("epoch", c_int64),
        ("boundary_lower_nodes", c_int64), ("boundary_upper_nodes", c_int64),
        ("algo", c_int32), ("reserved", c_int32),
        ("max_rank_dof", c_int64),
    ]


P = c_void_p  # every device / host pointer crosses the ABI as a plain address

# name -> (restype, argtypes); mirrors include/fea_b200.h one to one
PROTOTYPES = {
    "fea_version": (c_char_p, []),
    "fea_last_cuda_error": (c_char_p, []),
    "fea_profile_enable": (None, [c_int32]),
    "fea_profile_read": (None, [P]),
    "fea_ke_hex8": (c_int32, [P, P, c_int64, c_double, c_double, P, P, P]),
    "fea_ke_beam": (c_int32, [P, P, c_int64, P, P]),
    "fea_ke_truss": (c_int32, [P, P, P, c_int64, P, P, P]),
    "fea_csr_symbolic_workspace": (c_size_t, [c_int64, c_int64, c_int32]),
    "fea_csr_symbolic_count": (c_int32, [P, c_int64, c_int32, c_int64, P, P, P, P, P, c_size_t, P]),
    "fea_csr_symbolic_fill": (c_int32, [P, c_int64, c_int32, c_int64, P, P, P, P, c_int32, P]),
    "fea_csr_expand": (c_int32, [c_int64, c_int32, P, P, P, P, P]),
    "fea_assemble_hex8": (c_int32, [P, P, c_int64, c_int64, c_double, c_double, P, P, P, P, c_int32, P, c_int32,
                                    P, P, P, P]),
    "fea_assemble_beam": (c_int32, [P, P, P, c_int64, c_int64, P, P, P, P, P, c_int32, P, P, P]),
    "fea_assemble_truss": (c_int32, [P, P, P, c_int64, c_int64, P, P, P, P, P, c_int32, P, P, P, P]),
    "fea_jacobi_dinv": (c_int32, [c_int64, c_int32, P, P, P, P, P, P]),
    "fea_spmv": (c_int32, [c_int64, c_int32, P, P, P, c_int32, P, P, P]),
    "fea_spmm": (c_int32, [c_int64, c_int32, P, P, P, P, P, c_int32, P]),
    "fea_pcg_workspace": (c_size_t, [c_int64]),
    "fea_pcg_solve": (c_int32, [c_int64, c_int32, P, P, P, c_int32, P, P, P, c_double, c_int32, P, c_size_t, P,
                                ctypes.POINTER(PcgResult), P]),
    "fea_pcg_init": (c_int32, [c_int64, P, P, P, P, P, c_double, c_int32, P, P, P]),
    "fea_pcg_step_spmv": (c_int32, [c_int64, c_int32, P, P, P, c_int32, P, P, c_int64, P, P, P]),
    "fea_pcg_step_update": (c_int32, [c_int64, P, P, P, P, P, P, P, P]),
    "fea_pcg_step_direction": (c_int32, [c_int64, P, P, P, P, P, P]),
    "fea_slab_scan": (c_int32, [P, c_int32, c_int64, c_int32, P, c_int32, P]),
    "fea_comm_bytes": (c_size_t, [c_int64]),
    "fea_comm_alloc": (c_int32, [c_size_t, ctypes.POINTER(c_void_p)]),
    "fea_comm_free": (c_int32, [P]),
    "fea_comm_ipc_export": (c_int32, [P, P]),
    "fea_comm_ipc_open": (c_int32, [P, ctypes.POINTER(c_void_p)]),
    "fea_comm_ipc_close": (c_int32, [P]),
    "fea_peer_push": (c_int32, [P, P, P, P, c_int64, P, P, P, c_int64, P, c_int64, P]),
    "fea_peer_wait": (c_int32, [P, c_int32, c_int32, P, c_int64, P]),
    "fea_comm_error": (c_int32, [P, P, P]),
    "fea_pcg_solve_p2p": (c_int32, [c_int64, c_int32, P, P, P, c_int32, P, P, P, c_double, c_int32, P, c_size_t, P,
                                    ctypes.POINTER(PeerComm), ctypes.POINTER(PcgResult), P]),
    "fea_chain_solve_workspace": (c_size_t, [c_int64, c_int32, c_int32]),
    "fea_chain_solve": (c_int32, [c_int64, c_int32, P, P, P, P, P, P, c_int32, P, c_size_t, P, P]),
    "fea_pcg_multi_workspace": (c_size_t, [c_int64, c_int32]),
    "fea_pcg_solve_multi": (c_int32, [c_int64, c_int32, P, P, P, P, P, P, c_int32, c_double, c_int32, P, c_size_t,
                                      P, ctypes.POINTER(PcgResult), P]),
    "fea_pcg_multi_layout": (c_int32, [c_int32, P]),
    "fea_pcg_multi_init": (c_int32, [c_int64, c_int32, P, P, P, P, c_double, c_int32, P, c_size_t, P]),
    "fea_pcg_multi_activate": (c_int32, [c_int64, c_int32, P, P]),
    "fea_pcg_multi_step_spmm": (c_int32, [c_int64, c_int32, P, P, P, P, c_int64, c_int32, P, P]),
    "fea_pcg_multi_step_update": (c_int32, [c_int64, c_int32, P, P, P, P, P]),
    "fea_pcg_multi_step_direction": (c_int32, [c_int64, c_int32, P, P, P, P]),
    "fea_truss_member_forces": (c_int32, [P, P, P, c_int64, c_int64, P, P, P, P, c_int32, P]),
    "fea_truss_relax": (c_int32, [P, P, P, c_int64, c_int64, P, P, P, P, c_int64, c_double, c_int32, P, P, P, c_int32,
                                  P]),
    "fea_mesh_quad_grid": (c_int32, [c_int64, c_int64, c_double, c_double, P, P, P]),
    "fea_mesh_tube_section": (c_int32, [c_int64, c_double, c_double, P, P, P]),
    "fea_mesh_lattice_members": (c_int64, [c_int64]),
    "fea_mesh_lattice": (c_int32, [c_int64, c_double, P, P, P, P, P, P]),
    "fea_beam_moment_shear": (c_int32, [P, P, P, c_int64, P, P, P]),
    "fea_mesh_extrude": (c_int32, [P, c_int64, P, c_int64, P, c_int64, P, P, P]),
}

_LIB = None


class FeaLibraryError(RuntimeError):
    pass


def load(path: str | None = None) -> ctypes.CDLL:
    """Load the shared library and attach the prototypes.  Raises if it is missing."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    path = path or LIB_PATH
    if not os.path.isfile(path):
        raise FeaLibraryError(
            f"{path} not found: build it with `python -m fea_b200.build` (nvcc, sm_100a). "
            "fea_b200 has no CPU fallback."
        )
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError here = header and library disagree
        fn.restype = restype
        fn.argtypes = argtypes
    _LIB = lib
    return lib


def check(rc: int, what: str = "") -> None:
    """Map an FEA_ERR_* return code of an entry point to a Python exception."""
    if rc == FEA_OK:
        return
    lib = load()
    if rc == FEA_ERR_CUDA:
        raise FeaLibraryError(f"{what}: CUDA error: {lib.fea_last_cuda_error().decode()}")
    if rc == FEA_ERR_INVALID:
        raise ValueError(f"{what}: invalid argument")
    if rc == FEA_ERR_WORKSPACE:
        raise FeaLibraryError(f"{what}: workspace too small")
    raise FeaLibraryError(f"{what}: error code {rc}")


def raise_for_status(status_host: np.ndarray) -> None:
    """Map a data-dependent status slot {code, 0x7fffffff - index} to the reference's exceptions."""
    code = int(status_host[0])
    if code == FEA_OK:
        return
    index = 0x7FFFFFFF - int(status_host[1])
    if code == FEA_ERR_JACOBIAN:
        err = ValueError(JACOBIAN_MESSAGE)  # same text as utils.py:213-215
        err.element = index
        raise err
    if code == FEA_ERR_DEGENERATE:
        err = ValueError("zero-length truss member")
        err.element = index
        raise err
    if code == FEA_ERR_INVALID:
        raise ValueError(f"mesh is not a chain at node {index}")
    if code == FEA_ERR_BREAKDOWN:
        raise np.linalg.LinAlgError("Singular matrix")  # what np.linalg.solve raises, cubebeam.py:98
    if code == FEA_ERR_MAXITER:
        raise np.linalg.LinAlgError("PCG did not converge within max_iter")
    if code == FEA_ERR_STAGNATION:  # under-constrained body: the reduced K is singular (cubebeam.py:98)
        raise np.linalg.LinAlgError("Singular matrix (PCG residual stopped improving)")
    if code == FEA_ERR_PEER:
        raise FeaLibraryError("multi-GPU solve: a peer rank never delivered its halo / partial sum, "
                              "or the ranks disagree on the recurrence")
    raise FeaLibraryError(f"device status {code}")
