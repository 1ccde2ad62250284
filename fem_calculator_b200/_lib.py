"""ctypes binding of libfemb200.so (the C ABI declared in include/femb200.h).

The product path has no CPU fallback: if the library is missing or no CUDA device is
usable, the calls raise.  Nothing here imports ``oracle``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FEMB_LIB") or os.path.join(HERE, "libfemb200.so")   # FEMB_LIB: A/B another build

FEMB_OK, FEMB_ERR_ARG, FEMB_ERR_CUDA, FEMB_ERR_NOT_CONVERGED, FEMB_ERR_SINGULAR, FEMB_ERR_NOMEM = 0, -1, -2, -3, -4, -5
MAT_K, MAT_M = 0, 1
SOLVER_AUTO, SOLVER_PCG, SOLVER_CHAIN, SOLVER_DENSE = 0, 1, 2, 3
PRECOND_NONE, PRECOND_JACOBI, PRECOND_BLOCK_JACOBI, PRECOND_TWO_LEVEL, PRECOND_AUTO, PRECOND_LINES = 0, 1, 2, 3, 4, 5
OP_AUTO, OP_BSR, OP_EBE = 0, 1, 2


class SolveOpts(C.Structure):
    _fields_ = [("method", C.c_int32), ("precond", C.c_int32), ("max_iter", C.c_int32),
                ("check_every", C.c_int32), ("rtol", C.c_double), ("profile", C.c_int32),
                ("op", C.c_int32)]


class EigOpts(C.Structure):
    _fields_ = [("k", C.c_int32), ("block", C.c_int32), ("max_iter", C.c_int32), ("op", C.c_int32),
                ("rtol", C.c_double), ("lambda_min", C.c_double), ("precond", C.c_int32), ("reserved", C.c_int32),
                ("accept_rtol", C.c_double)]


class Stats(C.Structure):
    _fields_ = [("method_used", C.c_int32), ("iterations", C.c_int32), ("converged", C.c_int32),
                ("spmv_launches", C.c_int32), ("kernel_launches", C.c_int32), ("spmv_timed", C.c_int32),
                ("op_used", C.c_int32), ("coarse_dim", C.c_int32), ("precond_used", C.c_int32), ("reserved", C.c_int32),
                ("rel_residual", C.c_double), ("device_ms", C.c_double), ("spmv_ms", C.c_double),
                ("update_ms", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ }


class FembError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"femb200 error {code}: {msg}")
        self.code = code


_P = C.c_void_p
_F64 = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_I64 = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_I32 = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_U8 = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")

# name -> (restype, argtypes); every symbol include/femb200.h declares
SIGNATURES = {
    "femb_version": (C.c_int, []),
    "femb_device_count": (C.c_int, []),
    "femb_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "femb_destroy": (None, [_P]),
    "femb_last_error": (C.c_char_p, [_P]),
    "femb_frame_set_mesh": (C.c_int, [_P, C.c_int64, C.c_int64, _F64, _I64, _I32, C.c_int32, _F64,
                                      C.c_double, C.c_double, C.c_double]),
    "femb_frame_elements": (C.c_int, [_P, _P, _P]),
    "femb_tet10_set_mesh": (C.c_int, [_P, C.c_int64, C.c_int64, _F64, _I64, C.c_double, C.c_double]),
    "femb_tet10_elements": (C.c_int, [_P, _P]),
    "femb_tet10_negative_detj": (C.c_int64, [_P]),
    "femb_assemble": (C.c_int, [_P]),
    "femb_get_csr_size": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "femb_get_csr": (C.c_int, [_P, C.c_int, _I32, _I32, _F64]),
    "femb_set_bc": (C.c_int, [_P, C.c_int64, _P, _F64, _P]),
    "femb_solve_static": (C.c_int, [_P, C.POINTER(SolveOpts), C.c_int, _P, _P, C.POINTER(Stats)]),
    "femb_apply_k": (C.c_int, [_P, C.c_int, C.c_int, _F64, _F64, C.POINTER(C.c_int32)]),
    "femb_modal": (C.c_int, [_P, C.POINTER(EigOpts), _P, _P, C.POINTER(C.c_int32), C.POINTER(Stats)]),
    "femb_frame_stress": (C.c_int, [_P, _P, _P]),
    "femb_host_register": (C.c_int, [_P, C.c_void_p, C.c_int64]),
    "femb_host_unregister": (C.c_int, [_P, C.c_void_p]),
    "femb_frame_batch_solve": (C.c_int, [_P, C.c_int64, C.c_int64, _F64, _F64, C.c_double, C.c_double,
                                         _U8, _F64, _P, C.POINTER(Stats)]),
    "femb_dist_unique_id": (C.c_int, [_P]),
    "femb_dist_init": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "femb_dist_finalize": (None, [_P]),
    "femb_dist_set_halo": (C.c_int, [_P, C.c_int64, C.c_int32, _P, _P, _P, _P, _P]),
    "femb_dist_solve_static": (C.c_int, [_P, C.POINTER(SolveOpts), C.c_int, _P, _P, C.POINTER(Stats)]),
    "femb_dist_p2p_export": (C.c_int, [_P, _P]),
    "femb_dist_p2p_import": (C.c_int, [_P, _P, _P]),
    "femb_time_kernel": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "femb_timer": (C.c_int, [_P, C.c_int, C.POINTER(C.c_double)]),
    "femb_io_bytes": (None, [C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_int]),
    "femb_symbolic_pattern": (C.c_int, [C.c_int64, C.c_int64, C.c_int32, _I64, C.POINTER(C.c_int64), _P, _P]),
    "femb_symbolic_aggregates": (C.c_int, [C.c_int64, _P, C.c_int32, _P]),
    "femb_symbolic_lines": (C.c_int, [C.c_int64, C.c_int64, _P, _P, C.c_double, C.c_int32, C.POINTER(C.c_int64),
                                      C.POINTER(C.c_int64), _P, _P, _P, _P]),
    "femb_symbolic_coarse": (C.c_int, [C.c_int64, C.c_int64, _P, _P, C.c_int32, _P, _P, C.POINTER(C.c_int64), _P, _P]),
    "femb_dist_set_lines": (C.c_int, [_P, C.c_int32, _P, _P, _P, _P, _P]),
    "femb_symbolic_line_bundles": (C.c_int, [C.c_int64, C.c_int64, _P, _P, C.c_int32, _P, _P, _P, C.POINTER(C.c_int64),
                                             C.POINTER(C.c_int64), C.POINTER(C.c_double), _P, _P]),
}

_lib = None


def load():
    """Load libfemb200.so (built in-tree by fem_calculator_b200/build.py).  Raises if the
    library is absent — there is deliberately no pure-Python or CPU substitute."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not found: run `python -m fem_calculator_b200.build` "
                          "(or __graft_entry__.build()); femb200 has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def ptr(a):
    """void* of a C-contiguous numpy array, or NULL for None."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def device_count() -> int:
    return int(load().femb_device_count())


def io_bytes(reset=False):
    """(host->device, device->host) bytes copied through the library since the last reset."""
    a, b = C.c_int64(), C.c_int64()
    load().femb_io_bytes(C.byref(a), C.byref(b), int(reset))
    return a.value, b.value
