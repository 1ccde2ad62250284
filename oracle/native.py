"""TEST / BENCH INFRASTRUCTURE ONLY — builds and binds oracle/cg_omp.c (the multi-threaded C restatement of
the oracle's Jacobi-CG).  Imported by tests/, bench.py's CPU legs and oracle/make_golden_large.py only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "cg_omp.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libfemb_oracle_cg.so")
_lib = None


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["gcc", "-O3", "-march=x86-64-v2", "-fopenmp", "-shared", "-fPIC", SRC, "-o", LIB, "-lm"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode:
        raise RuntimeError("gcc failed building the oracle CG:\n" + r.stdout)
    return LIB


def load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.femb_oracle_threads.restype = C.c_int
        _lib.femb_oracle_pcg_jacobi.restype = C.c_int
        _lib.femb_oracle_pcg_jacobi.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_double, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_double)]
    return _lib


def threads() -> int:
    return int(load().femb_oracle_threads())


def use_all_cores() -> int:
    """Let the OpenMP CG use every core this process may run on, whatever OMP_NUM_THREADS says (torchrun sets it to
    1 for its workers).  Returns the thread count."""
    import os
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    lib = load()
    lib.femb_oracle_set_threads.argtypes = [C.c_int]
    lib.femb_oracle_set_threads.restype = None
    lib.femb_oracle_set_threads(int(n))
    return threads()


def pcg_jacobi(A, b, rtol=1e-13, max_iter=200_000):
    """Jacobi-PCG on a scipy CSR matrix with all host cores.  Returns (x, dict(iterations, rel_residual, flag))."""
    lib = load()
    A = A.tocsr()
    A.sort_indices()
    indptr = np.ascontiguousarray(A.indptr, dtype=np.int32)
    indices = np.ascontiguousarray(A.indices, dtype=np.int32)
    data = np.ascontiguousarray(A.data, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.zeros_like(b)
    it, rr = C.c_int32(0), C.c_double(0.0)
    flag = lib.femb_oracle_pcg_jacobi(A.shape[0], indptr.ctypes.data, indices.ctypes.data, data.ctypes.data,
                                      b.ctypes.data, x.ctypes.data, float(rtol), int(max_iter), C.byref(it), C.byref(rr))
    return x, {"iterations": int(it.value), "rel_residual": float(rr.value), "flag": int(flag)}
