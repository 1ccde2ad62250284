"""TEST/BENCH INFRASTRUCTURE ONLY — times the CPU oracle (oracle/ref_sparse.py, the sparse
restatement of the reference's numpy/scipy path) on a bounded sample of a lattice-frame
workload and scales it to the full configuration.  Used only by bench.py's ``cpu_baseline``
leg and ``--impl reference`` arm.

The reference itself (dense 6N x 6N + LAPACK, BeamSolver.py:360-418) cannot run beyond a few
thousand DOF, so the CPU baseline of record is the scalable port: vectorised element
formation -> scipy COO->CSR -> Jacobi-preconditioned scipy CG.
scipy's sparse mat-vec and numpy's element-wise kernels are single-threaded: cores = 1.
"""
from __future__ import annotations

import time

import numpy as np
import scipy.sparse.linalg as spla

from . import ref_sparse as S


def lattice_static_sample(mesh, elem_sec, props, bc_data, E, nu, cg_iters=150):
    """Assemble the sample lattice on the CPU and run ``cg_iters`` Jacobi-PCG iterations.
    Returns dict(n_elem, n_free, nnz, t_assemble, t_bc, t_per_iter)."""
    conn = mesh.cells_dict["line"]
    t0 = time.perf_counter()
    K, M = S.frame_assemble(mesh.points, conn, elem_sec, props, E, nu)
    t_asm = time.perf_counter() - t0
    t0 = time.perf_counter()
    fixed, free, f = S.frame_bc(mesh, bc_data)
    Kff = K[free][:, free].tocsr()
    ff = f[free]
    t_bc = time.perf_counter() - t0
    d = Kff.diagonal()
    Minv = spla.LinearOperator(Kff.shape, matvec=lambda x: x / d)
    t0 = time.perf_counter()
    spla.cg(Kff, ff, rtol=1e-30, atol=0.0, M=Minv, maxiter=cg_iters)
    t_it = (time.perf_counter() - t0) / cg_iters
    return {"n_elem": len(conn), "n_free": len(free), "nnz": int(Kff.nnz), "t_assemble": t_asm,
            "t_bc": t_bc, "t_per_iter": t_it}


def scaled_static_dof_per_s(sample, full_n_elem, full_n_free, full_iterations):
    """Full-size CPU estimate: per-element assembly/BC cost and per-iteration CG cost scale
    linearly with the element count; the iteration count is the one the same Jacobi-PCG
    needs on the full problem (measured on the GPU run, same algorithm and tolerance)."""
    s = full_n_elem / sample["n_elem"]
    t = s * (sample["t_assemble"] + sample["t_bc"]) + full_iterations * s * sample["t_per_iter"]
    return full_n_free / t, t
