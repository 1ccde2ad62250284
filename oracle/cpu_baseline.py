"""TEST/BENCH INFRASTRUCTURE ONLY — times the CPU oracle (oracle/ref_sparse.py, the sparse restatement of the
reference's numpy/scipy path, + oracle/cg_omp.c, the multi-threaded restatement of its Jacobi-CG) on the host
cores.  Used only by bench.py's ``cpu_baseline`` legs and its ``--impl reference`` arm.

The reference itself (dense 6N x 6N + LAPACK, BeamSolver.py:360-418) cannot run beyond a few thousand DOF, so the
CPU baseline of record is the scalable port: vectorised element formation -> scipy COO->CSR -> Jacobi-preconditioned
CG on every host core.  Every number returned here is MEASURED on a complete solve of the stated
problem (no extrapolation): the full workload for the reference arm, a smaller lattice of the same family for the
bounded cpu_baseline leg.
"""
from __future__ import annotations

import time

import numpy as np

from . import native
from . import ref_sparse as S


def lattice_static_solve(mesh, elem_sec, props, bc_data, E, nu, rtol=1e-12):
    """The whole static path of BeamSolver.py:360-418 on the CPU: element formation + COO->CSR assembly, BC
    partition, Jacobi-CG to ``rtol`` (all host cores), reactions K u - f.  Returns dict with the timings, the
    iteration count and the results (u, reactions) for the parity check."""
    conn = mesh.cells_dict["line"]
    t0 = time.perf_counter()
    K, _ = S.frame_assemble(mesh.points, conn, elem_sec, props, E, nu)
    t_asm = time.perf_counter() - t0
    t0 = time.perf_counter()
    fixed, free, f = S.frame_bc(mesh, bc_data)
    Kff = K[free][:, free].tocsr()
    ff = f[free]
    t_bc = time.perf_counter() - t0
    t0 = time.perf_counter()
    uf, info = native.pcg_jacobi(Kff, ff, rtol=rtol)
    t_cg = time.perf_counter() - t0
    t0 = time.perf_counter()
    u = np.zeros(K.shape[0])
    u[free] = uf
    r = K @ u - f
    t_react = time.perf_counter() - t0
    return {"n_elem": len(conn), "n_dof": K.shape[0], "n_free": len(free), "nnz": int(Kff.nnz), "t_assemble": t_asm,
            "t_bc": t_bc, "t_cg": t_cg, "t_reactions": t_react, "t_total": t_asm + t_bc + t_cg + t_react,
            "iterations": info["iterations"], "rel_residual": info["rel_residual"], "flag": info["flag"],
            "threads": native.threads(), "u": u, "reactions": r, "fixed": fixed, "f": f}


def lattice_modal_solve(mesh, elem_sec, props, bc_data, E, nu, k=20):
    """The modal path (pencil of BeamSolver.py:440-455) on the CPU: assembly + scipy eigsh(sigma=0) (SuperLU
    shift-invert Lanczos).  Returns dict(n_free, seconds, eigenvalues)."""
    conn = mesh.cells_dict["line"]
    t0 = time.perf_counter()
    K, M = S.frame_assemble(mesh.points, conn, elem_sec, props, E, nu)
    fixed, free, f = S.frame_bc(mesh, bc_data)
    t_asm = time.perf_counter() - t0
    t0 = time.perf_counter()
    lam, _ = S.frame_modal(K, M, free, k=k)
    t_eig = time.perf_counter() - t0
    return {"n_free": len(free), "t_assemble": t_asm, "t_eig": t_eig, "seconds": t_asm + t_eig, "eigenvalues": lam}


def chain_batch_solve(xyz, sec_props, E, G, fixed_mask, f, n_models):
    """BASELINE config 4 on the CPU for the first ``n_models`` models: per model the chain K from the reference's
    element (ref_sparse) and scipy's banded Cholesky solve.  Returns (u (n_models, ndof), seconds)."""
    import scipy.linalg as sla
    n_nodes = len(xyz)
    conn = np.stack([np.arange(n_nodes - 1), np.arange(1, n_nodes)], axis=1)
    free = np.flatnonzero(np.asarray(fixed_mask) == 0)
    out = np.zeros((n_models, 6 * n_nodes))
    nu = E / (2 * G) - 1.0
    t0 = time.perf_counter()
    for m in range(n_models):
        K, _ = S.frame_assemble(xyz, conn, np.zeros(len(conn), dtype=np.int32), sec_props[m:m + 1], E, nu)
        Kc = K[free][:, free].tocsr()
        nb = 12                                   # lower band storage for solveh_banded: half bandwidth 11 (6-DOF chain)
        ab = np.zeros((nb, len(free)))
        for d in range(nb):
            diag = Kc.diagonal(-d)
            ab[d, :len(diag)] = diag
        out[m, free] = sla.solveh_banded(ab, f[m][free], lower=True)
    return out, time.perf_counter() - t0


def tet10_assembly(points, conn, E, nu):
    """ReactionSolver.py:115-152 on the CPU (vectorised port): returns (K csr, seconds)."""
    t0 = time.perf_counter()
    K, _ = S.tet10_assemble(points, conn, E, nu)
    return K, time.perf_counter() - t0
