"""Development timing: Tet10 path (ReactionSolver.py) — fused element+assembly and PCG solve on box meshes."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fem_calculator_b200 import _lib as L, meshgen, compat
from fem_calculator_b200.api import Tet10Model
cases = [(8, 2, 8), (32, 8, 32), (64, 16, 64)] if len(sys.argv) < 4 else [tuple(int(v) for v in sys.argv[1:4])]
for dims in cases:
    t0 = time.time()
    mesh, fd, xd = meshgen.tet10_box_case(*dims)
    conn = mesh.cells_dict["tetra10"]
    m = Tet10Model(0)
    m.set_mesh(mesh.points, conn, 2e11, 0.3)
    t1 = time.time(); m.assemble(); t_first = time.time() - t1
    fa = compat.ForceAnalysisB200(mesh, fd, xd, 2e11, 0.3)
    fa.apply_boundary_conditions_only = True
    ms, by = m.time_kernel(1, 2, 10)
    ndof = 3 * len(mesh.points)
    # BC: fix the x = 0 face, load the opposite face
    pts = mesh.points
    fixed = np.flatnonzero(np.repeat(pts[:, 0] <= pts[:, 0].min() + 1e-12, 3)).astype(np.int64)
    f = np.zeros(ndof); f[3 * np.flatnonzero(pts[:, 0] >= pts[:, 0].max() - 1e-12) + 1] = -10.0
    m.set_bc(fixed, f)
    ms_spmv, by_spmv = m.time_kernel(0, 3, 20)
    u, r, st = m.solve_static(method=L.SOLVER_PCG, rtol=1e-10, want_u=False, want_reactions=False)
    print(f"{dims}: {len(conn)} tets, {ndof} DOF | assembly {ms:.3f} ms = {len(conn)/ms/1e3:.2f} M tets/s, {by/ms/1e6:.0f} GB/s "
          f"(first call incl. symbolic {t_first:.2f} s) | spmv {ms_spmv*1e3:.1f} us = {by_spmv/ms_spmv/1e6:.0f} GB/s | "
          f"PCG {st['iterations']} its {st['device_ms']:.1f} ms -> {(ndof-len(fixed))/st['device_ms']/1e3:.2f} M DOF/s", flush=True)
    m.close()
