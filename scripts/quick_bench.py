"""Development timing script (not the bench contract): C3-size frame, kernel timings."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fem_calculator_b200 import _lib as L, meshgen, compat
from fem_calculator_b200.api import FrameModel
from fem_calculator_b200.sections import calculate_section_properties as csp

nx, ny, nz = [int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (56, 56, 54))]
t0 = time.time()
mesh, sec, bc = meshgen.lattice_frame_case(nx, ny, nz, jitter=0.05)
E, nu = 2e11, 0.3
es, props, _ = compat.frame_section_table(mesh, sec, csp)
fixed, f = compat.frame_bc_vectors(mesh, bc, len(mesh.points))
print("meshgen s", time.time() - t0, "ndof", 6 * len(mesh.points), "elems", len(es))
m = FrameModel(0)
t0 = time.time(); m.set_mesh(mesh.points, mesh.cells_dict["line"], es, props, E, E / 2.6); print("set_mesh s", time.time() - t0)
t0 = time.time(); m.assemble(); print("first assemble (symbolic+numeric) s", time.time() - t0)
t0 = time.time(); m.assemble(); print("second assemble s", time.time() - t0)
m.set_bc(fixed, f)
for bulk in ("1", "0"):
    os.environ["FEMB_ASM_BULK"] = bulk
    ms, by = m.time_kernel(1, 3, 20)
    print(f"assembly bulk={bulk}: {ms:.4f} ms  {by/1e6:.1f} MB  {by/ms/1e6:.1f} GB/s  {len(es)/ms/1e3:.1f} M elem/s")
ms, by = m.time_kernel(0, 3, 50)
print(f"spmv: {ms:.4f} ms  {by/1e6:.1f} MB  {by/ms/1e6:.1f} GB/s")
for pc, name in ((L.PRECOND_JACOBI, "jacobi"), (L.PRECOND_BLOCK_JACOBI, "block-jacobi")):
    t0 = time.time()
    u, r, st = m.solve_static(method=L.SOLVER_PCG, precond=pc, rtol=1e-12, want_u=False, want_reactions=False)
    print(name, json.dumps(st), "wall", time.time() - t0, "DOF/s", (len(f) - len(fixed)) / (st["device_ms"] / 1e3))
u, r, st = m.solve_static(method=L.SOLVER_PCG, rtol=1e-12, profile=True, want_u=False, want_reactions=False)
print("profiled", json.dumps(st), "spmv share", st["spmv_ms"] / st["device_ms"], "spmv avg ms", st["spmv_ms"] / st["spmv_launches"])
