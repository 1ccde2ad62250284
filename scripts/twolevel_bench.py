"""Development timing script (not the bench contract): Jacobi-PCG vs the two-level PCG on the C3-size
frame — iterations, device ms, per-kernel in-loop times (every 8th iteration timed), setup cost."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fem_calculator_b200 import _lib as L, meshgen, compat
from fem_calculator_b200.api import FrameModel
from fem_calculator_b200.sections import calculate_section_properties as csp

nx, ny, nz = [int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (56, 56, 54))]
aggs = sys.argv[4:] or [""]
mesh, sec, bc = meshgen.lattice_frame_case(nx, ny, nz, jitter=0.05)
E, nu = meshgen.E_STEEL, meshgen.NU_STEEL
es, props, _ = compat.frame_section_table(mesh, sec, csp)
fixed, f = compat.frame_bc_vectors(mesh, bc, len(mesh.points))
nfree = len(f) - len(fixed)
m = FrameModel(0)
m.set_mesh(mesh.points, mesh.cells_dict["line"], es, props, E, E / (2 * (1 + nu)))
m.assemble(); m.set_bc(fixed, f)
uj, _, st = m.solve_static(method=L.SOLVER_PCG, precond=L.PRECOND_JACOBI, want_reactions=False)
uj, _, st = m.solve_static(method=L.SOLVER_PCG, precond=L.PRECOND_JACOBI, want_reactions=False)
print(f"jacobi     : {st['iterations']} its, {st['device_ms']:.1f} ms, {nfree / st['device_ms'] / 1e3:.2f} M DOF/s", flush=True)
for a in aggs:
    if a.startswith("w"):        # "w2.0": Jacobi weight omega, aggregates as before
        os.environ["FEMB_TL_OMEGA"] = a[1:]
    elif a:
        os.environ["FEMB_COARSE_AGGS"] = a
    m.set_mesh(mesh.points, mesh.cells_dict["line"], es[::-1].copy(), props, E, E / (2 * (1 + nu)))   # force a new symbolic phase
    m.set_mesh(mesh.points, mesh.cells_dict["line"], es, props, E, E / (2 * (1 + nu)))
    m.assemble(); m.set_bc(fixed, f)
    t0 = time.time()
    u, _, st0 = m.solve_static(method=L.SOLVER_PCG, precond=L.PRECOND_TWO_LEVEL, want_reactions=False)
    w0 = time.time() - t0
    m.assemble(); m.set_bc(fixed, f)       # numeric setup again, aggregate tables kept
    u, _, st1 = m.solve_static(method=L.SOLVER_PCG, precond=L.PRECOND_TWO_LEVEL, want_reactions=False)
    u, _, st2 = m.solve_static(method=L.SOLVER_PCG, precond=L.PRECOND_TWO_LEVEL, want_reactions=False)   # coarse inverse kept
    u, _, stp = m.solve_static(method=L.SOLVER_PCG, precond=L.PRECOND_TWO_LEVEL, want_reactions=False, profile=8)
    err = np.linalg.norm(u - uj) / np.linalg.norm(uj)
    print(f"two-level {a or 'default'}: coarse {st1['coarse_dim']}, {st1['iterations']} its, first call {st0['device_ms']:.1f} ms (wall {w0*1e3:.0f}), "
          f"with numeric setup {st1['device_ms']:.1f} ms, solve only {st2['device_ms']:.1f} ms = {st2['device_ms']/st2['iterations']*1e3:.1f} us/it, "
          f"{nfree / st1['device_ms'] / 1e3:.2f} M DOF/s; operator {stp['spmv_ms']/stp['spmv_timed']*1e3:.1f} us, "
          f"update+coarse {stp['update_ms']/stp['spmv_timed']*1e3:.1f} us; |u-u_jacobi|/|u| {err:.2e}", flush=True)
