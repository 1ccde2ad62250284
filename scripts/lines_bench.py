"""Development timing script (not the bench contract): Jacobi / rigid-body two-level / line-preconditioned PCG on a
lattice frame — iterations, device ms, per-kernel in-loop times, setup cost.
    python scripts/lines_bench.py 56 56 54 [bundles_per_family ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fem_calculator_b200 import _lib as L, meshgen, compat
from fem_calculator_b200.api import FrameModel
from fem_calculator_b200.sections import calculate_section_properties as csp

nx, ny, nz = [int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (56, 56, 54))]
variants = sys.argv[4:] or [""]
mesh, sec, bc = meshgen.lattice_frame_case(nx, ny, nz, jitter=0.05)
E, nu = meshgen.E_STEEL, meshgen.NU_STEEL
es, props, _ = compat.frame_section_table(mesh, sec, csp)
fixed, f = compat.frame_bc_vectors(mesh, bc, len(mesh.points))
nfree = len(f) - len(fixed)
m = FrameModel(0)
m.set_mesh(mesh.points, mesh.cells_dict["line"], es, props, E, E / (2 * (1 + nu)))
m.assemble(); m.set_bc(fixed, f)
ref = None
if not os.environ.get("SKIP_JACOBI"):
    for _ in range(2):
        ref, _, st = m.solve_static(method=L.SOLVER_PCG, precond=L.PRECOND_JACOBI, want_reactions=False)
    print(f"jacobi     : {st['iterations']} its, {st['device_ms']:.1f} ms, {nfree / st['device_ms'] / 1e3:.2f} M DOF/s", flush=True)
    for _ in range(2):
        u, _, st = m.solve_static(method=L.SOLVER_PCG, precond=L.PRECOND_TWO_LEVEL, want_reactions=False)
    print(f"rigid-body : {st['iterations']} its, {st['device_ms']:.1f} ms, {nfree / st['device_ms'] / 1e3:.2f} M DOF/s", flush=True)
for a in variants:
    if a.startswith("w"):
        os.environ["FEMB_LN_OMEGA"] = a[1:]
    elif a:
        os.environ["FEMB_LINE_BUNDLES"] = a
    m.set_mesh(mesh.points, mesh.cells_dict["line"], es[::-1].copy(), props, E, E / (2 * (1 + nu)))   # force a new symbolic phase
    m.set_mesh(mesh.points, mesh.cells_dict["line"], es, props, E, E / (2 * (1 + nu)))
    m.assemble(); m.set_bc(fixed, f)
    t0 = time.time()
    u, _, st0 = m.solve_static(method=L.SOLVER_PCG, precond=L.PRECOND_LINES, want_reactions=False)
    w0 = time.time() - t0
    m.assemble(); m.set_bc(fixed, f)       # numeric setup again, line tables kept
    u, _, st1 = m.solve_static(method=L.SOLVER_PCG, precond=L.PRECOND_LINES, want_reactions=False)
    u, _, st2 = m.solve_static(method=L.SOLVER_PCG, precond=L.PRECOND_LINES, want_reactions=False)   # factors kept
    u, _, stp = m.solve_static(method=L.SOLVER_PCG, precond=L.PRECOND_LINES, want_reactions=False, profile=8)
    err = np.linalg.norm(u - ref) / np.linalg.norm(ref) if ref is not None else float("nan")
    print(f"lines {a or 'default'}: precond {st1['precond_used']} coarse {st1['coarse_dim']}, {st1['iterations']} its, first call {st0['device_ms']:.1f} ms "
          f"(wall {w0*1e3:.0f}), with numeric setup {st1['device_ms']:.1f} ms, solve only {st2['device_ms']:.1f} ms = "
          f"{st2['device_ms']/max(1,st2['iterations'])*1e3:.1f} us/it, {nfree / st1['device_ms'] / 1e3:.2f} M DOF/s; operator "
          f"{stp['spmv_ms']/max(1,stp['spmv_timed'])*1e3:.1f} us, rest {stp['update_ms']/max(1,stp['spmv_timed'])*1e3:.1f} us; |u-u_jacobi|/|u| {err:.2e}", flush=True)
m.close()
