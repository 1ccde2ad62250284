"""Development A/B: PCG timing of the C3 frame with the library named by FEMB_LIB (default: in-tree)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fem_calculator_b200 import _lib as L, meshgen, compat
from fem_calculator_b200.api import FrameModel
from fem_calculator_b200.sections import calculate_section_properties as csp
mesh, sec, bc = meshgen.lattice_frame_case(56, 56, 54, jitter=0.05)
es, props, _ = compat.frame_section_table(mesh, sec, csp)
fixed, f = compat.frame_bc_vectors(mesh, bc, len(mesh.points))
m = FrameModel(0)
m.set_mesh(mesh.points, mesh.cells_dict["line"], es, props, 2e11, 2e11 / 2.6)
m.assemble(); m.set_bc(fixed, f)
ms, by = m.time_kernel(0, 5, 100)
try:
    ms9, by9 = m.time_kernel(9, 5, 100)
    print(f"stream-read ceiling (K values, {by9/1e6:.0f} MB): {ms9*1e3:.1f} us = {by9/ms9/1e6:.0f} GB/s")
except Exception as e:
    print("no stream-read hook", e)
out = [f"lib={os.environ.get('FEMB_LIB','in-tree')} spmv b2b {ms*1e3:.1f} us"]
for pc, name in ((L.PRECOND_JACOBI, "jacobi"), (L.PRECOND_BLOCK_JACOBI, "blockj")):
    for rep in range(2):
        u, r, st = m.solve_static(method=L.SOLVER_PCG, precond=pc, rtol=1e-12, want_u=False, want_reactions=False)
    out.append(f"{name} {st['iterations']} its {st['device_ms']:.1f} ms {st['device_ms']/st['iterations']*1e3:.2f} us/it")
print(" | ".join(out))
u, r, st = m.solve_static(method=L.SOLVER_PCG, precond=L.PRECOND_JACOBI, rtol=1e-12, want_u=False, want_reactions=False, profile=8)
if "update_ms" in st and st["spmv_timed"]:
    k = st["spmv_timed"]
    print(f"   in-loop (every 8th iteration, {k} timed): spmv {st['spmv_ms']/k*1e3:.1f} us, update {st['update_ms']/k*1e3:.1f} us, iteration {st['device_ms']/st['iterations']*1e3:.1f} us")
