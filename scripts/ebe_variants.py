"""Development: time the matrix-free operator variant selected by FEMB_EBE_VARIANT / FEMB_EBE_VARIANT4 /
FEMB_EBE_CTAS (one process per variant: the library reads the environment once)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fem_calculator_b200 import _lib as L, meshgen, compat
from fem_calculator_b200.api import FrameModel
from fem_calculator_b200.sections import calculate_section_properties as csp
mesh, sec, bc = meshgen.lattice_frame_case(56, 56, 54, jitter=0.05)
es, props, _ = compat.frame_section_table(mesh, sec, csp)
fixed, f = compat.frame_bc_vectors(mesh, bc, len(mesh.points))
m = FrameModel(0)
m.set_mesh(mesh.points, mesh.cells_dict["line"], es, props, 2e11, 2e11 / 2.6)
m.assemble(); m.set_bc(fixed, f)
tag = f"linked={os.environ.get('FEMB_PCG_LINKED','1')}"
ms1, _ = m.time_kernel(3, 5, 100)
ms4, _ = m.time_kernel(4, 5, 100)
ms5, _ = m.time_kernel(5, 5, 100)
rng = np.random.default_rng(1)
x = rng.standard_normal(len(f)); x[fixed] = 0
yb, _ = m.apply_k(x, op=L.OP_BSR, masked=True)
ye, _ = m.apply_k(x, op=L.OP_EBE, masked=True)
err = np.abs(ye - yb).max() / np.abs(yb).max()
u, r, st = m.solve_static(method=L.SOLVER_PCG, rtol=1e-12, want_u=False, want_reactions=False, op=L.OP_EBE)
u, r, st = m.solve_static(method=L.SOLVER_PCG, rtol=1e-12, want_u=False, want_reactions=False, op=L.OP_EBE, profile=8)
k = max(1, st["spmv_timed"])
print(f"{tag}: x1 b2b {ms1*1e3:.1f} us (with dot {ms5*1e3:.1f}) | x4 b2b {ms4*1e3:.1f} us | err {err:.1e} | pcg {st['iterations']} its {st['device_ms']/st['iterations']*1e3:.1f} us/it "
      f"(op {st['spmv_ms']/k*1e3:.1f}, upd {st['update_ms']/k*1e3:.1f})", flush=True)
m.close()
