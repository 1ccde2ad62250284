"""Small line-preconditioned solve for compute-sanitizer runs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fem_calculator_b200 import _lib as L, meshgen, compat
from fem_calculator_b200.api import FrameModel
from fem_calculator_b200.sections import calculate_section_properties as csp
mesh, sec, bc = meshgen.lattice_frame_case(9, 8, 7, jitter=0.05)
es, props, _ = compat.frame_section_table(mesh, sec, csp)
fixed, f = compat.frame_bc_vectors(mesh, bc, len(mesh.points))
m = FrameModel(0)
m.set_mesh(mesh.points, mesh.cells_dict["line"], es, props, 2e11, 2e11 / 2.6)
m.assemble(); m.set_bc(fixed, f)
uj, _, sj = m.solve_static(method=L.SOLVER_PCG, precond=L.PRECOND_JACOBI)
u, _, st = m.solve_static(method=L.SOLVER_PCG, precond=L.PRECOND_LINES)
print("jacobi", sj["iterations"], "lines", st["iterations"], st["precond_used"], st["coarse_dim"], np.linalg.norm(u - uj) / np.linalg.norm(uj))
m.close()
