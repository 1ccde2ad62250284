"""Short ncu target: one block step of the modal solve on the C3 frame (lockstep 4-RHS PCG kernels)."""
import os, sys
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
try:
    m.modal(k=20, max_iter=1)
except L.FembError as e:
    print("expected:", e)
print(m.last_stats)
m.close()
print("profile target done")
