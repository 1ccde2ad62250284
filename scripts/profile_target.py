"""Short ncu target: BASELINE config-3 frame, 3 assemblies + 40 PCG iterations (not a bench)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fem_calculator_b200 import _lib as L, meshgen, compat
from fem_calculator_b200.api import FrameModel
from fem_calculator_b200.sections import calculate_section_properties as csp

lat = [int(v) for v in sys.argv[1:4]] if len(sys.argv) > 3 else (56, 56, 54)
iters = int(os.environ.get("PCG_ITERS", "40"))
mesh, sec, bc = meshgen.lattice_frame_case(*lat, jitter=0.05)
es, props, _ = compat.frame_section_table(mesh, sec, csp)
fixed, f = compat.frame_bc_vectors(mesh, bc, len(mesh.points))
m = FrameModel(0)
m.set_mesh(mesh.points, mesh.cells_dict["line"], es, props, 2e11, 2e11 / 2.6)
for _ in range(3):
    m.assemble()
m.set_bc(fixed, f)
try:
    m.solve_static(method=L.SOLVER_PCG, precond=getattr(L, "PRECOND_" + os.environ.get("PCG_PRECOND", "JACOBI")), max_iter=iters, want_u=False, want_reactions=False)
except L.FembError as e:
    assert e.code == L.FEMB_ERR_NOT_CONVERGED, e
m.close()
print("profile target done")
