"""Development timing script: lowest-k modal solve on a lattice frame (not the bench contract)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fem_calculator_b200 import _lib as L, meshgen, compat
from fem_calculator_b200.api import FrameModel
from fem_calculator_b200.sections import calculate_section_properties as csp

nx, ny, nz = [int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (56, 56, 54))]
k = int(os.environ.get("K_MODES", "20"))
block = int(os.environ.get("BLOCK", "0"))
mesh, sec, bc = meshgen.lattice_frame_case(nx, ny, nz, jitter=0.05)
es, props, _ = compat.frame_section_table(mesh, sec, csp)
fixed, f = compat.frame_bc_vectors(mesh, bc, len(mesh.points))
m = FrameModel(0)
m.set_mesh(mesh.points, mesh.cells_dict["line"], es, props, 2e11, 2e11 / 2.6)
m.assemble()
m.set_bc(fixed, f)
t0 = time.time()
lam, phi, st = m.modal(k=k, block=block)
print("modal", nx, ny, nz, "ndof", len(f), json.dumps(st), "wall", time.time() - t0)
print("omega[rad/s]", np.sqrt(lam)[:8], "...", np.sqrt(lam)[-1])
m.close()
