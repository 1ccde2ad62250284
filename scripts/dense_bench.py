"""Development timing: dense blocked Cholesky (DMMA trailing update) on lattice frames."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fem_calculator_b200 import _lib as L, meshgen, compat
from fem_calculator_b200.api import FrameModel
from fem_calculator_b200.sections import calculate_section_properties as csp
cases = [(10, 10, 10), (12, 12, 11), (14, 14, 13)] if len(sys.argv) < 4 else [tuple(int(v) for v in sys.argv[1:4])]
for dims in cases:
    mesh, sec, bc = meshgen.lattice_frame_case(*dims, jitter=0.05)
    es, props, _ = compat.frame_section_table(mesh, sec, csp)
    fixed, f = compat.frame_bc_vectors(mesh, bc, len(mesh.points))
    m = FrameModel(0)
    m.set_mesh(mesh.points, mesh.cells_dict["line"], es, props, 2e11, 2e11 / 2.6)
    m.assemble(); m.set_bc(fixed, f)
    ms, flops = m.time_kernel(7, 1, 3)
    u, r, st = m.solve_static(method=L.SOLVER_DENSE)
    up, _, stp = m.solve_static(method=L.SOLVER_PCG)
    print(f"{dims}: n = {len(f)}: fill+factor {ms:.2f} ms = {flops/ms/1e9:.2f} TFLOP/s (n^3/3); full dense solve {st['device_ms']:.2f} ms; "
          f"PCG {stp['device_ms']:.2f} ms ({stp['iterations']} its); |u_dense-u_pcg|/|u| = {np.linalg.norm(u-up)/np.linalg.norm(up):.2e}", flush=True)
    m.close()
