"""Development: host-side time breakdown of one reference-shaped run_simulation call (FEMB_TRACE=1)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["FEMB_TRACE"] = "1"
from fem_calculator_b200 import _lib as L, meshgen, compat
mesh, sec, bc = meshgen.lattice_frame_case(56, 56, 54, jitter=0.05)
w = compat.BeamAnalysisB200(mesh, sec, bc, 2e11, 0.3)
for i in range(3):
    t0 = time.perf_counter()
    w.run_simulation(k_modes=0, solver=L.SOLVER_PCG, rtol=1e-12)
    print(f"== call {i}: {(time.perf_counter()-t0)*1e3:.1f} ms, device {w.solve_stats['device_ms']:.1f} ms\n", flush=True)
