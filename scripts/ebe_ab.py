"""Development A/B: assembled BSR operator vs the matrix-free (EBE) operator on the C3 frame."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fem_calculator_b200 import _lib as L, meshgen, compat
from fem_calculator_b200.api import FrameModel
from fem_calculator_b200.sections import calculate_section_properties as csp

nx, ny, nz = [int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (56, 56, 54))]
mesh, sec, bc = meshgen.lattice_frame_case(nx, ny, nz, jitter=0.05)
es, props, _ = compat.frame_section_table(mesh, sec, csp)
fixed, f = compat.frame_bc_vectors(mesh, bc, len(mesh.points))
m = FrameModel(0)
m.set_mesh(mesh.points, mesh.cells_dict["line"], es, props, 2e11, 2e11 / 2.6)
m.assemble(); m.set_bc(fixed, f)
for which, name in ((0, "bsr spmv"), (3, "ebe x1"), (4, "ebe x4")):
    ms, by = m.time_kernel(which, 5, 100)
    print(f"{name}: b2b {ms*1e3:.1f} us, {by/1e6:.1f} MB algorithmic -> {by/ms/1e6:.0f} GB/s")
us = {}
for op, name in ((L.OP_BSR, "bsr"), (L.OP_EBE, "ebe")):
    for rep in range(2):
        u, r, st = m.solve_static(method=L.SOLVER_PCG, rtol=1e-12, want_reactions=False, op=op)
    us[name] = u
    print(f"{name}: op_used {st['op_used']} {st['iterations']} its {st['device_ms']:.1f} ms {st['device_ms']/st['iterations']*1e3:.2f} us/it "
          f"res {st['rel_residual']:.2e} DOF/s {(len(f)-len(fixed))/(st['device_ms']/1e3):.3e}")
    u, r, st = m.solve_static(method=L.SOLVER_PCG, rtol=1e-12, want_u=False, want_reactions=False, op=op, profile=8)
    k = max(1, st["spmv_timed"])
    print(f"   in-loop ({k} timed): operator {st['spmv_ms']/k*1e3:.1f} us, update {st['update_ms']/k*1e3:.1f} us, {st['device_ms']/st['iterations']*1e3:.2f} us/it")
print("||u_ebe - u_bsr|| / ||u_bsr|| =", np.linalg.norm(us["ebe"] - us["bsr"]) / np.linalg.norm(us["bsr"]))
if os.environ.get("MODAL", "1") == "1":
    for op, name in ((L.OP_EBE, "ebe"),) + (((L.OP_BSR, "bsr"),) if os.environ.get("MODAL_BSR") else ()):
        lam, phi, st = m.modal(k=20, op=op)
        print(f"modal {name}: {json.dumps(st)}")
        print("   omega", np.sqrt(lam)[:4], "...", np.sqrt(lam)[-1])
m.close()
