"""Development timing: BASELINE config 4 (independent 2k-element cantilevers, batched chain solver)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fem_calculator_b200 import meshgen
from fem_calculator_b200.api import FrameModel
from fem_calculator_b200.sections import calculate_section_properties as csp

n_models = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
n_el = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
p = meshgen.batch_cantilever_params(n_models)
mesh, _, _ = meshgen.cantilever_case(n_el, 4.0)
xyz = mesh.points
nn = n_el + 1
props = np.array([csp("rectangular section", {"d": d, "b": b}) for d, b in zip(p["d"], p["b"])])
fixed_mask = np.zeros(6 * nn, dtype=np.uint8); fixed_mask[:6] = 1
f = np.zeros((n_models, 6 * nn))
f[:, 6 * (nn - 1) + 1] = p["tip_fy"]
f[:, 2::6] += p["nodal_fz"][:, None]
E, nu = meshgen.E_STEEL, meshgen.NU_STEEL
m = FrameModel(0)
for rep in range(3):
    t0 = time.perf_counter()
    u, st = m.batch_solve(xyz, props, E, E / (2 * (1 + nu)), fixed_mask, f)
    wall = time.perf_counter() - t0
    ndof = n_models * 6 * n_el
    print(json.dumps({"models": n_models, "elements": n_el, "free_dof": ndof, "device_ms": st["device_ms"], "wall_ms": wall * 1e3,
                      "dof_per_s_device": ndof / (st["device_ms"] * 1e-3), "dof_per_s_e2e": ndof / wall}))
# closed form check of the tip-load part on model 0 is in tests; here: finite + linearity spot check
assert np.isfinite(u).all()
m.close()
