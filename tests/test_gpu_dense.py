"""GPU parity tests of the dense blocked Cholesky (DMMA trailing update) behind FEMB_SOLVER_DENSE:
the direct counterpart of np.linalg.solve(k_ff, f_f) at BeamSolver.py:417 for reduced systems."""
import numpy as np
import pytest

from fem_calculator_b200 import _lib as L
from fem_calculator_b200 import compat, meshgen
from fem_calculator_b200.api import FrameModel
from fem_calculator_b200.sections import calculate_section_properties as csp
from oracle import ref_sparse as S

pytestmark = pytest.mark.gpu
E, NU = meshgen.E_STEEL, meshgen.NU_STEEL


@pytest.mark.parametrize("dims", [(3, 3, 2), (7, 6, 5), (8, 8, 7), (12, 10, 9)])   # 108 / 1260 / 2688 / 6480 DOF
def test_dense_cholesky_matches_oracle(dims):
    mesh, sec, bc = meshgen.lattice_frame_case(*dims, jitter=0.05)
    es, props, _ = compat.frame_section_table(mesh, sec, csp)
    fixed, f = compat.frame_bc_vectors(mesh, bc, len(mesh.points))
    m = FrameModel(0)
    m.set_mesh(mesh.points, mesh.cells_dict["line"], es, props, E, E / (2 * (1 + NU)))
    m.assemble()
    m.set_bc(fixed, f)
    u, r, st = m.solve_static(method=L.SOLVER_DENSE)
    u2, _, _ = m.solve_static(method=L.SOLVER_DENSE)
    m.close()
    assert st["method_used"] == L.SOLVER_DENSE
    assert np.array_equal(u, u2), "dense solve is not run-to-run reproducible"
    Ko, _ = S.frame_assemble(mesh.points, mesh.cells_dict["line"], es, props, E, NU)
    _, free, _ = S.frame_bc(mesh, bc)
    uo, _ = S.solve_static(Ko, f, fixed, free, method="direct")
    assert np.linalg.norm(u - uo) <= 1e-10 * np.linalg.norm(uo), (dims, np.linalg.norm(u - uo) / np.linalg.norm(uo))
    assert np.all(u[fixed] == 0.0)
    assert np.linalg.norm(r - (Ko @ uo - f)) <= 1e-9 * np.linalg.norm(f)


def test_dense_cholesky_reports_singular_and_size_limit():
    mesh, sec, bc = meshgen.lattice_frame_case(4, 4, 3, jitter=0.05)
    es, props, _ = compat.frame_section_table(mesh, sec, csp)
    m = FrameModel(0)
    m.set_mesh(mesh.points, mesh.cells_dict["line"], es, props, E, E / 2.6)
    m.assemble()
    f = np.zeros(6 * len(mesh.points)); f[7] = 1.0
    m.set_bc(np.zeros(0, dtype=np.int64), f)        # no supports: K_ff singular
    with pytest.raises(L.FembError) as ei:
        m.solve_static(method=L.SOLVER_DENSE)
    assert ei.value.code == L.FEMB_ERR_SINGULAR
    m.close()
    mesh, sec, bc = meshgen.lattice_frame_case(15, 15, 14, jitter=0.05)      # 18,900 DOF > 16,384
    es, props, _ = compat.frame_section_table(mesh, sec, csp)
    fixed, f = compat.frame_bc_vectors(mesh, bc, len(mesh.points))
    m = FrameModel(0)
    m.set_mesh(mesh.points, mesh.cells_dict["line"], es, props, E, E / 2.6)
    m.assemble(); m.set_bc(fixed, f)
    with pytest.raises(L.FembError) as ei:
        m.solve_static(method=L.SOLVER_DENSE)
    assert ei.value.code == L.FEMB_ERR_ARG
    m.close()
