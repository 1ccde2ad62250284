"""GPU parity tests of the two forms of the frame operator y = K x behind the Krylov loops:
the assembled block-CSR SpMV and the matrix-free (element-by-element) kernel, which rebuilds
R^T k R (BeamSolver.py:386-387, 646-660) per (node, element end) pair instead of reading K."""
import numpy as np
import pytest

from fem_calculator_b200 import _lib as L
from fem_calculator_b200 import compat, meshgen
from fem_calculator_b200.api import FrameModel
from fem_calculator_b200.sections import calculate_section_properties as csp
from oracle import ref_sparse as S

pytestmark = pytest.mark.gpu
E, NU = meshgen.E_STEEL, meshgen.NU_STEEL


def _setup(nx, ny, nz, jitter):
    mesh, sec, bc = meshgen.lattice_frame_case(nx, ny, nz, jitter=jitter)
    es, props, _ = compat.frame_section_table(mesh, sec, csp)
    fixed, f = compat.frame_bc_vectors(mesh, bc, len(mesh.points))
    m = FrameModel(0)
    m.set_mesh(mesh.points, mesh.cells_dict["line"], es, props, E, E / (2 * (1 + NU)))
    m.assemble()
    m.set_bc(fixed, f)
    return mesh, bc, es, props, fixed, f, m


@pytest.mark.parametrize("jitter", [0.0, 0.05])     # 0.0: axis-aligned, vertical-member branch (BeamSolver.py:380)
def test_matrix_free_product_matches_oracle_matrix(jitter):
    mesh, bc, es, props, fixed, f, m = _setup(9, 8, 7, jitter)
    Ko, _ = S.frame_assemble(mesh.points, mesh.cells_dict["line"], es, props, E, NU)
    rng = np.random.default_rng(20261018)
    x = rng.standard_normal(len(f))
    x.reshape(-1, 6)[:, 3:] *= 10.0                  # rotations and translations both matter
    yo = Ko @ x
    scale = np.abs(Ko).dot(np.abs(x))                # entry-wise rounding scale of the product
    for op in (L.OP_BSR, L.OP_EBE):
        y, used = m.apply_k(x, op=op)
        assert used == op
        assert (np.abs(y - yo) <= 1e-14 * scale).all(), (op, float((np.abs(y - yo) / scale).max()))
        y2, _ = m.apply_k(x, op=op)
        assert np.array_equal(y, y2), "operator is not run-to-run reproducible"
    # masked form: identity rows on the fixed DOFs, K_ff elsewhere (x zero on the fixed DOFs)
    xm = x.copy(); xm[fixed] = 0.0
    ym_o = Ko @ xm; ym_o[fixed] = 0.0
    for op in (L.OP_BSR, L.OP_EBE):
        ym, _ = m.apply_k(xm, op=op, masked=True)
        assert (np.abs(ym - ym_o) <= 1e-14 * scale).all()
        assert np.all(ym[fixed] == 0.0)
    m.close()


def test_pcg_with_either_operator_matches_oracle():
    mesh, bc, es, props, fixed, f, m = _setup(12, 10, 9, 0.05)
    Ko, _ = S.frame_assemble(mesh.points, mesh.cells_dict["line"], es, props, E, NU)
    _, free, _ = S.frame_bc(mesh, bc)
    uo, _ = S.solve_static(Ko, f, fixed, free, method="direct")
    its = {}
    for op in (L.OP_BSR, L.OP_EBE):
        u, r, st = m.solve_static(method=L.SOLVER_PCG, op=op)
        assert st["op_used"] == op and st["converged"] == 1
        assert np.linalg.norm(u - uo) <= 1e-10 * np.linalg.norm(uo), (op, st)
        assert np.linalg.norm(r - (Ko @ uo - f)) <= 1e-9 * np.linalg.norm(f)
        u2, _, _ = m.solve_static(method=L.SOLVER_PCG, op=op)
        assert np.array_equal(u, u2)
        its[op] = st["iterations"]
    assert abs(its[L.OP_BSR] - its[L.OP_EBE]) <= max(5, its[L.OP_BSR] // 50)   # same Krylov trajectory up to rounding
    # AUTO picks the matrix-free operator for a frame with unique members on one GPU
    _, _, st = m.solve_static(method=L.SOLVER_PCG)
    assert st["op_used"] == L.OP_EBE
    m.close()


def test_linked_pcg_edge_cases():
    """Zero load (u = 0 without iterating), iteration cap, unpreconditioned run and a singular system
    through the linked-reduction PCG of the matrix-free operator."""
    mesh, bc, es, props, fixed, f, m = _setup(6, 5, 5, 0.05)
    m.set_bc(fixed, np.zeros_like(f))
    u, r, st = m.solve_static(method=L.SOLVER_PCG, op=L.OP_EBE)
    assert st["converged"] == 1 and st["iterations"] == 0 and not u.any()
    m.set_bc(fixed, f)
    with pytest.raises(L.FembError) as ei:
        m.solve_static(method=L.SOLVER_PCG, op=L.OP_EBE, max_iter=7)
    assert ei.value.code == L.FEMB_ERR_NOT_CONVERGED and m.last_stats["iterations"] == 7
    u0, _, st0 = m.solve_static(method=L.SOLVER_PCG, op=L.OP_EBE, precond=L.PRECOND_NONE)
    u1, _, st1 = m.solve_static(method=L.SOLVER_PCG, op=L.OP_BSR, precond=L.PRECOND_NONE)
    assert st0["converged"] == 1 and np.linalg.norm(u0 - u1) <= 1e-9 * np.linalg.norm(u1)
    assert abs(st0["iterations"] - st1["iterations"]) <= max(5, st1["iterations"] // 50)
    m.set_bc(np.zeros(0, dtype=np.int64), f)          # no supports: K_ff singular
    with pytest.raises(L.FembError):
        m.solve_static(method=L.SOLVER_PCG, op=L.OP_EBE, max_iter=3000)
    m.close()


def test_duplicate_members_fall_back_to_assembled_operator():
    """Two members between the same nodes: the pair view does not apply, AUTO keeps the BSR operator
    and asking for EBE explicitly is an argument error, not a wrong answer."""
    pts = np.array([[0, 0, 0], [1.0, 0.2, 0.1], [2.0, 0.5, -0.3], [2.5, 1.5, 0.3]])
    conn = np.array([[0, 1], [1, 2], [1, 0], [2, 3]])
    props = np.array([[5e-3, 4e-6, 1e-6, 2.8e-6, 0.83, 0.83, 0.02, 0.05], [2e-3, 1e-6, 2e-6, 1e-6, 0.5, 0.6, 0.02, 0.05]])
    es = np.array([0, 1, 1, 0], dtype=np.int32)
    m = FrameModel(0)
    m.set_mesh(pts, conn, es, props, E, E / 2.6)
    m.assemble()
    f = np.zeros(24); f[6 * 3 + 1] = -100.0
    fixed = np.arange(6, dtype=np.int64)
    m.set_bc(fixed, f)
    u, r, st = m.solve_static(method=L.SOLVER_PCG)
    assert st["op_used"] == L.OP_BSR
    Ko, _ = S.frame_assemble(pts, conn, es, props, E, 0.3)
    free = np.setdiff1d(np.arange(24), fixed)
    uo, _ = S.solve_static(Ko, f, fixed, free, method="direct")
    assert np.linalg.norm(u - uo) <= 1e-10 * np.linalg.norm(uo)
    with pytest.raises(L.FembError):
        m.apply_k(np.ones(24), op=L.OP_EBE)
    m.close()


def test_modal_with_either_operator_agrees():
    """The 4-vector matrix-free operator inside the lockstep PCG of the shift-invert modal solve."""
    mesh, bc, es, props, fixed, f, m = _setup(9, 8, 8, 0.05)      # > 2048 DOF: PCG-based K^-1
    lam = {}
    for op in (L.OP_BSR, L.OP_EBE):
        lam[op], phi, st = m.modal(k=8, op=op)
        assert st["op_used"] == op and len(lam[op]) == 8
    m.close()
    assert np.abs(lam[L.OP_EBE] - lam[L.OP_BSR]).max() <= 1e-8 * lam[L.OP_BSR].max()
    Ko, Mo = S.frame_assemble(mesh.points, mesh.cells_dict["line"], es, props, E, NU)
    _, free, _ = S.frame_bc(mesh, bc)
    lo, _ = S.frame_modal(Ko, Mo, free, k=8)
    assert np.abs(lam[L.OP_EBE] - lo[:8]).max() <= 1e-8 * lo[:8].max()
