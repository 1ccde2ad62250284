"""GPU parity tests of the line-preconditioned PCG (csrc/lines.cu: Jacobi + exact solves along member lines +
bundle coarse space) against the CPU oracle: same answer as the direct solve within north_star's static
tolerance (1e-10 on ||u||, 1e-9 on reactions), far fewer iterations than Jacobi, bit-reproducible; and at the
production configuration (AUTO on a frame with >= 50,000 nodes, 768 bundles per family) against the committed
sampled goldens the oracle produced with its own Jacobi-CG (oracle/make_golden_large.py)."""
import os

import numpy as np
import pytest

from fem_calculator_b200 import _lib as L
from fem_calculator_b200 import compat, meshgen
from fem_calculator_b200.api import FrameModel
from fem_calculator_b200.sections import calculate_section_properties as csp
from oracle import ref_sparse as S

pytestmark = pytest.mark.gpu

E, NU = meshgen.E_STEEL, meshgen.NU_STEEL
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _setup(nx, ny, nz, jitter, bc=None):
    mesh, sec, bc0 = meshgen.lattice_frame_case(nx, ny, nz, jitter=jitter)
    bc = bc or bc0
    es, props, _ = compat.frame_section_table(mesh, sec, csp)
    fixed, f = compat.frame_bc_vectors(mesh, bc, len(mesh.points))
    m = FrameModel(0)
    m.set_mesh(mesh.points, mesh.cells_dict["line"], es, props, E, E / (2 * (1 + NU)))
    m.assemble()
    m.set_bc(fixed, f)
    return mesh, bc, es, props, fixed, f, m


@pytest.mark.parametrize("jitter,bundles", [(0.05, None), (0.0, None), (0.05, "1"), (0.05, "7"), (0.05, "40")])
def test_lines_pcg_matches_oracle(jitter, bundles, monkeypatch):
    """bundles = FEMB_LINE_BUNDLES override (bundles per family): default (one line per bundle on this small
    frame), a single bundle per family, counts that leave ragged bundles / a padded coarse dimension."""
    if bundles:
        monkeypatch.setenv("FEMB_LINE_BUNDLES", bundles)
    mesh, bc, es, props, fixed, f, m = _setup(14, 12, 11, jitter)
    Ko, _ = S.frame_assemble(mesh.points, mesh.cells_dict["line"], es, props, E, NU)
    _, free, _ = S.frame_bc(mesh, bc)
    uo, _ = S.solve_static(Ko, f, fixed, free, method="direct")
    uj, _, stj = m.solve_static(method=L.SOLVER_PCG, precond=L.PRECOND_JACOBI)
    u, r, st = m.solve_static(method=L.SOLVER_PCG, precond=L.PRECOND_LINES)
    assert st["converged"] == 1 and st["op_used"] == L.OP_EBE and st["precond_used"] == L.PRECOND_LINES, st
    per_family = [12 * 11, 14 * 11, 14 * 12]
    want = sum(min(n, int(bundles)) if bundles else n for n in per_family)
    assert st["coarse_dim"] == want, st
    assert np.linalg.norm(u - uo) <= 1e-10 * np.linalg.norm(uo), st
    assert np.linalg.norm(r - (Ko @ uo - f)) <= 1e-9 * np.linalg.norm(f)
    assert (u[fixed] == 0).all()
    assert st["iterations"] < 0.25 * stj["iterations"], (st["iterations"], stj["iterations"])
    u2, _, st2 = m.solve_static(method=L.SOLVER_PCG, precond=L.PRECOND_LINES)
    assert np.array_equal(u, u2) and st2["iterations"] == st["iterations"], "not run-to-run reproducible"
    m.close()


def test_lines_follow_new_bc_and_coordinates():
    """Line tables depend on the topology only; directions, factors and inverses are rebuilt when the
    coordinates, K or the BC mask change (pinned supports: translations only, loads in y)."""
    mesh, bc, es, props, fixed, f, m = _setup(10, 9, 8, 0.05)
    m.solve_static(method=L.SOLVER_PCG, precond=L.PRECOND_LINES)
    fixed2 = np.array([d for d in fixed if d % 6 < 3], dtype=np.int64)
    f2 = f.copy()
    f2.reshape(-1, 6)[:, 1] += 50.0
    f2[fixed2] = 0.0
    pts2 = mesh.points * np.array([1.0, 1.3, 0.8])
    m.set_mesh(pts2, mesh.cells_dict["line"], es, props, E, E / (2 * (1 + NU)))
    m.assemble()
    m.set_bc(fixed2, f2)
    u2, r2, st2 = m.solve_static(method=L.SOLVER_PCG, precond=L.PRECOND_LINES)
    Ko, _ = S.frame_assemble(pts2, mesh.cells_dict["line"], es, props, E, NU)
    free2 = np.setdiff1d(np.arange(len(f2)), fixed2)
    uo, _ = S.solve_static(Ko, f2, fixed2, free2, method="direct")
    assert st2["precond_used"] == L.PRECOND_LINES and st2["converged"] == 1
    assert np.linalg.norm(u2 - uo) <= 1e-10 * np.linalg.norm(uo), st2
    m.close()


def test_lines_zero_load_iteration_cap_and_fallbacks():
    mesh, bc, es, props, fixed, f, m = _setup(8, 8, 8, 0.05)
    m.set_bc(fixed, np.zeros_like(f))
    u, _, st = m.solve_static(method=L.SOLVER_PCG, precond=L.PRECOND_LINES)
    assert st["converged"] == 1 and st["iterations"] == 0 and not u.any()
    m.set_bc(fixed, f)
    with pytest.raises(L.FembError) as ei:
        m.solve_static(method=L.SOLVER_PCG, precond=L.PRECOND_LINES, max_iter=7)
    assert ei.value.code == -3 and m.last_stats["iterations"] == 7
    # BSR operator requested -> the Jacobi path runs and says so
    u, _, st = m.solve_static(method=L.SOLVER_PCG, precond=L.PRECOND_LINES, op=L.OP_BSR)
    assert st["converged"] == 1 and st["coarse_dim"] == 0 and st["op_used"] == L.OP_BSR
    m.close()


def test_lines_on_a_frame_without_lines_falls_back():
    """A zig-zag truss whose consecutive members are never collinear has no member lines: asking for LINES
    runs the Jacobi path (precond_used says so) and still matches the oracle."""
    n = 16
    x = np.arange(n, dtype=float)
    pts = np.stack([x, (np.arange(n) % 2) * 1.0, 0.3 * ((np.arange(n) // 2) % 2)], axis=1)
    conn = np.stack([np.arange(n - 1), np.arange(1, n)], axis=1)
    conn = np.concatenate([conn, np.stack([np.arange(n - 2), np.arange(2, n)], axis=1)[::3]])   # bracing: not a chain
    props = np.asarray([csp("rectangular section", {"d": 0.1, "b": 0.05}, False)])
    es = np.zeros(len(conn), dtype=np.int32)
    fixed = np.arange(6, dtype=np.int64)
    f = np.zeros(6 * n)
    f[6 * (n - 1) + 1] = -100.0
    m = FrameModel(0)
    m.set_mesh(pts, conn, es, props, E, E / (2 * (1 + NU)))
    m.assemble()
    m.set_bc(fixed, f)
    u, _, st = m.solve_static(method=L.SOLVER_PCG, precond=L.PRECOND_LINES, max_iter=100000)
    Ko, _ = S.frame_assemble(pts, conn, es, props, E, NU)
    free = np.setdiff1d(np.arange(6 * n), fixed)
    uo, _ = S.solve_static(Ko, f, fixed, free, method="direct")
    assert st["converged"] == 1
    assert st["precond_used"] != L.PRECOND_LINES and st["coarse_dim"] == 0, st
    assert np.linalg.norm(u - uo) <= 1e-8 * np.linalg.norm(uo), st     # cond(K) of a 16-node cantilever truss: ~1e7
    m.close()


def test_modal_with_line_preconditioned_inner_solves_matches_oracle():
    """Lowest modes of K phi = lambda M phi (BeamSolver.py:440-455) with the line-preconditioned PCG behind the
    shift-invert operator: eigenvalues within 1e-8 of the oracle's, shapes M-orthonormal."""
    mesh, bc, es, props, fixed, f, m = _setup(10, 9, 8, 0.05)
    Ko, Mo = S.frame_assemble(mesh.points, mesh.cells_dict["line"], es, props, E, NU)
    _, free, _ = S.frame_bc(mesh, bc)
    lam_o, _ = S.frame_modal(Ko, Mo, free, k=8)
    lam, phi, st = m.modal(k=8, precond=L.PRECOND_LINES)
    assert st["precond_used"] == L.PRECOND_LINES and st["coarse_dim"] > 0, st
    assert len(lam) == 8
    assert np.abs(lam - lam_o[:8]).max() <= 1e-8 * np.abs(lam_o[:8]).max(), (lam, lam_o[:8])
    G = phi.T @ (Mo @ phi)
    assert np.abs(G - np.eye(8)).max() <= 1e-8
    m.close()


def _check_against_sampled_golden(lat, want_precond):
    g = np.load(os.path.join(GOLD, f"lattice_{lat[0]}x{lat[1]}x{lat[2]}_sampled.npz"))
    mesh, bc, es, props, fixed, f, m = _setup(*lat, float(g["jitter"]))
    assert len(f) == int(g["n_dof"]) and len(fixed) == int(g["n_fixed"])
    u, r, st = m.solve_static(method=L.SOLVER_PCG, rtol=1e-13)             # AUTO preconditioner, default bundle target
    m.close()
    assert st["converged"] == 1 and st["precond_used"] == want_precond, st
    idx = g["idx"]
    # ||u - u_ref|| / ||u_ref|| on the 4,096 sampled DOFs and on the norm (north_star: 1e-10 on ||u||)
    assert np.linalg.norm(u[idx] - g["u_idx"]) <= 1e-10 * np.linalg.norm(g["u_idx"]), st
    assert abs(np.linalg.norm(u) - float(g["u_norm"])) <= 1e-10 * float(g["u_norm"])
    assert abs(np.abs(u).max() - float(g["u_absmax"])) <= 1e-10 * float(g["u_absmax"])
    rs = r[fixed].reshape(-1, 6)[:, :3].sum(axis=0)
    assert np.abs(rs - g["reaction_sum"]).max() <= 1e-9 * float(g["f_norm"])
    assert abs(np.linalg.norm(r[fixed]) - float(g["reaction_norm"])) <= 1e-9 * float(g["reaction_norm"])
    return st


def test_production_configuration_40x40x38_matches_oracle_golden():
    """60,800 nodes: AUTO picks the line preconditioner with its production parameters (768 bundles per family)."""
    st = _check_against_sampled_golden((40, 40, 38), L.PRECOND_LINES)
    assert st["coarse_dim"] == 3 * 768 and st["iterations"] < 400, st


def test_baseline_config3_full_size_matches_oracle_golden():
    """BASELINE configs[2] at full size (1,016,064 DOF) against the oracle's sampled golden."""
    st = _check_against_sampled_golden((56, 56, 54), L.PRECOND_LINES)
    assert st["coarse_dim"] == 3 * 768 and st["iterations"] < 450, st
