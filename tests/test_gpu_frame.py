"""GPU parity tests for the frame path (through the C ABI) against the golden vectors of
the unmodified reference and against the CPU oracle on seeded inputs."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

import golden_util as G
from fem_calculator_b200 import _lib as L
from fem_calculator_b200 import compat, meshgen
from fem_calculator_b200.api import FrameModel
from fem_calculator_b200.sections import calculate_section_properties as csp
from oracle import ref_sparse as S

pytestmark = pytest.mark.gpu


def _model_from_case(c):
    es, props = G.elem_sec_and_props(c)
    m = FrameModel(0)
    E, nu = c["E"], c["nu"]
    m.set_mesh(c["mesh"].points, c["mesh"].cells_dict["line"], es, props, E, E / (2 * (1 + nu)))
    return m


def test_element_matrices_match_reference():
    e = G.load_elements()
    E, nu = float(e["E"]), float(e["nu"])
    m = FrameModel(0)
    m.set_mesh(e["points"], e["line"], e["elem_sec"], e["props"], E, E / (2 * (1 + nu)))
    ke, me = m.elements()
    m.close()
    for a, b in ((ke, e["ref_ke_global"]), (me, e["ref_me_global"])):
        scale = np.abs(b).max(axis=(1, 2), keepdims=True)
        assert (np.abs(a - b) <= 1e-14 * scale).all(), float((np.abs(a - b) / scale).max())


def test_reference_helper_signatures():
    """get_timoshenko_stiffness_matrix / get_lumped_mass_matrix keep their signatures
    (BeamSolver.py:646,662) and values (vs the reference's own local matrices)."""
    e = G.load_elements()
    E, nu = float(e["E"]), float(e["nu"])
    w = compat.BeamAnalysisB200()
    L0, _ = S.frame_rotation(e["points"], e["line"])
    for i in (0, 4, 7):
        A, Ix, Iy, J, ky, kz = e["props"][i, :6]
        k = w.get_timoshenko_stiffness_matrix(L0[i], E, E / (2 * (1 + nu)), A, Ix, Iy, J, ky, kz)
        ref = e["ref_k_local"][i]
        assert np.abs(k - ref).max() <= 1e-14 * np.abs(ref).max()
        mm = w.get_lumped_mass_matrix(L0[i], A, Ix, Iy, J, 7850)
        refm = S.lumped_local_mass([L0[i]], A, Ix, Iy, J, 7850)[0]
        assert np.abs(mm - refm).max() <= 1e-14 * np.abs(refm).max()


@pytest.mark.parametrize("name", G.BEAM_CASES)
def test_assembly_matches_reference(name):
    c = G.load_beam(name)
    m = _model_from_case(c)
    m.assemble()
    indptr, indices, data = m.get_csr(L.MAT_K)
    mp, mi, md = m.get_csr(L.MAT_M)
    # run-to-run reproducibility: bit-identical values
    m.assemble()
    _, _, data2 = m.get_csr(L.MAT_K)
    assert np.array_equal(data, data2)
    m.close()
    es, props = G.elem_sec_and_props(c)
    Ko, Mo = S.frame_assemble(c["mesh"].points, c["mesh"].cells_dict["line"], es, props, c["E"], c["nu"])
    assert np.array_equal(indptr, Ko.indptr) and np.array_equal(indices, Ko.indices)   # pattern bit-exact
    n = Ko.shape[0]
    K = sp.csr_matrix((data, indices, indptr), shape=(n, n)).toarray()
    M = sp.csr_matrix((md, mi, mp), shape=(n, n)).toarray()
    Kr, Mr = c["ref"]["K_dense"], c["ref"]["M_dense"]
    assert np.abs(K - Kr).max() <= 1e-14 * np.abs(Kr).max()
    assert np.abs(M - Mr).max() <= 1e-14 * np.abs(Mr).max()
    assert np.abs(K - K.T).max() <= 1e-15 * np.abs(K).max()


def test_bulk_and_plain_store_paths_agree():
    c = G.load_beam("c3_lattice_3x3x4_jitter")
    vals = []
    os.environ["FEMB_ASM_GENERIC"] = "1"   # the bulk (cp.async.bulk) store path lives in the generic tile kernel
    for flag in ("1", "0"):
        os.environ["FEMB_ASM_BULK"] = flag
        m = _model_from_case(c)
        m.assemble()
        vals.append(m.get_csr(L.MAT_K)[2])
        vals.append(m.get_csr(L.MAT_M)[2])
        m.close()
    os.environ.pop("FEMB_ASM_BULK")
    os.environ.pop("FEMB_ASM_GENERIC")
    assert np.array_equal(vals[0], vals[2]) and np.array_equal(vals[1], vals[3])


def test_pair_kernel_agrees_with_generic_contribution_kernel():
    """The frame fast path (one thread per element end) and the generic one-thread-per-
    contribution kernel build the same K and M (same closed forms; only the association of the
    diagonal sums differs by design: both add in element-ascending order)."""
    mesh, sec, bc = meshgen.lattice_frame_case(7, 6, 5, jitter=0.05)
    es, props, _ = compat.frame_section_table(mesh, sec, csp)
    out = {}
    for flag in ("0", "1"):
        os.environ["FEMB_ASM_GENERIC"] = flag
        m = FrameModel(0)
        m.set_mesh(mesh.points, mesh.cells_dict["line"], es, props, 2e11, 2e11 / 2.6)
        m.assemble()
        out[flag] = (m.get_csr(L.MAT_K)[2], m.get_csr(L.MAT_M)[2])
        m.close()
    os.environ.pop("FEMB_ASM_GENERIC")
    for a, b in zip(out["0"], out["1"]):
        assert np.abs(a - b).max() <= 1e-15 * np.abs(b).max()


def test_duplicate_members_and_isolated_points():
    """Two members between the same nodes (ordered accumulation in one block) and a mesh
    point no element touches (all-zero rows, BeamSolver.py:354,360)."""
    pts = np.array([[0, 0, 0], [1.0, 0.2, 0.1], [2.0, 0.5, -0.3], [5.0, 5.0, 5.0]])
    conn = np.array([[0, 1], [1, 2], [1, 0], [0, 1]])
    props = np.array([[5e-3, 4e-6, 1e-6, 2.8e-6, 0.83, 0.83, 0.02, 0.05], [2e-3, 1e-6, 2e-6, 1e-6, 0.5, 0.6, 0.02, 0.05]])
    es = np.array([0, 1, 1, 0], dtype=np.int32)
    m = FrameModel(0)
    m.set_mesh(pts, conn, es, props, 2e11, 2e11 / 2.6)
    m.assemble()
    indptr, indices, data = m.get_csr(L.MAT_K)
    mp, mi, md = m.get_csr(L.MAT_M)
    m.close()
    Ko, Mo = S.frame_assemble(pts, conn, es, props, 2e11, 0.3)
    assert np.array_equal(indptr, Ko.indptr) and np.array_equal(indices, Ko.indices)
    assert np.abs(data - Ko.data).max() <= 1e-14 * np.abs(Ko.data).max()
    M = sp.csr_matrix((md, mi, mp), shape=Ko.shape).toarray()
    assert np.abs(M - Mo.toarray()).max() <= 1e-14 * np.abs(Mo.toarray()).max()
    assert np.all(data[indptr[18]:indptr[24]] == 0.0)


def _methods_for(name):
    if name.startswith("c3"):
        return [L.SOLVER_AUTO, L.SOLVER_PCG, L.SOLVER_DENSE]
    return [L.SOLVER_AUTO, L.SOLVER_CHAIN, L.SOLVER_DENSE]


@pytest.mark.parametrize("name", G.BEAM_CASES)
def test_static_solve_reactions_stress_match_reference(name):
    c = G.load_beam(name)
    fixed, f = compat.frame_bc_vectors(c["mesh"], c["bc"], len(c["mesh"].points))
    ref = c["ref"]
    oracle = S.frame_run(c["mesh"], c["props"], c["bc"], c["E"], c["nu"], k_modes=0)
    for method in _methods_for(name):
        m = _model_from_case(c)
        m.assemble()
        m.set_bc(fixed, f)
        u, r, st = m.solve_static(method=method)
        sig = m.stress()
        m.close()
        err = np.linalg.norm(u - ref["u"]) / np.linalg.norm(ref["u"])
        assert err <= 1e-10, (name, method, err, st)
        assert np.all(u[fixed] == 0.0)
        # reactions K u - f: vanish on free DOFs, balance the load on fixed ones, match the oracle
        assert np.linalg.norm(r - oracle["reactions"]) <= 1e-9 * np.linalg.norm(f), (name, method)
        smax = np.abs(ref["smoothed_stresses"]).max()
        assert np.abs(sig - ref["smoothed_stresses"]).max() <= 1e-8 * smax, (name, method)


def test_auto_picks_chain_for_chains_and_reports_it():
    c = G.load_beam("c2_simply_supported_24")
    fixed, f = compat.frame_bc_vectors(c["mesh"], c["bc"], len(c["mesh"].points))
    m = _model_from_case(c)
    m.assemble(); m.set_bc(fixed, f)
    _, _, st = m.solve_static()
    m.close()
    assert st["method_used"] == L.SOLVER_CHAIN


def test_compat_run_simulation_matches_reference_attributes():
    """The reference-shaped entry point fills u / smoothed_stresses like BeamSolver.py:418,438."""
    c = G.load_beam("c1_cantilever_beam")
    w = compat.BeamAnalysisB200(c["mesh"], [], c["bc"], c["E"], c["nu"],
                                props_fn=lambda t, p, r=False: c["props"][p["group"]])
    w.section_data = [{"group": g, "type": "x", "params": {"group": g}, "rotate": False} for g in c["props"]]
    w.run_simulation(k_modes=0)
    assert np.linalg.norm(w.u - c["ref"]["u"]) <= 1e-10 * np.linalg.norm(c["ref"]["u"])
    tip = w.u.reshape(-1, 6)[1]
    assert abs(tip[1] - (-0.01280624)) < 5e-9 and abs(tip[5] - (-0.0096)) < 1e-9      # closed form, SURVEY §4
    assert abs(w.smoothed_stresses[0] - 48e6) < 1.0
    r = w.reaction_forces.reshape(-1, 6)[0]
    assert abs(r[1] - 1000.0) < 1e-6 and abs(r[5] - 2000.0) < 1e-6
    # missing section group -> same error text as BeamSolver.py:368
    w.section_data = []
    msgs = []
    w.run_simulation(k_modes=0, on_error=lambda t, msg: msgs.append(msg))
    assert msgs and "Section properties not defined for physical group 'beam'" in msgs[0]


def test_singular_system_is_reported():
    """No supports -> K_ff singular: the reference's LAPACK raises; we return FEMB_ERR_SINGULAR
    or fail to converge, never garbage."""
    mesh, sec, bc = meshgen.lattice_frame_case(3, 3, 3, jitter=0.05)
    es, props, _ = compat.frame_section_table(mesh, sec, csp)
    m = FrameModel(0)
    m.set_mesh(mesh.points, mesh.cells_dict["line"], es, props, 2e11, 2e11 / 2.6)
    m.assemble()
    f = np.zeros(6 * len(mesh.points)); f[6 * 5 + 1] = 1.0
    m.set_bc(np.zeros(0, dtype=np.int64), f)
    with pytest.raises(L.FembError):
        m.solve_static(method=L.SOLVER_PCG, max_iter=2000)
    m.close()


@pytest.mark.parametrize("jitter", [0.0, 0.05])
def test_medium_lattice_pcg_vs_oracle(jitter):
    mesh, sec, bc = meshgen.lattice_frame_case(12, 10, 9, jitter=jitter)
    E, nu = meshgen.E_STEEL, meshgen.NU_STEEL
    es, props, _ = compat.frame_section_table(mesh, sec, csp)
    fixed, f = compat.frame_bc_vectors(mesh, bc, len(mesh.points))
    m = FrameModel(0)
    m.set_mesh(mesh.points, mesh.cells_dict["line"], es, props, E, E / (2 * (1 + nu)))
    m.assemble()
    indptr, indices, data = m.get_csr(L.MAT_K)
    m.set_bc(fixed, f)
    res = {}
    for pc in (L.PRECOND_JACOBI, L.PRECOND_BLOCK_JACOBI):
        u, r, st = m.solve_static(method=L.SOLVER_PCG, precond=pc)
        res[pc] = (u, r, st)
        u2, _, _ = m.solve_static(method=L.SOLVER_PCG, precond=pc)
        assert np.array_equal(u, u2), "PCG is not run-to-run reproducible"
    sig = m.stress()
    m.close()
    Ko, Mo = S.frame_assemble(mesh.points, mesh.cells_dict["line"], es, props, E, nu)
    assert np.array_equal(indptr, Ko.indptr) and np.array_equal(indices, Ko.indices)
    assert np.abs(data - Ko.data).max() <= 1e-14 * np.abs(Ko.data).max()
    _, free, _ = S.frame_bc(mesh, bc)
    uo, _ = S.solve_static(Ko, f, fixed, free, method="direct")
    for pc, (u, r, st) in res.items():
        assert np.linalg.norm(u - uo) <= 1e-10 * np.linalg.norm(uo), (pc, st)
        ro = Ko @ uo - f
        assert np.linalg.norm(r - ro) <= 1e-9 * np.linalg.norm(f)
    so = S.frame_stress(mesh.points, mesh.cells_dict["line"], es, props, E, nu, uo)
    assert np.abs(sig - so).max() <= 1e-8 * np.abs(so).max()


def test_full_size_c3_properties():
    """BASELINE config 3 at full size (56x56x54 lattice, 1,016,064 DOF): size-independent
    properties — residual on free DOFs, global equilibrium, reproducibility."""
    mesh, sec, bc = meshgen.lattice_frame_case(56, 56, 54, jitter=0.05)
    E, nu = meshgen.E_STEEL, meshgen.NU_STEEL
    es, props, _ = compat.frame_section_table(mesh, sec, csp)
    fixed, f = compat.frame_bc_vectors(mesh, bc, len(mesh.points))
    assert 6 * len(mesh.points) == 1016064 and len(fixed) == 18816
    m = FrameModel(0)
    m.set_mesh(mesh.points, mesh.cells_dict["line"], es, props, E, E / (2 * (1 + nu)))
    m.assemble()
    m.set_bc(fixed, f)
    u, r, st = m.solve_static(method=L.SOLVER_PCG, rtol=1e-12)
    m.close()
    assert st["converged"] == 1 and st["rel_residual"] <= 1e-12
    # FEMB_PRECOND_AUTO at this size: the line preconditioner, 768 bundles per member direction
    assert st["precond_used"] == L.PRECOND_LINES and st["coarse_dim"] == 3 * 768, st
    free = np.ones(len(f), dtype=bool); free[fixed] = False
    assert np.linalg.norm(r[free]) <= 1e-10 * np.linalg.norm(f)          # K u = f on free DOFs
    R = r.reshape(-1, 6)[:, :3].sum(axis=0)
    F = f.reshape(-1, 6)[:, :3].sum(axis=0)
    assert np.abs(R + F).max() <= 1e-7 * np.abs(F).max()                 # sum of reactions = -sum of loads
    assert np.all(u[fixed] == 0.0) and np.isfinite(u).all()


def test_same_topology_reuses_symbolic_analysis_correctly():
    """Re-running on the same connectivity keeps the symbolic analysis (pattern, pair records): new
    coordinates, new section assignment and a new topology must each give the oracle's K."""
    E, nu = meshgen.E_STEEL, meshgen.NU_STEEL
    m = FrameModel(0)
    variants = []
    mesh, sec, bc = meshgen.lattice_frame_case(5, 4, 4, jitter=0.0)
    es, props, _ = compat.frame_section_table(mesh, sec, csp)
    conn = mesh.cells_dict["line"]
    variants.append((mesh.points, conn, es))
    mesh2, _, _ = meshgen.lattice_frame_case(5, 4, 4, jitter=0.08)
    variants.append((mesh2.points, conn, es))                       # same topology, new coordinates
    variants.append((mesh2.points, conn, (es + 1) % len(props)))    # new section assignment
    variants.append((mesh2.points, conn[::-1].copy(), es[::-1].copy()))   # new element order -> new topology
    for pts, cn, sec_ids in variants:
        m.set_mesh(pts, cn, sec_ids, props, E, E / (2 * (1 + nu)))
        m.assemble()
        indptr, indices, data = m.get_csr(L.MAT_K)
        Ko, _ = S.frame_assemble(pts, cn, sec_ids, props, E, nu)
        assert np.array_equal(indptr, Ko.indptr) and np.array_equal(indices, Ko.indices)
        assert np.abs(data - Ko.data).max() <= 1e-14 * np.abs(Ko.data).max()
        fixed, f = compat.frame_bc_vectors(mesh, bc, len(pts))
        m.set_bc(fixed, f)
        u, _, st = m.solve_static(method=L.SOLVER_PCG)
        _, free, _ = S.frame_bc(mesh, bc)
        uo, _ = S.solve_static(Ko, f, fixed, free, method="direct")
        assert np.linalg.norm(u - uo) <= 1e-10 * np.linalg.norm(uo)
    m.close()


def test_qr_algorithm_helper_reproduces_the_reference_iteration():
    """compat.BeamAnalysisB200.qr_algorithm (BeamSolver.py:467-481 signature) on inv(M_ff) K_ff of the shipped
    cantilever: same eigenvalues as the unmodified reference's own loop (tests/golden, omega = sqrt(lambda)) and
    orthonormal accumulated vectors."""
    c = G.load_beam("c1_cantilever_beam")
    raw = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c1_cantilever_beam.npz"), allow_pickle=True)
    K, M = raw["ref_K_dense"], raw["ref_M_dense"]
    _, free, _ = S.frame_bc(c["mesh"], c["bc"])
    A = np.linalg.inv(M[np.ix_(free, free)]) @ K[np.ix_(free, free)]
    w = compat.BeamAnalysisB200.__new__(compat.BeamAnalysisB200)
    w.device = 0
    lam, V = w.qr_algorithm(A)
    assert lam.shape == (len(free),) and V.shape == (len(free), len(free)) and np.all(np.diff(lam) >= 0)
    om = np.sqrt(lam[lam > 1e-6])
    ref = np.sort(np.asarray(raw["ref_natural_frequencies"]))
    assert np.abs(om[:len(ref)] - ref).max() <= 1e-6 * ref.max(), (om[:4], ref[:4])
    assert np.abs(V.T @ V - np.eye(len(free))).max() <= 1e-10
