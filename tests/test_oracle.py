"""Pins the CPU oracle (oracle/ref_sparse.py) to the unmodified reference's outputs
(tests/golden/*.npz) and to closed-form known answers.  CPU only."""
import numpy as np
import pytest

from oracle import ref_sparse as S
import golden_util as G


@pytest.mark.parametrize("name", G.BEAM_CASES)
def test_frame_matrices_match_reference(name):
    c = G.load_beam(name)
    es, props = G.elem_sec_and_props(c)
    K, M = S.frame_assemble(c["mesh"].points, c["mesh"].cells_dict["line"], es, props, c["E"], c["nu"])
    Kr, Mr = c["ref"]["K_dense"], c["ref"]["M_dense"]
    assert np.abs(K.toarray() - Kr).max() <= 1e-14 * np.abs(Kr).max()
    assert np.abs(M.toarray() - Mr).max() <= 1e-14 * np.abs(Mr).max()
    # structural pattern covers every non-zero of the reference's dense matrix
    assert not (Kr != 0)[~(K.toarray() != 0) & ~_structural(K)].any()


def _structural(K):
    m = np.zeros(K.shape, dtype=bool)
    coo = K.tocoo()
    m[coo.row, coo.col] = True
    return m


@pytest.mark.parametrize("name", G.BEAM_CASES)
def test_frame_static_stress_modal_match_reference(name):
    c = G.load_beam(name)
    out = S.frame_run(c["mesh"], c["props"], c["bc"], c["E"], c["nu"], k_modes=10)
    ref = c["ref"]
    assert np.linalg.norm(out["u"] - ref["u"]) <= 1e-10 * np.linalg.norm(ref["u"])
    smax = np.abs(ref["smoothed_stresses"]).max()
    assert np.abs(out["smoothed_stresses"] - ref["smoothed_stresses"]).max() <= 1e-9 * smax
    # the reference's QR iteration is only loosely converged: 1e-5 cross-check
    k = min(6, len(ref["natural_frequencies"]))
    rel = np.abs(out["natural_frequencies"][:k] - ref["natural_frequencies"][:k]) / ref["natural_frequencies"][:k]
    assert rel.max() <= 1e-5
    # equilibrium: support reactions balance the applied load
    R = out["reactions"].reshape(-1, 6)[:, :3].sum(axis=0)
    F = out["f"].reshape(-1, 6)[:, :3].sum(axis=0)
    assert np.abs(R + F).max() <= 1e-8 * max(1.0, np.abs(F).max())
    fixed = out["fixed"]
    rf = np.zeros_like(out["u"]); rf[fixed] = out["reactions"][fixed]
    assert np.abs(rf.reshape(-1, 6)[:, :3].sum(axis=0) + F).max() <= 1e-7 * max(1.0, np.abs(F).max())


def test_element_vectors_match_reference():
    e = G.load_elements()
    E, nu = float(e["E"]), float(e["nu"])
    ke, me = S.frame_element_matrices(e["points"], e["line"], e["elem_sec"], e["props"], E, nu)
    for a, b in ((ke, e["ref_ke_global"]), (me, e["ref_me_global"])):
        scale = np.abs(b).max(axis=(1, 2), keepdims=True)
        assert (np.abs(a - b) <= 1e-14 * scale).all()
    L, _ = S.frame_rotation(e["points"], e["line"])
    pr = e["props"]
    kl = S.timoshenko_local_stiffness(L, E, E / (2 * (1 + nu)), pr[:, 0], pr[:, 1], pr[:, 2], pr[:, 3], pr[:, 4], pr[:, 5])
    scale = np.abs(e["ref_k_local"]).max(axis=(1, 2), keepdims=True)
    assert (np.abs(kl - e["ref_k_local"]) <= 4e-16 * scale).all()
    assert ((kl != 0) == (e["ref_k_local"] != 0)).all()


def test_cantilever_closed_form():
    """Known answer on the shipped mesh: 2-node Timoshenko element is nodally exact for
    end loads: u_y = PL^3/3EI + PL/(kappa G A), theta_z = PL^2/2EI."""
    c = G.load_beam("c1_cantilever_beam")
    out = S.frame_run(c["mesh"], c["props"], c["bc"], c["E"], c["nu"], k_modes=4)
    E, nu, P, L = c["E"], c["nu"], -1000.0, 2.0
    A, Ix, Iy, J, ky, kz, cy, cz = c["props"]["beam"]
    Gm = E / (2 * (1 + nu))
    uy = P * L**3 / (3 * E * Iy) + P * L / (ky * Gm * A)
    tip = out["u"].reshape(-1, 6)[1]
    assert abs(tip[1] - uy) <= 1e-12 * abs(uy)
    assert abs(tip[5] - P * L**2 / (2 * E * Iy)) <= 1e-12 * abs(tip[5])
    assert abs(tip[1] - (-0.01280624)) < 5e-9
    r = out["reactions"].reshape(-1, 6)[0]
    assert abs(r[1] - 1000.0) < 1e-6 and abs(r[5] - 2000.0) < 1e-6
    assert abs(out["smoothed_stresses"][0] - 48e6) < 1e-3


@pytest.mark.parametrize("name", G.TET_CASES)
def test_tet10_matches_reference(name):
    c = G.load_tet(name)
    out = S.tet10_run(c["mesh"], c["force_data"], c["fix_data"], c["E"], c["nu"])
    ref = c["ref"]
    assert out["negative_detJ_count"] == int(ref["negative_detJ_count"])
    assert np.array_equal(out["fixed_dofs"], ref["fixed_dofs"])
    assert np.array_equal(out["active_dofs"], ref["active_dofs"])
    assert np.array_equal(out["f"], ref["f"])
    assert np.linalg.norm(out["u"] - ref["u"]) <= 1e-10 * np.linalg.norm(ref["u"])
    assert np.linalg.norm(out["reaction_forces"] - ref["reaction_forces"]) <= 1e-9 * np.linalg.norm(ref["f"])
    if "K_data" in ref:
        import scipy.sparse as sp
        Kr = sp.csr_matrix((ref["K_data"], ref["K_indices"], ref["K_indptr"]), shape=out["K"].shape)
        K = out["K"]
        assert abs(K - Kr).max() <= 1e-14 * abs(Kr).max()
        # the reference's lil->csr pattern (exact zeros dropped) is a subset of the structural one
        S_ = K.copy(); S_.data[:] = 1.0
        R_ = Kr.copy(); R_.data[:] = 1.0
        assert (R_ - R_.multiply(S_)).nnz == 0
    # equilibrium (ReactionSolver.py:218-224)
    nodes = ref["fixed_nodes"]
    tot = out["reaction_forces"].reshape(-1, 3)[nodes].sum(axis=0)
    assert np.abs(tot + np.array([0.0, 3000.0, 0.0])).max() < 1e-6
