"""GPU parity tests for the Tet10 path (ReactionSolver.py) through the C ABI."""
import numpy as np
import pytest
import scipy.sparse as sp

import golden_util as G
from fem_calculator_b200 import _lib as L
from fem_calculator_b200 import compat, meshgen
from fem_calculator_b200.api import Tet10Model
from oracle import ref_sparse as S

pytestmark = pytest.mark.gpu


def test_tet10_element_matrices_match_oracle():
    c = G.load_tet("tet10_box_2x1x2")
    m = Tet10Model(0)
    m.set_mesh(c["mesh"].points, c["mesh"].cells_dict["tetra10"], c["E"], c["nu"])
    ke = m.elements()
    m.close()
    ko, skipped = S.tet10_element_matrices(c["mesh"].points, c["mesh"].cells_dict["tetra10"], c["E"], c["nu"])
    scale = np.abs(ko).max(axis=(1, 2), keepdims=True)
    assert (np.abs(ke - ko) <= 1e-13 * scale).all(), float((np.abs(ke - ko) / scale).max())


@pytest.mark.parametrize("name", G.TET_CASES)
def test_force_analysis_matches_reference(name):
    c = G.load_tet(name)
    ref = c["ref"]
    fa = compat.ForceAnalysisB200(c["mesh"], c["force_data"], c["fix_data"], c["E"], c["nu"])
    fa.assemble_stiffness_matrix()
    fa.apply_boundary_conditions()
    fa.solve()
    assert fa.negative_detJ_count == int(ref["negative_detJ_count"])
    assert np.array_equal(fa.fixed_dofs, ref["fixed_dofs"]) and np.array_equal(fa.active_dofs, ref["active_dofs"])
    assert np.array_equal(fa.f, ref["f"])
    assert [int(i["node_idx"]) for i in fa.fixed_nodes_info] == ref["fixed_nodes"].tolist()
    assert np.linalg.norm(fa.u - ref["u"]) <= 1e-10 * np.linalg.norm(ref["u"]), fa.solve_stats
    assert np.linalg.norm(fa.reaction_forces - ref["reaction_forces"]) <= 1e-9 * np.linalg.norm(ref["f"])
    Ko, _ = S.tet10_assemble(c["mesh"].points, c["mesh"].cells_dict["tetra10"], c["E"], c["nu"])
    assert np.array_equal(fa.K.indptr, Ko.indptr) and np.array_equal(fa.K.indices, Ko.indices)
    assert abs(fa.K - Ko).max() <= 1e-13 * abs(Ko).max()
    if "K_data" in ref:
        Kr = sp.csr_matrix((ref["K_data"], ref["K_indices"], ref["K_indptr"]), shape=fa.K.shape)
        assert abs(fa.K - Kr).max() <= 1e-13 * abs(Kr).max()
    # equilibrium print-out (ReactionSolver.py:218-224): sum of reactions = -applied
    tot = sum(fa.reaction_forces[3 * i["node_idx"]: 3 * i["node_idx"] + 3] for i in fa.fixed_nodes_info)
    assert np.abs(tot + np.array([0.0, 3000.0, 0.0])).max() < 1e-6
    fa.close()


def test_tet10_solvers_agree_and_are_reproducible():
    mesh, fd, xd = meshgen.tet10_box_case(8, 2, 8)
    out = S.tet10_run(mesh, fd, xd, 2e11, 0.3)
    fa = compat.ForceAnalysisB200(mesh, fd, xd, 2e11, 0.3)
    fa.assemble_stiffness_matrix(export_csr=False)
    fa.apply_boundary_conditions()
    fa.solve(method=L.SOLVER_PCG)
    u1 = fa.u.copy()
    fa.solve(method=L.SOLVER_PCG)
    assert np.array_equal(u1, fa.u)
    assert np.linalg.norm(fa.u - out["u"]) <= 1e-10 * np.linalg.norm(out["u"]), fa.solve_stats
    assert np.linalg.norm(fa.reaction_forces - out["reaction_forces"]) <= 1e-9 * np.linalg.norm(out["f"])
    fa.close()


def test_inverted_elements_are_counted_and_skipped():
    """ReactionSolver.py:133-135: Gauss points with detJ <= 1e-12 are skipped and counted."""
    mesh, fd, xd = meshgen.tet10_box_case(2, 1, 2)
    conn = mesh.cells_dict["tetra10"].copy()
    conn[3] = conn[3][[0, 2, 1, 3, 6, 5, 4, 7, 9, 8]]        # swap two corners: negative Jacobian
    Ko, skipped = S.tet10_assemble(mesh.points, conn, 2e11, 0.3)
    assert skipped == 4
    m = Tet10Model(0)
    m.set_mesh(mesh.points, conn, 2e11, 0.3)
    m.assemble()
    indptr, indices, data = m.get_csr(L.MAT_K)
    assert m.negative_detj == skipped
    m.close()
    assert np.abs(data - Ko.data).max() <= 1e-13 * np.abs(Ko.data).max()
