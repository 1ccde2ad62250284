"""GPU tests of the row-block distributed path that need only ONE GPU: world = 1 goes through
csrc/dist.cu end to end (no neighbours, NCCL never loaded), and the per-rank local assemblies of a
3-way partition are run one after the other and compared bit for bit with the global matrix.
The multi-rank NCCL run itself is exercised by scripts/dist_solve.py under torchrun (profiles/)."""
import numpy as np
import pytest
import scipy.sparse as sp

from fem_calculator_b200 import _lib as L
from fem_calculator_b200 import compat, meshgen, partition as P
from fem_calculator_b200.api import DistFrameModel, FrameModel
from fem_calculator_b200.sections import calculate_section_properties as csp
from oracle import ref_sparse as S

pytestmark = pytest.mark.gpu


def _case(nx=9, ny=6, nz=7):
    mesh, sec, bc = meshgen.lattice_frame_case(nx, ny, nz, jitter=0.05)
    es, props, _ = compat.frame_section_table(mesh, sec, csp)
    fixed, f = compat.frame_bc_vectors(mesh, bc, len(mesh.points))
    return mesh, es, props, fixed, f


@pytest.mark.parametrize("precond", [L.PRECOND_JACOBI, L.PRECOND_BLOCK_JACOBI])
def test_world1_dist_path_matches_oracle_and_single_gpu_path(precond):
    mesh, es, props, fixed, f = _case()
    E, nu = meshgen.E_STEEL, meshgen.NU_STEEL
    G = E / (2 * (1 + nu))
    conn = mesh.cells_dict["line"]
    d = DistFrameModel(0)
    d.setup(mesh.points, conn, es, props, E, G, fixed, f, rank=0, world=1)
    u, r, st = d.solve_static_dist(precond=precond)
    u2, _, _ = d.solve_static_dist(precond=precond)
    d.close()
    assert st["converged"] == 1 and np.array_equal(u, u2)
    m = FrameModel(0)
    m.set_mesh(mesh.points, conn, es, props, E, G)
    m.assemble(); m.set_bc(fixed, f)
    us, rs, sts = m.solve_static(method=L.SOLVER_PCG, precond=precond)
    m.close()
    K, _ = S.frame_assemble(mesh.points, conn, es, props, E, nu)
    free = np.setdiff1d(np.arange(len(f)), fixed)
    uo, _ = S.solve_static(K, f, fixed, free, method="direct")
    assert np.linalg.norm(u - uo) <= 1e-10 * np.linalg.norm(uo)
    assert np.linalg.norm(u - us) <= 1e-10 * np.linalg.norm(uo)
    assert np.linalg.norm(r - (K @ uo - f)) <= 1e-9 * np.linalg.norm(f)
    assert abs(st["iterations"] - sts["iterations"]) <= 2


def test_local_assemblies_reproduce_global_rows_bit_for_bit():
    mesh, es, props, fixed, f = _case()
    E, nu = meshgen.E_STEEL, meshgen.NU_STEEL
    G = E / (2 * (1 + nu))
    conn = mesh.cells_dict["line"]
    n = len(mesh.points)
    m = FrameModel(0)
    m.set_mesh(mesh.points, conn, es, props, E, G)
    m.assemble()
    ip, ix, iv = m.get_csr(L.MAT_K)
    m.close()
    Kg = sp.csr_matrix((iv, ix, ip), shape=(6 * n, 6 * n))
    for r in range(3):
        p = P.partition_mesh(conn, n, 3, r)
        ml = FrameModel(0)
        ml.set_mesh(mesh.points[p.local_nodes], p.conn_local, es[p.elem_ids], props, E, G)
        ml.assemble()
        lp, lx, lv = ml.get_csr(L.MAT_K)
        ml.close()
        nl = 6 * len(p.local_nodes)
        Kl = sp.csr_matrix((lv, lx, lp), shape=(nl, nl))
        ld = p.local_dofs(6)
        rows = slice(0, 6 * p.n_owned)
        ref = Kg[ld[rows]][:, ld]
        assert abs(Kl[rows] - ref).max() == 0.0      # owner computes: identical contributions in identical order
