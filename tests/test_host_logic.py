"""CPU-only tests: C-ABI library loads and exports every declared symbol, host-side symbolic
phase, MSH reader, section front end, BC bookkeeping (no compute calls: no GPU here)."""
import os
import re

import numpy as np
import pytest

import golden_util as G
from fem_calculator_b200 import _lib, api, compat, meshgen, msh, sections
from oracle import ref_sparse as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "femb200.h")).read()
    declared = set(re.findall(r"\b(femb_[a-z0-9_]+)\s*\(", header))
    declared -= {"femb_handle"}
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), f"libfemb200.so does not export {name}"
        assert name in _lib.SIGNATURES, f"ctypes binding missing for {name}"
    assert lib.femb_version() >= 100


def test_no_cpu_fallback_without_gpu():
    if _lib.device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(_lib.FembError):
        api.FrameModel(0)


@pytest.mark.parametrize("name", G.BEAM_CASES)
def test_symbolic_pattern_matches_oracle_frame(name):
    c = G.load_beam(name)
    mesh = c["mesh"]
    rowptr, colidx = api.symbolic_pattern(len(mesh.points), mesh.cells_dict["line"])
    es, props = G.elem_sec_and_props(c)
    K, _ = S.frame_assemble(mesh.points, mesh.cells_dict["line"], es, props, c["E"], c["nu"])
    # expand block pattern to scalar CSR pattern and compare bit-exact with the oracle's
    indptr, indices = _expand(rowptr, colidx, 6)
    assert np.array_equal(indptr, K.indptr)
    assert np.array_equal(indices, K.indices)


def test_symbolic_pattern_matches_oracle_tet10():
    c = G.load_tet("tet10_box_2x1x2")
    mesh = c["mesh"]
    rowptr, colidx = api.symbolic_pattern(len(mesh.points), mesh.cells_dict["tetra10"])
    K, _ = S.tet10_assemble(mesh.points, mesh.cells_dict["tetra10"], c["E"], c["nu"])
    indptr, indices = _expand(rowptr, colidx, 3)
    assert np.array_equal(indptr, K.indptr)
    assert np.array_equal(indices, K.indices)


def _expand(rowptr, colidx, bs):
    n = len(rowptr) - 1
    nbr = np.diff(rowptr)
    indptr = np.zeros(n * bs + 1, dtype=np.int64)
    indptr[1:] = np.cumsum(np.repeat(nbr * bs, bs))
    idx = []
    for i in range(n):
        cols = (colidx[rowptr[i]:rowptr[i + 1]][:, None] * bs + np.arange(bs)[None, :]).ravel()
        for _ in range(bs):
            idx.append(cols)
    return indptr, np.concatenate(idx)


def test_symbolic_isolated_node_and_duplicates():
    # node 3 is touched by no element; elements 0 and 2 connect the same pair
    conn = np.array([[0, 1], [1, 2], [1, 0]])
    rowptr, colidx = api.symbolic_pattern(4, conn)
    assert rowptr.tolist() == [0, 2, 5, 7, 8]
    assert colidx.tolist() == [0, 1, 0, 1, 2, 1, 2, 3]


def test_msh_reader_on_reference_layout(tmp_path):
    text = """$MeshFormat
4.1 0 8
$EndMeshFormat
$PhysicalNames
3
0 2 "fix"
0 3 "load_y"
1 4 "beam"
$EndPhysicalNames
$Entities
2 1 0 0
1 0 0 0 1 2 
2 2 0 0 1 3 
1 0 0 0 2 0 0 1 4 2 1 -2 
$EndEntities
$Nodes
3 3 1 3
0 1 0 1
1
0 0 0
0 2 0 1
2
2 0 0
1 1 0 1
3
0.9999999999973884 0 0
$EndNodes
$Elements
3 4 1 4
0 1 15 1
1 1 
0 2 15 1
2 2 
1 1 1 2
3 1 3 
4 3 2 
$EndElements
"""
    p = tmp_path / "cantilever_beam"
    p.write_text(text)
    m = msh.read_msh(str(p))
    c = G.load_beam("c1_cantilever_beam")["mesh"]   # parsed from the shipped file when the fixture was made
    assert np.array_equal(m.points, c.points)
    assert np.array_equal(m.cells_dict["line"], c.cells_dict["line"])
    assert np.array_equal(m.cells_dict["vertex"], c.cells_dict["vertex"])
    assert {k: v.tolist() for k, v in m.field_data.items()} == {"fix": [2, 0], "load_y": [3, 0], "beam": [4, 1]}
    assert m.cell_data_dict["gmsh:physical"]["line"].tolist() == [4, 4]
    assert m.group_nodes("fix").tolist() == [0] and m.group_nodes("load_y").tolist() == [1]


def test_msh_roundtrip(tmp_path):
    for mesh in (meshgen.lattice_frame_case(3, 2, 3, jitter=0.05)[0], meshgen.tet10_box_case(2, 1, 2)[0]):
        p = str(tmp_path / "m.msh")
        msh.write_msh(p, mesh)
        back = msh.read_msh(p)
        assert np.array_equal(back.points, mesh.points)
        for k in mesh.cells_dict:
            # entity grouping may reorder cells between physical tags; compare as sorted sets
            a = np.concatenate([mesh.cells_dict[k], mesh.cell_data_dict["gmsh:physical"][k][:, None]], axis=1)
            b = np.concatenate([back.cells_dict[k], back.cell_data_dict["gmsh:physical"][k][:, None]], axis=1)
            assert np.array_equal(a[np.lexsort(a.T[::-1])], b[np.lexsort(b.T[::-1])])
        assert {k: v.tolist() for k, v in back.field_data.items()} == {k: v.tolist() for k, v in mesh.field_data.items()}


def test_sections_closed_form_and_rotate():
    A, Ix, Iy, J, ky, kz, cy, cz = sections.calculate_section_properties("rectangular section", {"d": 0.1, "b": 0.05})
    assert abs(A - 5e-3) < 1e-15 and abs(Ix - 0.05 * 0.1**3 / 12) < 1e-18 and abs(Iy - 0.1 * 0.05**3 / 12) < 1e-18
    assert abs(ky - 5 / 6) < 1e-15 and (cy, cz) == (0.025, 0.05)
    r = sections.calculate_section_properties("rectangular section", {"d": 0.1, "b": 0.05}, rotate=True)
    assert (r[1], r[2], r[6], r[7]) == (Iy, Ix, cz, cy)          # BeamSolver.py:76-77
    assert sections.calculate_section_properties("nonsense", {}) == (0,) * 8   # :55-57
    for t, p in [("I section", {"d": 0.2, "b": 0.1, "t_f": 0.0085, "t_w": 0.0056, "r": 0}),
                 ("C section", {"d": 0.1, "b": 0.05, "t_f": 0.005, "t_w": 0.005, "r": 0}),
                 ("L section", {"d": 0.075, "b": 0.075, "t": 0.006, "r_r": 0, "r_t": 0}),
                 ("hollow box section", {"d": 0.1, "b": 0.1, "t": 0.005, "r_out": 0}),
                 ("circular section", {"d": 0.05}), ("hollow circular section", {"d": 0.12, "t": 0.008})]:
        v = sections.calculate_section_properties(t, p)
        assert all(x > 0 for x in v), (t, v)


@pytest.mark.parametrize("name", G.BEAM_CASES)
def test_bc_bookkeeping_matches_oracle(name):
    c = G.load_beam(name)
    fixed, f = compat.frame_bc_vectors(c["mesh"], c["bc"], len(c["mesh"].points))
    ofixed, ofree, of = S.frame_bc(c["mesh"], c["bc"])
    assert np.array_equal(fixed, ofixed) and np.array_equal(f, of)


def test_mesh_generators_sizes():
    mesh, sec, bc = meshgen.lattice_frame_case(5, 4, 3)
    assert len(mesh.points) == 60 and len(mesh.cells_dict["line"]) == 4 * 4 * 3 + 5 * 3 * 3 + 5 * 4 * 2
    # node numbering: z fastest
    assert np.allclose(mesh.points[1], [0, 0, 1]) and np.allclose(mesh.points[3], [0, 1, 0])
    # BASELINE config 3 sizes: 56x56x54 -> 169,344 nodes, 498,848 elements
    assert 56 * 56 * 54 == 169344 and 55 * 56 * 54 * 2 + 56 * 56 * 53 == 498848
    tm, fd, xd = meshgen.tet10_box_case(2, 1, 2)
    assert len(tm.cells_dict["tetra10"]) == 24 and len(tm.points) == 75


def test_aggregates_are_a_balanced_partition():
    """Two-level preconditioner, host side (csrc/coarse.cpp): recursive coordinate bisection gives
    every aggregate floor/ceil(n/parts) nodes, is deterministic, and cuts a lattice into boxes."""
    mesh, _, _ = meshgen.lattice_frame_case(12, 10, 9, jitter=0.05)
    for parts in (1, 2, 7, 40, 444):
        agg = api.symbolic_aggregates(mesh.points, parts)
        cnt = np.bincount(agg, minlength=parts)
        n = len(mesh.points)
        assert agg.min() >= 0 and agg.max() < parts and cnt.sum() == n
        assert cnt.max() - cnt.min() <= 1 or parts > n // 2, (parts, cnt.min(), cnt.max())
        assert np.array_equal(agg, api.symbolic_aggregates(mesh.points, parts))
    agg = api.symbolic_aggregates(mesh.points, 8)
    for a in range(8):                      # 8 parts of a box: every part is itself a box of ~1/8 the volume
        p = mesh.points[agg == a]
        ext = p.max(0) - p.min(0)
        assert (ext <= np.array([12, 10, 9]) * 0.5 + 0.2).all(), (a, ext)
    assert api.symbolic_aggregates(np.zeros((0, 3)), 4).size == 0


@pytest.mark.parametrize("parts", [1, 5, 37])
def test_coarse_symbolic_slots_and_adjacency(parts):
    """Aggregate adjacency and per-block slots (csrc/coarse.cpp) against a numpy restatement from the
    block pattern: neighbour lists sorted, unique, self included, symmetric; every block's slot points
    at the aggregate of its column node."""
    mesh, _, _ = meshgen.lattice_frame_case(9, 7, 6, jitter=0.05)
    conn = mesh.cells_dict["line"]
    n = len(mesh.points)
    agg, nbr_ptr, nbr, slot = api.symbolic_coarse(mesh.points, conn, parts)
    assert np.array_equal(agg, api.symbolic_aggregates(mesh.points, parts))
    rowptr, colidx = api.symbolic_pattern(n, conn)
    rows = np.repeat(np.arange(n), np.diff(rowptr))
    pairs = set(zip(agg[rows].tolist(), agg[colidx].tolist()))
    for a in range(parts):
        lst = nbr[nbr_ptr[a]:nbr_ptr[a + 1]]
        assert np.all(np.diff(lst) > 0) and a in lst
        assert set(lst.tolist()) == {j for (i, j) in pairs if i == a}
        for j in lst:
            assert a in nbr[nbr_ptr[j]:nbr_ptr[j + 1]]
    assert np.array_equal(nbr[nbr_ptr[agg[rows]] + slot], agg[colidx])


def _check_lines_cover(conn, line_ptr, line_nodes):
    """consecutive nodes of a line are joined by an element; no element is used twice"""
    key = {tuple(sorted(e)): k for k, e in enumerate(conn.tolist())}
    used = set()
    for a, b in zip(line_ptr[:-1], line_ptr[1:]):
        nodes = line_nodes[a:b].tolist()
        for u, v in zip(nodes[:-1], nodes[1:]):
            k = key[tuple(sorted((u, v)))]
            assert k not in used
            used.add(k)
    return used


@pytest.mark.parametrize("jitter", [0.0, 0.05])
def test_member_lines_of_a_lattice(jitter):
    """Member-line detection (csrc/coarse.cpp): a 6x5x4 lattice has 20 + 24 + 30 lines of 6 / 5 / 4
    nodes, grouped by direction, and together they use every member exactly once."""
    mesh, _, _ = meshgen.lattice_frame_case(6, 5, 4, jitter=jitter)
    conn = mesh.cells_dict["line"]
    lp, ln, ld, lf = api.symbolic_lines(mesh.points, conn)
    assert len(lf) == 74 and np.array_equal(np.bincount(lf), [20, 24, 30])
    for fam, length in ((0, 6), (1, 5), (2, 4)):
        assert np.all(np.diff(lp)[lf == fam] == length)
        assert np.all(ld[lf == fam][:, fam] > 0.98)
    assert len(_check_lines_cover(conn, lp, ln)) == len(conn)
    again = api.symbolic_lines(mesh.points, conn)
    assert all(np.array_equal(x, y) for x, y in zip((lp, ln, ld, lf), again))


def test_member_lines_corners_rings_and_short_members():
    # an L: 4 members along x, then 3 along y -> two lines meeting at the corner node
    pts = np.array([[i, 0, 0] for i in range(5)] + [[4, j, 0] for j in range(1, 4)], dtype=float)
    conn = np.array([[i, i + 1] for i in range(7)])
    lp, ln, ld, lf = api.symbolic_lines(pts, conn)
    assert len(lf) == 2 and sorted(np.diff(lp).tolist()) == [4, 5] and sorted(lf.tolist()) == [0, 1]
    # a square ring with subdivided sides -> four lines; a regular 60-gon (6 degrees per corner) -> one closed line
    side = [[i, 0, 0] for i in range(4)] + [[4, i, 0] for i in range(4)] + [[4 - i, 4, 0] for i in range(4)] + [[0, 4 - i, 0] for i in range(4)]
    ring = np.array([[i, (i + 1) % 16] for i in range(16)])
    lp, ln, ld, lf = api.symbolic_lines(np.array(side, dtype=float), ring)
    assert len(lf) == 4 and np.all(np.diff(lp) == 5)
    t = 2 * np.pi * np.arange(60) / 60
    lp, ln, ld, lf = api.symbolic_lines(np.stack([np.cos(t), np.sin(t), 0 * t], 1), np.array([[i, (i + 1) % 60] for i in range(60)]))
    assert len(lf) == 1 and lp[1] == 61 and ln[0] == ln[-1]
    # isolated single members are dropped at min_nodes = 3 and kept at 2; empty meshes are fine
    pts = np.array([[0, 0, 0], [1, 0, 0], [0, 5, 0], [0, 5, 1]], dtype=float)
    two = np.array([[0, 1], [2, 3]])
    assert len(api.symbolic_lines(pts, two)[3]) == 0
    assert len(api.symbolic_lines(pts, two, min_nodes=2)[3]) == 2
    assert len(api.symbolic_lines(np.zeros((0, 3)), np.zeros((0, 2), dtype=np.int64))[3]) == 0


def test_member_line_modes_cut_the_iteration_count():
    """Why the lines matter: with one axial translation mode per detected line next
    to the rigid-body modes of the aggregates, the additive two-level PCG of csrc/twolevel.cu (restated in
    numpy on the oracle's K) needs several times fewer iterations."""
    import scipy.sparse as sp
    from prototypes.coarse_space_study import pcg, rigid_body_modes
    mesh, sec, bc = meshgen.lattice_frame_case(10, 9, 8, jitter=0.05)
    es, props = meshgen.section_table(mesh, sec)
    K, _ = S.frame_assemble(mesh.points, mesh.cells_dict["line"], es, props, meshgen.E_STEEL, meshgen.NU_STEEL)
    fixed, free, f = S.frame_bc(mesh, bc)
    n = K.shape[0]
    mask = np.zeros(n, bool); mask[free] = True
    Dm = sp.diags(mask.astype(float))
    A = (Dm @ K @ Dm + sp.diags((~mask).astype(float))).tocsr()
    b, d = f * mask, A.diagonal()
    agg = api.symbolic_aggregates(mesh.points, 4).astype(np.int64)
    P = rigid_body_modes(mesh.points, agg, 4, mask)
    lp, ln, ld, lf = api.symbolic_lines(mesh.points, mesh.cells_dict["line"])
    rows, cols, vals = [], [], []
    for k in range(len(lf)):
        nodes = ln[lp[k]:lp[k + 1]]
        for c in range(3):
            rows.append(6 * nodes + c); cols.append(np.full(len(nodes), k)); vals.append(np.full(len(nodes), ld[k, c]))
    Pl = Dm @ sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, len(lf)))

    def iterations(Pc):
        Kc = (Pc.T @ A @ Pc).toarray()
        zero = np.diag(Kc) <= 0
        Kc[zero, zero] = 1.0
        Kci = np.linalg.inv(Kc)
        x, it = pcg(A, b, lambda r: 2 * r / d + Pc @ (Kci @ (Pc.T @ r)))
        assert np.linalg.norm(A @ x - b) <= 1e-11 * np.linalg.norm(b)
        return it

    it_rbm, it_lines = iterations(P.tocsr()), iterations(sp.hstack([P, Pl]).tocsr())
    assert it_lines < 0.3 * it_rbm, (it_rbm, it_lines)


def test_line_precond_spec_matches_matrix_form():
    """tests/prototypes/line_precond_spec.py (the loop-level specification the round-2 kernels will be
    transcribed from) against plain sparse algebra: per-direction Galerkin blocks, segment diagonal and
    one application of the preconditioner."""
    import scipy.sparse as sp
    from prototypes import line_precond_spec as LS
    mesh, sec, bc = meshgen.lattice_frame_case(7, 6, 5, jitter=0.05)
    es, props = meshgen.section_table(mesh, sec)
    conn = mesh.cells_dict["line"]
    K, _ = S.frame_assemble(mesh.points, conn, es, props, meshgen.E_STEEL, meshgen.NU_STEEL)
    fixed, free, f = S.frame_bc(mesh, bc)
    n = K.shape[0]
    mask = np.zeros(n, bool); mask[free] = True
    Dm = sp.diags(mask.astype(float))
    A = (Dm @ K @ Dm + sp.diags((~mask).astype(float))).tocsr()
    lp, ln, ld, lf = api.symbolic_lines(mesh.points, conn)
    Kb = K.tobsr((6, 6))
    Kb.sort_indices()
    rowptr, colidx = api.symbolic_pattern(len(mesh.points), conn)
    assert np.array_equal(rowptr, Kb.indptr) and np.array_equal(colidx, Kb.indices)
    Kf, row_in_family, Dseg = LS.galerkin(rowptr, colidx, Kb.data, mask, lp, ln, ld, lf)
    # matrix form
    rows, cols, vals, srows, scols = [], [], [], [], []
    seg_first, n_seg = LS.segments(lp)
    for k in range(len(lf)):
        nodes = ln[lp[k]:lp[k + 1]]
        for c in range(3):
            rows.append(6 * nodes + c); cols.append(np.full(len(nodes), k)); vals.append(np.full(len(nodes), ld[k, c]))
            scols.append(seg_first[k] + np.arange(len(nodes)) // LS.SEG)
    Pl = (Dm @ sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, len(lf)))).tocsr()
    Ps = (Dm @ sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(scols))), shape=(n, n_seg))).tocsr()
    G = (Pl.T @ A @ Pl).toarray()
    dead = np.diag(G) <= 0.0                 # lines inside the fixed base: identity rows, as in the spec
    G[dead, dead] = 1.0
    for fam in range(3):
        idx = np.flatnonzero(lf == fam)
        idx = idx[np.argsort(row_in_family[idx])]
        ref = G[np.ix_(idx, idx)]
        assert np.abs(Kf[fam] - ref).max() <= 1e-12 * np.abs(ref).max()
    dref = (Ps.T @ A @ Ps).diagonal()
    assert np.abs(Dseg - dref).max() <= 1e-12 * np.abs(dref).max()
    Kf_inv = [np.linalg.inv(m) for m in Kf]
    rng = np.random.default_rng(7)
    r = rng.standard_normal(n) * mask
    d = A.diagonal()
    z = LS.apply(r, 1.0 / d, 2.0, mask, lp, ln, ld, lf, Kf_inv, row_in_family, Dseg)
    Gbd = np.zeros_like(G)
    for fam in range(3):
        idx = np.flatnonzero(lf == fam)
        Gbd[np.ix_(idx, idx)] = G[np.ix_(idx, idx)]
    zref = 2.0 * r / d + Pl @ np.linalg.solve(Gbd, Pl.T @ r) + Ps @ ((Ps.T @ r) / np.where(dref > 0, dref, 1.0))
    assert np.abs(z - zref).max() <= 1e-10 * np.abs(zref).max()


def test_qr_algorithm_helper_has_no_cpu_fallback():
    """BeamSolver.py:467 keeps its signature on the mirror class and runs on the GPU (tests/test_gpu_frame.py compares
    it with the unmodified reference's output); without a device it raises instead of computing on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the GPU test")
    w = compat.BeamAnalysisB200.__new__(compat.BeamAnalysisB200)
    w.device = 0
    with pytest.raises(RuntimeError, match="no CUDA device"):
        w.qr_algorithm(np.eye(3))


def test_result_tables_follow_the_reference_report(tmp_path):
    """BeamSolver.py:530-547: nodal table (node, x, y, z, u_x, u_y, u_z, stress in MPa) and the
    first-10 frequency table (mode, rad/s, Hz), as arrays and CSV with the report's formats."""
    w = compat.BeamAnalysisB200.__new__(compat.BeamAnalysisB200)
    w.points = np.array([[0.0, 0.0, 0.0], [1.0, 0.5, 0.25]])
    w.u = np.arange(12, dtype=float) * 1e-3
    w.smoothed_stresses = np.array([2.5e6, 48e6])
    w.natural_frequencies = np.linspace(10.0, 130.0, 13)
    t = w.nodal_result_table()
    assert t.shape == (2, 8) and np.allclose(t[1], [1, 1.0, 0.5, 0.25, 6e-3, 7e-3, 8e-3, 48.0])
    m = w.modal_result_table()
    assert m.shape == (10, 3) and np.allclose(m[0], [1, 10.0, 10.0 / (2 * np.pi)])
    paths = w.write_result_tables(str(tmp_path / "run"))
    nodes = open(paths[0]).read().splitlines()
    assert nodes[0].startswith("Node ID,X (m)") and nodes[2] == "1,1.0000,0.5000,0.2500,6.0000e-03,7.0000e-03,8.0000e-03,48.0000"
    modes = open(paths[1]).read().splitlines()
    assert len(modes) == 11 and modes[1] == "1,10.0000,1.5915"
    w.natural_frequencies = None
    assert w.modal_result_table().shape == (0, 3) and len(w.write_result_tables(str(tmp_path / "static"))) == 1


# ---- closed-form section front end: value tests for all seven dialog types ------------------------
def _polygon_props(poly):
    """Area, centroid and centroidal second moments of a simple polygon (exact: Green's theorem)."""
    x, y = np.asarray(poly, dtype=float).T
    x1, y1 = np.roll(x, -1), np.roll(y, -1)
    c = x * y1 - x1 * y
    A = c.sum() / 2
    cx = ((x + x1) * c).sum() / (6 * A)
    cy = ((y + y1) * c).sum() / (6 * A)
    ixx = ((y ** 2 + y * y1 + y1 ** 2) * c).sum() / 12 - A * cy ** 2
    iyy = ((x ** 2 + x * x1 + x1 ** 2) * c).sum() / 12 - A * cx ** 2
    s = 1.0 if A > 0 else -1.0
    return s * A, cx, cy, s * ixx, s * iyy


def _shape_polygons(kind, p):
    """(outer polygon, hole polygon or None) of the sharp-cornered dialog shapes, picture axes (x horizontal)."""
    if kind == "rectangular section":
        d, b = p["d"], p["b"]
        return [(0, 0), (b, 0), (b, d), (0, d)], None
    if kind == "I section":
        d, b, tf, tw = p["d"], p["b"], p["t_f"], p["t_w"]
        x0, x1 = (b - tw) / 2, (b + tw) / 2
        return [(0, 0), (b, 0), (b, tf), (x1, tf), (x1, d - tf), (b, d - tf), (b, d), (0, d), (0, d - tf), (x0, d - tf), (x0, tf), (0, tf)], None
    if kind == "C section":
        d, b, tf, tw = p["d"], p["b"], p["t_f"], p["t_w"]
        return [(0, 0), (b, 0), (b, tf), (tw, tf), (tw, d - tf), (b, d - tf), (b, d), (0, d)], None
    if kind == "L section":
        d, b, t = p["d"], p["b"], p["t"]
        return [(0, 0), (b, 0), (b, t), (t, t), (t, d), (0, d)], None
    if kind == "hollow box section":
        d, b, t = p["d"], p["b"], p["t"]
        return [(0, 0), (b, 0), (b, d), (0, d)], [(t, t), (b - t, t), (b - t, d - t), (t, d - t)]
    raise KeyError(kind)


@pytest.mark.parametrize("kind,params", [
    ("rectangular section", {"d": 0.1, "b": 0.05}),
    ("I section", {"d": 0.2, "b": 0.1, "t_f": 0.0085, "t_w": 0.0056, "r": 0.0}),
    ("C section", {"d": 0.1, "b": 0.05, "t_f": 0.005, "t_w": 0.005, "r": 0.0}),
    ("L section", {"d": 0.075, "b": 0.075, "t": 0.006, "r_r": 0.0, "r_t": 0.0}),
    ("L section", {"d": 0.12, "b": 0.07, "t": 0.008, "r_r": 0.0, "r_t": 0.0}),
    ("hollow box section", {"d": 0.1, "b": 0.1, "t": 0.005, "r_out": 0.0}),
    ("hollow box section", {"d": 0.15, "b": 0.08, "t": 0.006, "r_out": 0.0}),
])
def test_polygonal_sections_match_exact_polygon_integrals(kind, params):
    """A, I_x, I_y and the extreme fibre distances of the five polygonal dialog shapes against exact polygon integrals;
    `rotate` swaps the (I, kappa, c) pairs (BeamSolver.py:73-79)."""
    from fem_calculator_b200.sections import calculate_section_properties as csp
    outer, hole = _shape_polygons(kind, params)
    A, cx, cy, ixx, iyy = _polygon_props(outer)
    if hole:
        Ah, hx, hy, hxx, hyy = _polygon_props(hole)
        ixx = ixx + A * cy ** 2 - (hxx + Ah * hy ** 2)
        iyy = iyy + A * cx ** 2 - (hyy + Ah * hx ** 2)
        cx, cy = (A * cx - Ah * hx) / (A - Ah), (A * cy - Ah * hy) / (A - Ah)
        A -= Ah
        ixx -= A * cy ** 2
        iyy -= A * cx ** 2
    xs, ys = np.asarray(outer, dtype=float).T
    got = csp(kind, params, False)
    assert abs(got[0] - A) <= 1e-12 * A
    assert abs(got[1] - ixx) <= 1e-11 * ixx and abs(got[2] - iyy) <= 1e-11 * iyy, (got, ixx, iyy)
    assert abs(got[6] - max(cx - xs.min(), xs.max() - cx)) <= 1e-12 and abs(got[7] - max(cy - ys.min(), ys.max() - cy)) <= 1e-12
    assert 0.0 < got[4] < 1.0 and 0.0 < got[5] < 1.0 and got[3] > 0.0
    rot = csp(kind, params, True)
    assert rot[1] == got[2] and rot[2] == got[1] and rot[4] == got[5] and rot[5] == got[4] and rot[6] == got[7] and rot[7] == got[6]


def test_section_torsion_constants_and_round_sections_match_published_values():
    """J: exact for the circle (pi d^4 / 32) and the tube; Roark's table for the solid rectangle (beta = 0.229 at
    a/b = 2, 0.141 for the square); Bredt's formula for the closed box; sum(b t^3 / 3) for the open thin-walled shapes."""
    from fem_calculator_b200.sections import calculate_section_properties as csp
    d = 0.08
    c = csp("circular section", {"d": d})
    assert abs(c[0] - np.pi * d * d / 4) <= 1e-15 and abs(c[1] - np.pi * d ** 4 / 64) <= 1e-18 and c[1] == c[2]
    assert abs(c[3] - np.pi * d ** 4 / 32) <= 1e-18 and abs(c[4] - 6 / 7) <= 1e-12 and c[6] == d / 2
    t = 0.004
    h = csp("hollow circular section", {"d": d, "t": t})
    di = d - 2 * t
    assert abs(h[0] - np.pi * (d * d - di * di) / 4) <= 1e-15 and abs(h[3] - np.pi * (d ** 4 - di ** 4) / 32) <= 1e-18
    assert 0.5 < h[4] < 6 / 7                       # thin tube -> 1/2, solid -> 6/7 (Cowper, nu = 0)
    r = csp("rectangular section", {"d": 0.1, "b": 0.05})
    assert abs(r[3] / (0.1 * 0.05 ** 3) - 0.229) <= 1e-3 and abs(r[4] - 5 / 6) <= 1e-12
    sq = csp("rectangular section", {"d": 0.06, "b": 0.06})
    assert abs(sq[3] / 0.06 ** 4 - 0.1406) <= 1e-3
    bx = csp("hollow box section", {"d": 0.15, "b": 0.08, "t": 0.006, "r_out": 0.0})
    Am = (0.15 - 0.006) * (0.08 - 0.006)
    assert abs(bx[3] - 4 * Am ** 2 / (2 * ((0.15 - 0.006) + (0.08 - 0.006)) / 0.006)) <= 1e-12 * bx[3]
    i = csp("I section", {"d": 0.2, "b": 0.1, "t_f": 0.0085, "t_w": 0.0056, "r": 0.0})
    assert abs(i[3] - (2 * 0.1 * 0.0085 ** 3 + (0.2 - 0.0085) * 0.0056 ** 3) / 3) <= 1e-18
    ell = csp("L section", {"d": 0.075, "b": 0.075, "t": 0.006, "r_r": 0.0, "r_t": 0.0})
    assert abs(ell[3] - (0.075 + 0.075 - 0.006) * 0.006 ** 3 / 3) <= 1e-18


def test_fillet_radii_are_reported_not_silently_dropped():
    from fem_calculator_b200.sections import calculate_section_properties as csp
    with pytest.warns(RuntimeWarning, match="fillet"):
        a = csp("I section", {"d": 0.2, "b": 0.1, "t_f": 0.0085, "t_w": 0.0056, "r": 0.012})
    b = csp("I section", {"d": 0.2, "b": 0.1, "t_f": 0.0085, "t_w": 0.0056, "r": 0.0})
    assert a == b


def test_msh_reader_rejects_other_format_versions(tmp_path):
    from fem_calculator_b200 import msh
    p = tmp_path / "old.msh"
    p.write_text("$MeshFormat\n4.0 0 8\n$EndMeshFormat\n")
    with pytest.raises(ValueError, match="4.1"):
        msh.read_msh(str(p))


def test_accelerated_force_analysis_keeps_the_reference_bookkeeping():
    """compat.accelerate_force_analysis: the subclass inherits the reference's apply_boundary_conditions /
    print_reactions and only adds the device hand-over (here with a stand-in model: no GPU)."""
    from fem_calculator_b200 import compat

    class Ref:                                       # the shape of ReactionSolver.ForceAnalysis
        def __init__(self):
            self.calls = []

        def apply_boundary_conditions(self):
            self.calls.append("ref-bc")
            self.fixed_dofs = np.array([0, 1, 2])
            self.f = np.arange(9.0)

        def print_reactions(self):
            self.calls.append("ref-print")

    class Model:
        def set_bc(self, fixed, f):
            self.got = (fixed, f)

    G = compat.accelerate_force_analysis(Ref)
    g = G()
    g._model = Model()
    g.apply_boundary_conditions()
    g.print_reactions()
    assert g.calls == ["ref-bc", "ref-print"] and g._model.got[0].dtype == np.int64 and np.array_equal(g._model.got[1], np.arange(9.0))
    assert G.__name__ == "RefGPU" and G.solve is compat.ForceAnalysisB200.solve
