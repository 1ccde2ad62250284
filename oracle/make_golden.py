"""TEST INFRASTRUCTURE ONLY.  Generates the committed golden vectors under
``tests/golden/`` by executing the UNMODIFIED reference (``/root/reference``) through
``oracle/ref_harness.py``.  Run in the build container only:

    python -m oracle.make_golden

Every fixture stores its full inputs next to the reference's outputs, so the GPU box
(which has no /root/reference) can replay them.  Section records come from
``fem_calculator_b200.sections`` (sectionproperties is not installed); they are inputs,
stored verbatim in the fixture.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from fem_calculator_b200 import meshgen, msh  # noqa: E402
from fem_calculator_b200.sections import calculate_section_properties as csp  # noqa: E402
from oracle import ref_harness as H  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
E, NU = 2.0e11, 0.3


def _mesh_arrays(mesh, line=True):
    d = {"points": mesh.points, "vertex": mesh.cells_dict["vertex"],
         "vertex_phys": mesh.cell_data_dict["gmsh:physical"]["vertex"],
         "field_json": np.frombuffer(json.dumps({k: [int(x) for x in v] for k, v in mesh.field_data.items()}).encode(), dtype=np.uint8)}
    if line:
        d["line"] = mesh.cells_dict["line"]
        d["line_phys"] = mesh.cell_data_dict["gmsh:physical"]["line"]
    else:
        d["tetra10"] = mesh.cells_dict["tetra10"]
        d["tetra10_phys"] = mesh.cell_data_dict["gmsh:physical"]["tetra10"]
    return d


def _beam_fixture(name, mesh, sec, bc, eb=False, dense=True):
    props = {s["group"]: csp(s["type"], s["params"], s.get("rotate", False)) for s in sec}
    if eb:
        props = {k: meshgen.euler_bernoulli(v) for k, v in props.items()}
    ref = H.run_beam_reference(mesh, props, bc, E, NU, with_modal=True)
    d = _mesh_arrays(mesh)
    d["props_names_json"] = np.frombuffer(json.dumps(list(props.keys())).encode(), dtype=np.uint8)
    d["props"] = np.asarray([props[k] for k in props], dtype=np.float64)
    d["bc_json"] = np.frombuffer(json.dumps(bc).encode(), dtype=np.uint8)
    d["E"], d["nu"] = np.float64(E), np.float64(NU)
    d["ref_u"] = ref["u"]
    d["ref_smoothed_stresses"] = ref["smoothed_stresses"]
    d["ref_natural_frequencies"] = ref["natural_frequencies"]
    d["ref_mode_shapes"] = ref["mode_shapes"]
    if dense:
        K, M = H.beam_reference_matrices(mesh, props, E, NU)
        d["ref_K_dense"], d["ref_M_dense"] = K, M
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(f"{name}: {len(mesh.points)} nodes, {len(mesh.cells_dict['line'])} elements, "
          f"{len(ref['natural_frequencies'])} reference modes")


def _element_fixture():
    """Element-level vectors straight from the reference's helpers
    (BeamSolver.py:646-675) plus the rotation of :378-388, for awkward directions."""
    kfun, mfun, _ = H.beam_reference_helpers()
    rng = np.random.default_rng(meshgen.SEED)
    dirs = [(1, 0, 0), (0, 1, 0), (0, 0, 1), (0, 0, -1), (1, 1, 1), (-1, 2, -3), (1e-7, 0, 1), (2e-6, 0, 1),
            (0, 5e-7, -1), (3, -4, 0)] + [tuple(v) for v in rng.normal(size=(22, 3))]
    recs = []
    for i, dvec in enumerate(dirs):
        p1 = rng.normal(size=3)
        L = rng.uniform(0.05, 3.0)
        dvec = np.asarray(dvec, dtype=float)
        p2 = p1 + L * dvec / np.linalg.norm(dvec)
        A, Ix, Iy, J = rng.uniform(1e-4, 1e-2), rng.uniform(1e-8, 1e-5), rng.uniform(1e-8, 1e-5), rng.uniform(1e-8, 1e-5)
        ky, kz = (0.0, 0.0) if i % 5 == 4 else (rng.uniform(0.2, 0.9), rng.uniform(0.2, 0.9))
        recs.append((p1, p2, (A, Ix, Iy, J, ky, kz, 0.01, 0.02)))
    pts = np.concatenate([np.stack([r[0], r[1]]) for r in recs])
    conn = np.arange(2 * len(recs)).reshape(-1, 2)
    props = np.asarray([r[2] for r in recs])
    G = E / (2 * (1 + NU))
    ke, me, kl = [], [], []
    for (p1, p2, pr) in recs:
        A, Ix, Iy, J, ky, kz, _, _ = pr
        # verbatim replay of BeamSolver.py:373-388 through the reference's helpers
        L = np.linalg.norm(p2 - p1)
        k_ = kfun(L, E, G, A, Ix, Iy, J, ky, kz)
        m_ = mfun(L, A, Ix, Iy, J, 7850)
        Cxx, Cyx, Czx = (p2 - p1) / L
        if Cxx**2 + Cyx**2 < 1e-6**2:
            lam = np.array([[0., 0., 1. if Czx > 0 else -1.], [0., 1., 0.], [-1. if Czx > 0 else 1., 0., 0.]])
        else:
            D = np.sqrt(Cxx**2 + Cyx**2)
            lam = np.array([[Cxx, Cyx, Czx], [-Cyx / D, Cxx / D, 0], [-Cxx * Czx / D, -Cyx * Czx / D, D]])
        R = np.kron(np.eye(4, dtype=float), lam)
        kl.append(k_)
        ke.append(R.T @ k_ @ R)
        me.append(R.T @ m_ @ R)
    np.savez_compressed(os.path.join(OUT, "frame_elements.npz"), points=pts, line=conn,
                        props=props, elem_sec=np.arange(len(recs), dtype=np.int32), E=np.float64(E), nu=np.float64(NU),
                        ref_k_local=np.asarray(kl), ref_ke_global=np.asarray(ke), ref_me_global=np.asarray(me))
    print(f"frame_elements: {len(recs)} elements")


def _tet_fixture(name, nx, ny, nz, store_K=True):
    mesh, fd, xd = meshgen.tet10_box_case(nx, ny, nz)
    fa = H.run_tet10_reference(mesh, fd, xd, E, NU)
    d = _mesh_arrays(mesh, line=False)
    d["force_json"] = np.frombuffer(json.dumps(fd).encode(), dtype=np.uint8)
    d["fix_json"] = np.frombuffer(json.dumps(xd).encode(), dtype=np.uint8)
    d["E"], d["nu"] = np.float64(E), np.float64(NU)
    K = fa.K.tocsr()
    K.sort_indices()
    if store_K:
        d["ref_K_indptr"], d["ref_K_indices"], d["ref_K_data"] = K.indptr, K.indices, K.data
    d["ref_negative_detJ_count"] = np.int64(fa.negative_detJ_count)
    d["ref_fixed_dofs"], d["ref_active_dofs"] = fa.fixed_dofs, fa.active_dofs
    d["ref_f"], d["ref_u"], d["ref_reaction_forces"] = fa.f, fa.u, fa.reaction_forces
    d["ref_fixed_nodes"] = np.asarray([i["node_idx"] for i in fa.fixed_nodes_info])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(f"{name}: {len(mesh.points)} nodes, {len(mesh.cells_dict['tetra10'])} tets")


def main():
    if not H.available():
        raise SystemExit("reference not present; golden vectors can only be generated in the build container")
    os.makedirs(OUT, exist_ok=True)
    # C1: the shipped mesh file, rectangular 0.1 x 0.05, tip load -1000 N
    mesh = msh.read_msh(os.path.join(H.REFERENCE_DIR, "cantilever_beam"))
    sec = [{"group": "beam", "type": "rectangular section", "params": {"d": 0.1, "b": 0.05}, "rotate": False}]
    bc = [meshgen._fix_bc("fix"), meshgen._force_bc("load_y", fy=-1000.0)]
    _beam_fixture("c1_cantilever_beam", mesh, sec, bc)
    # C2 slice: simply supported Euler-Bernoulli I-beam, 24 elements
    mesh, sec, bc = meshgen.simply_supported_case(24, 10.0)
    _beam_fixture("c2_simply_supported_24", mesh, sec, bc, eb=True)
    # C3 slices: 3x3x4 lattice, axis-aligned (vertical branch) and jittered (generic rotation)
    mesh, sec, bc = meshgen.lattice_frame_case(3, 3, 4, jitter=0.0)
    _beam_fixture("c3_lattice_3x3x4_aligned", mesh, sec, bc)
    mesh, sec, bc = meshgen.lattice_frame_case(3, 3, 4, jitter=0.05)
    _beam_fixture("c3_lattice_3x3x4_jitter", mesh, sec, bc)
    # inclined cantilever chain (generic direction, Timoshenko), 8 elements
    mesh = meshgen.chain_mesh(8, 3.0, {"fix": [0], "load_y": [8]}, axis=(1.0, 2.0, -0.5))
    sec = [{"group": "beam", "type": "hollow circular section", "params": {"d": 0.12, "t": 0.008}, "rotate": False}]
    bc = [meshgen._fix_bc("fix"), meshgen._force_bc("load_y", fx=50.0, fy=-700.0, fz=120.0)]
    _beam_fixture("chain_inclined_8", mesh, sec, bc)
    _element_fixture()
    _tet_fixture("tet10_box_2x1x2", 2, 1, 2, store_K=True)
    _tet_fixture("tet10_box_4x1x4", 4, 1, 4, store_K=False)


if __name__ == "__main__":
    main()
