"""TEST INFRASTRUCTURE ONLY — CPU restatement ("port") of the reference's numerical
hot path.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module; the product path
(``fem_calculator_b200``) never does and fails loudly without its CUDA library.

What it restates (all citations into /root/reference):
  * frame element stiffness / lumped mass       BeamSolver.py:646-660, 662-675
  * element geometry, rotation, scatter         BeamSolver.py:364-393
  * BC bookkeeping, load vector, static solve   BeamSolver.py:395-418, 677-686
  * stress recovery                             BeamSolver.py:420-438
  * modal pencil K_ff phi = lambda M_ff phi     BeamSolver.py:440-455 (see note)
  * Tet10 material, shape fns, assembly         ReactionSolver.py:87-152
  * Tet10 BC, solve, reactions                  ReactionSolver.py:154-205

The reference builds DENSE (6N,6N) matrices and solves with LAPACK; this restatement is
the same arithmetic vectorised over elements and stored sparse (scipy CSR), so that it
runs at the BASELINE.json sizes and serves as the CPU baseline.  Third-party arithmetic
it stands on (not vendored by the reference, no pinned versions, Dependencies.txt:1-7):
numpy.linalg.solve/inv/qr/det (LAPACK), scipy.sparse lil->csr, scipy.sparse.linalg.spsolve
(SuperLU); here numpy 2.3 / scipy 1.18.

Parity pinning: the reference ships NO tests and NO golden vectors.
This oracle is pinned instead against (1) outputs of the unmodified reference executed
in the build container through ``oracle/ref_harness.py`` — committed as fixtures under
``tests/golden/`` by ``oracle/make_golden.py`` — and (2) closed-form known answers on
the shipped ``cantilever_beam`` mesh.  tests/test_oracle.py checks both.

Modal note: the reference's ``qr_algorithm`` (BeamSolver.py:467-481) is an unshifted QR
iteration on inv(M_ff) K_ff stopped at ~1e-5 relative change; its eigenvalues carry
1e-9..1e-3 errors and its "mode shapes" are Schur vectors.  The pencil it
approximates is K_ff phi = lambda M_ff phi, which is what ``frame_modal`` solves; the
reference's eigenvalues are used as a loose cross-check only.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp
import scipy.sparse.linalg as spla

RHO_REFERENCE = 7850.0      # literal at BeamSolver.py:376 (the UI density field is ignored)
EPS_VERTICAL = 1e-6         # BeamSolver.py:362
DETJ_MIN = 1e-12            # ReactionSolver.py:133
GAUSS_A, GAUSS_B = 0.58541020, 0.13819660   # 8-digit literals, ReactionSolver.py:120-123
GAUSS_W = 1 / 4             # ReactionSolver.py:124 (no 1/6 reference-volume factor)


# ------------------------------------------------------------------ frame elements

def timoshenko_local_stiffness(L, E, G, A, I_x, I_y, J, kappa_y, kappa_z):
    """Vectorised BeamSolver.py:646-660.  All arguments broadcast to (n,).  Returns
    (n,12,12).  phi_* = 0 when its denominator is <= 0 (Euler-Bernoulli fallback,
    :647-648); every term is 0 when L <= 0."""
    L, A, I_x, I_y, J, kappa_y, kappa_z = np.broadcast_arrays(
        *(np.asarray(v, dtype=np.float64) for v in (L, A, I_x, I_y, J, kappa_y, kappa_z)))
    n = L.shape[0]
    with np.errstate(divide="ignore", invalid="ignore"):
        den_z = G * kappa_y * A * L**2
        den_y = G * kappa_z * A * L**2
        phi_z = np.where(den_z > 0, (12 * E * I_y) / den_z, 0.0)
        phi_y = np.where(den_y > 0, (12 * E * I_x) / den_y, 0.0)
        ok = L > 0
        z = lambda v: np.where(ok, v, 0.0)
        k11_z = z((12 * E * I_y) / (L**3 * (1 + phi_z)))
        k12_z = z((6 * E * I_y) / (L**2 * (1 + phi_z)))
        k22_z = z(((4 + phi_z) * E * I_y) / (L * (1 + phi_z)))
        k23_z = z(((2 - phi_z) * E * I_y) / (L * (1 + phi_z)))
        k11_y = z((12 * E * I_x) / (L**3 * (1 + phi_y)))
        k12_y = z((6 * E * I_x) / (L**2 * (1 + phi_y)))
        k22_y = z(((4 + phi_y) * E * I_x) / (L * (1 + phi_y)))
        k23_y = z(((2 - phi_y) * E * I_x) / (L * (1 + phi_y)))
        tor = z(G * J / L)
        ax = z(A * E / L)
    k = np.zeros((n, 12, 12))
    def put(entries):
        for (r, c, v) in entries:
            k[:, r, c] = v
    put([(0, 0, ax), (0, 6, -ax), (6, 0, -ax), (6, 6, ax)])
    put([(3, 3, tor), (3, 9, -tor), (9, 3, -tor), (9, 9, tor)])
    # bending in local x-y plane (uses I_y), rows 1,5,7,11            :655-660
    put([(1, 1, k11_z), (1, 5, k12_z), (1, 7, -k11_z), (1, 11, k12_z),
         (5, 1, k12_z), (5, 5, k22_z), (5, 7, -k12_z), (5, 11, k23_z),
         (7, 1, -k11_z), (7, 5, -k12_z), (7, 7, k11_z), (7, 11, -k12_z),
         (11, 1, k12_z), (11, 5, k23_z), (11, 7, -k12_z), (11, 11, k22_z)])
    # bending in local x-z plane (uses I_x), rows 2,4,8,10
    put([(2, 2, k11_y), (2, 4, -k12_y), (2, 8, -k11_y), (2, 10, -k12_y),
         (4, 2, -k12_y), (4, 4, k22_y), (4, 8, k12_y), (4, 10, k23_y),
         (8, 2, -k11_y), (8, 4, k12_y), (8, 8, k11_y), (8, 10, k12_y),
         (10, 2, -k12_y), (10, 4, k23_y), (10, 8, k12_y), (10, 10, k22_y)])
    return k


def lumped_local_mass(L, A, Ix, Iy, J, rho):
    """Vectorised BeamSolver.py:662-675: diagonal (n,12,12)."""
    L, A, Ix, Iy, J = np.broadcast_arrays(*(np.asarray(v, dtype=np.float64) for v in (L, A, Ix, Iy, J)))
    m = np.zeros((L.shape[0], 12, 12))
    tm = rho * A * L / 2
    rx = rho * J * L / 2
    ry = rho * Ix * L / 2
    rz = rho * Iy * L / 2
    for o in (0, 6):
        m[:, o, o] = m[:, o + 1, o + 1] = m[:, o + 2, o + 2] = tm
        m[:, o + 3, o + 3] = rx
        m[:, o + 4, o + 4] = ry
        m[:, o + 5, o + 5] = rz
    return m


def frame_rotation(points, conn):
    """L and the 3x3 direction-cosine matrix per element, BeamSolver.py:372-384."""
    p1 = points[conn[:, 0]]
    p2 = points[conn[:, 1]]
    d = p2 - p1
    L = np.sqrt((d * d).sum(axis=1))           # np.linalg.norm: sqrt(sum of squares)
    with np.errstate(divide="ignore", invalid="ignore"):
        c = d / L[:, None]
        cx, cy, cz = c[:, 0], c[:, 1], c[:, 2]
        vert = cx**2 + cy**2 < EPS_VERTICAL**2
        D = np.sqrt(cx**2 + cy**2)
        lam = np.zeros((len(conn), 3, 3))
        lam[:, 0, 0], lam[:, 0, 1], lam[:, 0, 2] = cx, cy, cz
        lam[:, 1, 0], lam[:, 1, 1] = -cy / D, cx / D
        lam[:, 2, 0], lam[:, 2, 1], lam[:, 2, 2] = -cx * cz / D, -cy * cz / D, D
    s = np.where(cz > 0, 1.0, -1.0)
    lv = np.zeros((len(conn), 3, 3))
    lv[:, 0, 2] = s
    lv[:, 1, 1] = 1.0
    lv[:, 2, 0] = -s
    lam[vert] = lv[vert]
    return L, lam


def frame_element_matrices(points, conn, elem_sec, props, E, nu, rho=RHO_REFERENCE):
    """Global-axis element matrices K_e = R^T k R, M_e = R^T m R (BeamSolver.py:375-388).
    props: (S,8) section records; elem_sec: (E,) index into props.  Returns (ke, me)
    each (E,12,12)."""
    G = E / (2 * (1 + nu))
    pr = np.asarray(props, dtype=np.float64)[np.asarray(elem_sec)]
    A, I_x, I_y, J, ky, kz = (pr[:, i] for i in range(6))
    L, lam = frame_rotation(np.asarray(points, dtype=np.float64), np.asarray(conn))
    k_ = timoshenko_local_stiffness(L, E, G, A, I_x, I_y, J, ky, kz)
    m_ = lumped_local_mass(L, A, I_x, I_y, J, rho)
    R = np.zeros((len(L), 12, 12))
    for b in range(4):
        R[:, 3 * b:3 * b + 3, 3 * b:3 * b + 3] = lam
    Rt = R.transpose(0, 2, 1)
    return Rt @ k_ @ R, Rt @ m_ @ R


def _scatter_blocks(conn, emat, n_nodes, bs, nper):
    """COO -> CSR scatter of element matrices with node-major DOFs (dof = bs*node + c);
    a structural diagonal block is added for every node so isolated points keep their
    (all-zero) rows like the reference's dense matrix does (BeamSolver.py:354,360)."""
    ne = len(conn)
    dofs = (bs * np.repeat(conn, bs, axis=1) + np.tile(np.arange(bs), nper)[None, :])  # (E, nper*bs)
    nd = nper * bs
    rows = np.repeat(dofs, nd, axis=1).ravel()
    cols = np.tile(dofs, (1, nd)).ravel()
    vals = np.asarray(emat).reshape(ne, nd * nd).ravel()
    dn = bs * np.repeat(np.arange(n_nodes), bs * bs) + np.tile(np.repeat(np.arange(bs), bs), n_nodes)
    dc = bs * np.repeat(np.arange(n_nodes), bs * bs) + np.tile(np.tile(np.arange(bs), bs), n_nodes)
    rows = np.concatenate([rows, dn])
    cols = np.concatenate([cols, dc])
    vals = np.concatenate([vals, np.zeros(len(dn))])
    n = bs * n_nodes
    K = sp.coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsr()
    K.sum_duplicates()
    K.sort_indices()
    return K


def frame_assemble(points, conn, elem_sec, props, E, nu, rho=RHO_REFERENCE):
    """Sparse equivalent of the scatter at BeamSolver.py:390-393.  Returns CSR (K, M) on
    the STRUCTURAL pattern (6x6 blocks over node adjacency + every node's diagonal
    block; explicit zeros kept)."""
    ke, me = frame_element_matrices(points, conn, elem_sec, props, E, nu, rho)
    n = len(points)
    return _scatter_blocks(np.asarray(conn), ke, n, 6, 2), _scatter_blocks(np.asarray(conn), me, n, 6, 2)


def frame_bc(mesh, bc_data):
    """BeamSolver.py:395-410: fixed DOF list (sorted unique), free list (ascending
    complement) and the nodal load vector.  Every node of a Force group receives the
    full force vector; rotational loads do not exist."""
    n = 6 * len(mesh.points)
    f = np.zeros(n)
    up = []
    for bc in bc_data:
        nodes = mesh_group_nodes(mesh, "vertex", bc["group"])
        if bc["type"] == "Fix":
            for c, key in enumerate(("fix_x", "fix_y", "fix_z", "fix_rx", "fix_ry", "fix_rz")):
                if bc.get(key):
                    up.append(6 * nodes + c)
        elif bc["type"] == "Force":
            f[6 * nodes + 0] += bc.get("force_x", 0)
            f[6 * nodes + 1] += bc.get("force_y", 0)
            f[6 * nodes + 2] += bc.get("force_z", 0)
    fixed = np.unique(np.concatenate(up)) if up else np.zeros(0, dtype=np.int64)
    mask = np.ones(n, dtype=bool)
    mask[fixed] = False
    free = np.nonzero(mask)[0]
    return fixed.astype(np.int64), free.astype(np.int64), f


def mesh_group_nodes(mesh, element_type, name):
    """BeamSolver.py:677-686 / ReactionSolver.py:75-85."""
    cells = mesh.cells_dict.get(element_type)
    phys = mesh.cell_data_dict.get("gmsh:physical", {}).get(element_type)
    if cells is None or phys is None or name not in mesh.field_data:
        return np.zeros(0, dtype=np.int64)
    return np.unique(np.asarray(cells)[np.asarray(phys) == mesh.field_data[name][0]].ravel())


def solve_static(K, f, fixed, free, method="auto", rtol=1e-13):
    """u[free] = K_ff^-1 f_f, u[fixed] = 0 (BeamSolver.py:412-418; ReactionSolver.py:199-203).
    'direct' = SuperLU (the reference's Tet10 solver), 'cg' = Jacobi-PCG (the scalable
    CPU baseline)."""
    Kff = K[free][:, free].tocsc()
    ff = f[free]
    if method == "auto":
        method = "direct" if Kff.shape[0] <= 150_000 else "cg"
    info = {"method": method, "iterations": 0}
    if method == "direct":
        uf = spla.spsolve(Kff, ff)
    else:
        d = Kff.diagonal()
        Minv = spla.LinearOperator(Kff.shape, matvec=lambda x: x / d)
        it = [0]
        uf, flag = spla.cg(Kff.tocsr(), ff, rtol=rtol, atol=0.0, M=Minv, maxiter=200_000,
                           callback=lambda xk: it.__setitem__(0, it[0] + 1))
        info["iterations"] = it[0]
        info["flag"] = flag
    u = np.zeros(K.shape[0])
    u[free] = uf
    return u, info


def reactions(K, u, f=None):
    """ReactionSolver.py:205: reaction_forces = K_full @ u (f is NOT subtracted there;
    pass f to get the support reaction K u - f used for the frame path)."""
    r = K @ u
    return r if f is None else r - f


def frame_stress(points, conn, elem_sec, props, E, nu, u):
    """BeamSolver.py:420-438: f_local = k_ (R u_e); sigma = f_local[6]/A +
    |M c / I| at each end; node-averaged.  Nodes touched by no element keep 0."""
    G = E / (2 * (1 + nu))
    conn = np.asarray(conn)
    pr = np.asarray(props, dtype=np.float64)[np.asarray(elem_sec)]
    A, I_x, I_y, J, ky, kz, cy, cz = (pr[:, i] for i in range(8))
    L, lam = frame_rotation(np.asarray(points, dtype=np.float64), conn)
    k_ = timoshenko_local_stiffness(L, E, G, A, I_x, I_y, J, ky, kz)
    ue = np.concatenate([u.reshape(-1, 6)[conn[:, 0]], u.reshape(-1, 6)[conn[:, 1]]], axis=1)  # (E,12)
    ul = np.einsum("eij,ebj->ebi", lam, ue.reshape(-1, 4, 3)).reshape(-1, 12)
    fl = np.einsum("eij,ej->ei", k_, ul)
    with np.errstate(divide="ignore", invalid="ignore"):
        sa = np.where(A > 0, fl[:, 6] / A, 0.0)
        b1 = np.abs(np.where(I_x > 0, fl[:, 4] * cz / I_x, 0.0)) + np.abs(np.where(I_y > 0, fl[:, 5] * cy / I_y, 0.0))
        b2 = np.abs(np.where(I_x > 0, fl[:, 10] * cz / I_x, 0.0)) + np.abs(np.where(I_y > 0, fl[:, 11] * cy / I_y, 0.0))
    n = len(points)
    s = np.zeros(n)
    cnt = np.zeros(n)
    np.add.at(s, conn[:, 0], sa + b1)
    np.add.at(s, conn[:, 1], sa + b2)
    np.add.at(cnt, conn[:, 0], 1)
    np.add.at(cnt, conn[:, 1], 1)
    out = np.zeros(n)
    np.divide(s, cnt, out=out, where=cnt != 0)
    return out


def frame_modal(K, M, free, k=20, dense_limit=3000):
    """Lowest-k eigenpairs of K_ff phi = lambda M_ff phi (the pencil behind
    BeamSolver.py:440-455), lambda > 1e-6 kept (:448), omega = sqrt(lambda) rad/s (:451).
    Returns (lambda (k,), Phi (ndof,k)) with M-normalised columns, zeros on fixed DOFs."""
    Kff = K[free][:, free]
    Mff = M[free][:, free]
    nf = Kff.shape[0]
    if nf <= dense_limit:
        w, v = sla.eigh(Kff.toarray(), Mff.toarray())
    else:
        w, v = spla.eigsh(Kff.tocsc(), k=min(k, nf - 2), M=Mff.tocsc(), sigma=0.0, which="LM", tol=1e-12)
        o = np.argsort(w)
        w, v = w[o], v[:, o]
    keep = w > 1e-6
    w, v = w[keep][:k], v[:, keep][:, :k]
    nrm = np.sqrt(np.einsum("ij,ij->j", v, Mff @ v))
    v = v / nrm
    phi = np.zeros((K.shape[0], v.shape[1]))
    phi[free] = v
    return w, phi


def frame_run(mesh, section_props, bc_data, E, nu, k_modes=20, solver="auto"):
    """Whole BeamSolver.run_simulation pipeline (:345-455) on sparse storage.
    section_props: group -> 8-tuple (the calculate_section_properties record)."""
    names = list(section_props.keys())
    gid2name = {int(v[0]): kk for kk, v in mesh.field_data.items()}
    tags = np.asarray(mesh.cell_data_dict["gmsh:physical"]["line"])
    elem_sec = np.array([names.index(gid2name[int(t)]) for t in tags], dtype=np.int32)
    props = np.asarray([section_props[nm] for nm in names], dtype=np.float64)
    conn = mesh.cells_dict["line"]
    K, M = frame_assemble(mesh.points, conn, elem_sec, props, E, nu)
    fixed, free, f = frame_bc(mesh, bc_data)
    u, info = solve_static(K, f, fixed, free, method=solver)
    out = {"K": K, "M": M, "fixed": fixed, "free": free, "f": f, "u": u, "solve_info": info,
           "reactions": reactions(K, u, f),
           "smoothed_stresses": frame_stress(mesh.points, conn, elem_sec, props, E, nu, u)}
    if k_modes:
        lam, phi = frame_modal(K, M, free, k_modes)
        out["eigenvalues"] = lam
        out["natural_frequencies"] = np.sqrt(lam)
        out["mode_shapes"] = phi
    return out


# ------------------------------------------------------------------------ Tet10

def tet10_material(E, v):
    """ReactionSolver.py:87-98."""
    C1 = E / ((1 + v) * (1 - 2 * v))
    C2 = (1 - 2 * v) / 2
    C = np.zeros((6, 6))
    C[:3, :3] = v
    C[0, 0] = C[1, 1] = C[2, 2] = 1 - v
    C[3, 3] = C[4, 4] = C[5, 5] = C2
    return C1 * C


def tet10_shape_derivs(xi, eta, zeta):
    """ReactionSolver.py:100-113: dN/d(xi,eta,zeta), shape (3,10), meshio node order."""
    L2, L3, L4 = xi, eta, zeta
    L1 = 1 - xi - eta - zeta
    dN_L = np.array([
        [4 * L1 - 1, 0, 0, 0], [0, 4 * L2 - 1, 0, 0], [0, 0, 4 * L3 - 1, 0], [0, 0, 0, 4 * L4 - 1],
        [4 * L2, 4 * L1, 0, 0], [0, 4 * L3, 4 * L2, 0], [4 * L3, 0, 4 * L1, 0], [4 * L4, 0, 0, 4 * L1],
        [0, 4 * L4, 0, 4 * L2], [0, 0, 4 * L4, 4 * L3]]).T
    dL = np.array([[-1, -1, -1], [1, 0, 0], [0, 1, 0], [0, 0, 1]])
    return dL.T @ dN_L


TET10_GAUSS = np.array([[GAUSS_A, GAUSS_B, GAUSS_B], [GAUSS_B, GAUSS_A, GAUSS_B],
                        [GAUSS_B, GAUSS_B, GAUSS_A], [GAUSS_B, GAUSS_B, GAUSS_B]])


def tet10_element_matrices(points, conn, E, v):
    """Vectorised ReactionSolver.py:126-146.  Returns (Ke (E,30,30), skipped_count) where
    skipped_count = Gauss points with detJ <= 1e-12 (:133-135), which contribute nothing."""
    X = np.asarray(points, dtype=np.float64)[np.asarray(conn)]        # (E,10,3)
    C = tet10_material(E, v)
    ne = len(conn)
    Ke = np.zeros((ne, 30, 30))
    skipped = 0
    for g in TET10_GAUSS:
        dN = tet10_shape_derivs(*g)                                   # (3,10)
        J = np.einsum("ik,ekj->eij", dN, X)                           # (E,3,3)
        detJ = np.linalg.det(J)
        good = detJ > DETJ_MIN
        skipped += int((~good).sum())
        Jsafe = J.copy()
        Jsafe[~good] = np.eye(3)
        dNg = np.linalg.inv(Jsafe) @ dN                               # (E,3,10)
        B = np.zeros((ne, 6, 30))
        dx, dy, dz = dNg[:, 0], dNg[:, 1], dNg[:, 2]
        B[:, 0, 0::3] = dx; B[:, 1, 1::3] = dy; B[:, 2, 2::3] = dz
        B[:, 3, 0::3] = dy; B[:, 3, 1::3] = dx
        B[:, 4, 1::3] = dz; B[:, 4, 2::3] = dy
        B[:, 5, 0::3] = dz; B[:, 5, 2::3] = dx
        contrib = (B.transpose(0, 2, 1) @ C @ B) * (detJ * GAUSS_W)[:, None, None]
        contrib[~good] = 0.0
        Ke += contrib
    return Ke, skipped


def tet10_assemble(points, conn, E, v):
    """ReactionSolver.py:115-152.  Returns (K CSR on the structural 3x3-block pattern,
    skipped Gauss point count).  The reference's lil->csr drops exact zeros: compare
    patterns after ``eliminate_zeros()`` on both sides."""
    Ke, skipped = tet10_element_matrices(points, conn, E, v)
    return _scatter_blocks(np.asarray(conn), Ke, len(points), 3, 10), skipped


def tet10_bc(points, diri_nodes, neumann_nodes, force_data, fix_data):
    """ReactionSolver.py:154-194: nearest node inside the group; DOF fixed iff flag == 0."""
    points = np.asarray(points)
    n = 3 * len(points)
    f = np.zeros(n)
    fixed, fixed_info, force_info = [], [], []
    for fi in fix_data:
        pos = np.array([fi["pos_x"], fi["pos_y"], fi["pos_z"]])
        node = diri_nodes[np.argmin(np.linalg.norm(points[diri_nodes] - pos, axis=1))]
        dofs = [3 * node + c for c, kk in enumerate(("fix_x", "fix_y", "fix_z")) if fi[kk] == 0]
        fixed.extend(dofs)
        fixed_info.append({"node_idx": node, "pos": points[node], "dofs": dofs})
    fixed = np.unique(fixed).astype(np.int64)
    for fo in force_data:
        vec = np.array([fo["force_x"], fo["force_y"], fo["force_z"]])
        pos = np.array([fo["force_x_pstn"], fo["force_y_pstn"], fo["force_z_pstn"]])
        node = neumann_nodes[np.argmin(np.linalg.norm(points[neumann_nodes] - pos, axis=1))]
        f[3 * node:3 * node + 3] += vec
        force_info.append({"node_idx": node, "pos": points[node], "force_vec": vec})
    active = np.setdiff1d(np.arange(n), fixed)
    return fixed, active, f, fixed_info, force_info


def tet10_run(mesh, force_data, fix_data, E, v, solver="auto"):
    conn = mesh.cells_dict["tetra10"]
    K, skipped = tet10_assemble(mesh.points, conn, E, v)
    diri = mesh_group_nodes(mesh, "vertex", "Diri_BCs")
    neu = mesh_group_nodes(mesh, "vertex", "Neumann_BCs")
    fixed, active, f, finfo, ginfo = tet10_bc(mesh.points, diri, neu, force_data, fix_data)
    u, info = solve_static(K, f, fixed, active, method=solver)
    return {"K": K, "negative_detJ_count": skipped, "fixed_dofs": fixed, "active_dofs": active, "f": f,
            "u": u, "reaction_forces": reactions(K, u), "fixed_nodes_info": finfo,
            "applied_forces_info": ginfo, "solve_info": info}
