"""CPU study (scipy, not a test and not part of the product path): iteration counts of the additive
two-level PCG  M^-1 = omega D^-1 + P (P^T A P)^-1 P^T  on the lattice frame of BASELINE config 3 for
different coarse spaces P.  It motivated csrc/twolevel.cu (rigid-body modes per RCB aggregate) and
records what the next coarse space should be.

    python tests/prototypes/coarse_space_study.py 24 24 22

Results, rtol 1e-12, omega 2 (iterations):
  24x24x22 lattice (76k DOF)     Jacobi 2,9xx | rigid-body modes of 33 RCB aggregates 1,167
                                 | + one axial translation mode per lattice line (1,632 modes) 220
                                 | + line segments of 8 nodes (4,896 modes) 120 | segments of 4: 114
                                 | smoothed-aggregation prolongator (1 / 2 Jacobi sweeps) 1,057 / 1,033
  56x56x54 lattice (1M DOF)      Jacobi 6,931 | rigid-body modes of 444 aggregates 1,372 (GPU: same count)
Additive multilevel variants that avoid one big coarse solve (second part of main(); 36x36x34, 280k DOF):
  whole lines only, exact solve of the 3,744 line unknowns                                   333
  whole lines, exact solve per member direction ("family": three ~1,250^2 blocks, cross terms dropped)
    + Jacobi on line segments of 8 (one extra diagonal scaling of an aggregated residual)    209
    (the full 3,744^2 solve instead of the three blocks: also 209)
  ... + rigid-body modes of 115 / 688 RCB aggregates                                         183 / 165
  Jacobi on segments of 2,4,8,16,32 (BPX along the lines) instead of 8 only                  217
  (24x24x22: 148 / 135 with rigid-body modes; rigid-body modes only: 1,167)
So: M^-1 = 2 D^-1 + sum over the three member directions of P_f (P_f^T A P_f)^-1 P_f^T
          + P_seg diag(P_seg^T A P_seg)^-1 P_seg^T + P_rbm (P_rbm^T A P_rbm)^-1 P_rbm^T
needs three more dense inverses of the size the product already builds (about 3,000^2 each at 1M DOF),
no extra operator application, and cuts the iteration count another ~7x.

Reading: after diagonal scaling the slow modes of a frame are not only the locally rigid ones.  A row of
collinear members moving along its own axis costs bending energy of the crossing members only
(12 EI / L^3) while its diagonal carries the axial stiffness EA / L, so every "line translation" is a
near-null vector of D^-1 A — the classic anisotropy that strength-of-connection (line) aggregates
resolve.  With line segments of length s the count follows the 1-D Laplacian along the segment,
about 14 (2 s / pi), independent of the mesh size.
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from fem_calculator_b200 import api, meshgen  # noqa: E402
from oracle import ref_sparse as S  # noqa: E402


def pcg(A, b, Minv, rtol=1e-12, maxit=50000):
    x = np.zeros_like(b); r = b.copy(); z = Minv(r); p = z.copy(); rz = r @ z; nb = np.linalg.norm(b)
    for it in range(1, maxit + 1):
        q = A @ p; a = rz / (p @ q); x += a * p; r -= a * q
        if np.linalg.norm(r) <= rtol * nb:
            return x, it
        z = Minv(r); rz2 = r @ z; p = z + (rz2 / rz) * p; rz = rz2
    return x, maxit


def rigid_body_modes(points, agg, nagg, mask):
    N = len(points)
    cen = np.stack([np.bincount(agg, weights=points[:, d], minlength=nagg) for d in range(3)], 1) / \
        np.maximum(np.bincount(agg, minlength=nagg), 1)[:, None]
    rho = points - cen[agg]
    n = np.arange(N)
    rows, cols, vals = [], [], []
    for i in range(3):
        rows.append(6 * n + i); cols.append(6 * agg + i); vals.append(np.ones(N))
    for j in range(3):
        e = np.zeros(3); e[j] = 1
        u = np.cross(np.broadcast_to(e, rho.shape), rho)
        for i in range(3):
            rows.append(6 * n + i); cols.append(6 * agg + 3 + j); vals.append(u[:, i])
        rows.append(6 * n + 3 + j); cols.append(6 * agg + 3 + j); vals.append(np.ones(N))
    P = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(6 * N, 6 * nagg))
    return (sp.diags(mask.astype(float)) @ P).tocsr()


def line_modes(lat, mask, seg):
    nx, ny, nz = lat
    N = nx * ny * nz
    ix, iy, iz = np.unravel_index(np.arange(N), lat)
    rows, cols, c0 = [], [], 0
    for comp, (a, b, c, na) in enumerate([(iy, iz, ix, nx), (ix, iz, iy, ny), (ix, iy, iz, nz)]):
        nseg = -(-na // seg)
        _, inv = np.unique((a * (b.max() + 1) + b) * nseg + c // seg, return_inverse=True)
        rows.append(6 * np.arange(N) + comp); cols.append(c0 + inv); c0 += inv.max() + 1
    Pl = sp.csr_matrix((np.ones(3 * N), (np.concatenate(rows), np.concatenate(cols))), shape=(6 * N, c0))
    return (sp.diags(mask.astype(float)) @ Pl).tocsr()


def main():
    lat = tuple(int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (16, 16, 14)
    mesh, sec, bc = meshgen.lattice_frame_case(*lat, jitter=0.05)
    es, props = meshgen.section_table(mesh, sec)
    K, _ = S.frame_assemble(mesh.points, mesh.cells_dict["line"], es, props, meshgen.E_STEEL, meshgen.NU_STEEL)
    fixed, free, f = S.frame_bc(mesh, bc)
    ndof = K.shape[0]
    mask = np.zeros(ndof, bool); mask[free] = True
    Dm = sp.diags(mask.astype(float))
    A = (Dm @ K @ Dm + sp.diags((~mask).astype(float))).tocsr()
    b = f * mask
    d = A.diagonal()
    print("lattice", lat, "DOF", ndof)
    print("jacobi", pcg(A, b, lambda r: r / d)[1], flush=True)

    def run(P, label, om=2.0):
        Kc = (P.T @ A @ P).toarray()
        dz = np.diag(Kc) <= 1e-300
        Kc[dz, dz] = 1.0
        Kci = np.linalg.inv(Kc + 1e-10 * np.diag(np.diag(Kc)))
        print(label, "coarse dim", P.shape[1], "iterations", pcg(A, b, lambda r: om * r / d + P @ (Kci @ (P.T @ r)))[1], flush=True)

    nagg = max(1, len(mesh.points) // 381)
    agg = api.symbolic_aggregates(mesh.points, nagg)       # the product's own RCB (csrc/coarse.cpp, host-only)
    P = rigid_body_modes(mesh.points, agg.astype(np.int64), nagg, mask)
    run(P, "rigid-body modes per aggregate")
    for seg in (10 ** 6, 8, 4):
        run(sp.hstack([P, line_modes(lat, mask, seg)]).tocsr(), f"rigid-body + line segments of {min(seg, max(lat))}")

    # additive multilevel: whole lines solved per member direction, Jacobi on line segments, rigid-body modes
    def inv(M):
        M = M.toarray() if sp.issparse(M) else M
        dz = np.diag(M) <= 1e-300
        M[dz, dz] = 1.0
        return np.linalg.inv(M + 1e-10 * np.diag(np.diag(M)))

    nx, ny, nz = lat
    Pl = line_modes(lat, mask, 10 ** 6)
    off = np.cumsum([0, ny * nz, nx * nz, nx * ny])
    Kl = (Pl.T @ A @ Pl).toarray()
    Kbd = np.zeros_like(Kl)
    for k in range(3):
        Kbd[off[k]:off[k + 1], off[k]:off[k + 1]] = Kl[off[k]:off[k + 1], off[k]:off[k + 1]]
    Kbd_i = inv(Kbd)
    Ps = line_modes(lat, mask, 8)
    ds = (Ps.T @ A @ Ps).diagonal()
    ds[ds <= 0] = 1.0
    Kr_i = inv(P.T @ A @ P)
    print("whole lines (per-direction exact solve) only, iterations",
          pcg(A, b, lambda r: 2 * r / d + Pl @ (Kbd_i @ (Pl.T @ r)))[1], flush=True)
    print("  + Jacobi on segments of 8, iterations",
          pcg(A, b, lambda r: 2 * r / d + Pl @ (Kbd_i @ (Pl.T @ r)) + Ps @ ((Ps.T @ r) / ds))[1], flush=True)
    print("  + rigid-body modes per aggregate, iterations",
          pcg(A, b, lambda r: 2 * r / d + Pl @ (Kbd_i @ (Pl.T @ r)) + Ps @ ((Ps.T @ r) / ds) + P @ (Kr_i @ (P.T @ r)))[1], flush=True)


if __name__ == "__main__":
    main()
