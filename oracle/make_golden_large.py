"""TEST INFRASTRUCTURE — generates the sampled golden vectors of the large lattice frames
(tests/golden/lattice_*_sampled.npz) with the CPU oracle: oracle/ref_sparse.py assembly (the restatement
pinned against the unmodified reference in tests/test_oracle.py) + the multi-threaded Jacobi-CG of
oracle/cg_omp.c at rtol 1e-13 followed by two steps of iterative refinement.  A full displacement vector of BASELINE configs[2] is 8 MB, so the fixture keeps
u at 4,096 seeded DOFs, ||u||_2, the support-reaction resultant and the norm of K u - f on the fixed DOFs.

    python -m oracle.make_golden_large 40 40 38
    python -m oracle.make_golden_large 56 56 54        # BASELINE configs[2], ~10 min on 8 cores
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fem_calculator_b200 import meshgen  # noqa: E402  (mesh generator + closed-form sections: inputs, not numerics)
from oracle import native, ref_sparse as S  # noqa: E402

JITTER = 0.05
N_SAMPLE = 4096


def main():
    lat = tuple(int(v) for v in sys.argv[1:4])
    mesh, sec, bc = meshgen.lattice_frame_case(*lat, jitter=JITTER)
    es, props = meshgen.section_table(mesh, sec)
    t0 = time.time()
    K, _ = S.frame_assemble(mesh.points, mesh.cells_dict["line"], es, props, meshgen.E_STEEL, meshgen.NU_STEEL)
    fixed, free, f = S.frame_bc(mesh, bc)
    Kff = K[free][:, free].tocsr()
    t1 = time.time()
    uf, info = native.pcg_jacobi(Kff, f[free], rtol=1e-13)
    assert info["flag"] == 0, info
    # CG's recurrence residual drifts from the true one (4e-11 after 5,135 iterations at 40x40x38): two steps of
    # iterative refinement on the true residual f - K u bring the golden down to the rounding floor of K u itself
    refine = []
    for _ in range(2):
        res = f[free] - Kff @ uf
        refine.append(float(np.linalg.norm(res) / np.linalg.norm(f[free])))
        e, i2 = native.pcg_jacobi(Kff, res, rtol=1e-6)
        assert i2["flag"] == 0, i2
        uf = uf + e
        info["iterations"] += i2["iterations"]
    t2 = time.time()
    u = np.zeros(K.shape[0])
    u[free] = uf
    r = K @ u - f
    true_res = np.linalg.norm((K @ u - f)[free]) / np.linalg.norm(f[free])
    rng = np.random.default_rng(meshgen.SEED)
    idx = np.sort(rng.choice(free, size=min(N_SAMPLE, len(free)), replace=False))
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                       f"lattice_{lat[0]}x{lat[1]}x{lat[2]}_sampled.npz")
    np.savez_compressed(out, lattice=np.asarray(lat), jitter=JITTER, idx=idx, u_idx=u[idx], u_norm=np.linalg.norm(u),
                        u_absmax=np.abs(u).max(), reaction_sum=r[fixed].reshape(-1, 6)[:, :3].sum(axis=0),
                        reaction_norm=np.linalg.norm(r[fixed]), f_norm=np.linalg.norm(f), n_dof=K.shape[0],
                        n_fixed=len(fixed), iterations=info["iterations"], rel_residual=info["rel_residual"],
                        true_rel_residual=true_res)
    print(f"{lat}: {K.shape[0]} DOF, assembly {t1 - t0:.1f} s, Jacobi-CG {info['iterations']} its in {t2 - t1:.1f} s on "
          f"{native.threads()} threads, true residual before / between / after refinement {refine[0]:.2e} / {refine[1]:.2e} / "
          f"{true_res:.2e} -> {out}")


if __name__ == "__main__":
    main()
