"""Host logic of the row-block multi-GPU path, on CPU:

* partition invariants (owned ranges tile the nodes; send/recv lists of neighbouring ranks
  mirror each other; owned rows of the locally assembled K equal the global rows);
* a world_size-2 `gloo` run of the distributed Chronopoulos-Gear PCG — the same sequence of
  halo exchange / one all-reduce of three scalars / fused update that csrc/dist.cu enqueues —
  with the local operators built by the oracle, compared with the global solve.
"""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp

from fem_calculator_b200 import compat, meshgen, partition as P
from fem_calculator_b200.sections import calculate_section_properties as csp
from oracle import ref_sparse as S


def _case(nx=6, ny=4, nz=5):
    mesh, sec, bc = meshgen.lattice_frame_case(nx, ny, nz, jitter=0.05)
    es, props, _ = compat.frame_section_table(mesh, sec, csp)
    fixed, f = compat.frame_bc_vectors(mesh, bc, len(mesh.points))
    return mesh, es, props, fixed, f


@pytest.mark.parametrize("kind", ["slabs", "boxes"])
@pytest.mark.parametrize("world", [1, 2, 3, 5, 8])
def test_partition_invariants(world, kind):
    mesh, es, props, fixed, f = _case()
    conn = mesh.cells_dict["line"]
    n = len(mesh.points)
    owner = P.box_owner(mesh.points, world) if kind == "boxes" else None
    if owner is not None:
        cnt = np.bincount(owner, minlength=world)
        assert cnt.max() - cnt.min() <= 1                             # equal node counts
    parts = [P.partition_mesh(conn, n, world, r, owner=owner) for r in range(world)]
    owned = np.concatenate([p.local_nodes[:p.n_owned] for p in parts])
    assert np.array_equal(np.sort(owned), np.arange(n))
    for p in parts:
        assert np.all(np.diff(p.owned_nodes) > 0)                     # owned rows in ascending global order
    covered = np.zeros(len(conn), dtype=int)
    for p in parts:
        assert np.all(np.diff(p.elem_ids) > 0)                       # ascending global element order
        assert np.array_equal(p.local_nodes[p.conn_local], conn[p.elem_ids])
        ghosts = p.local_nodes[p.n_owned:]
        for k in range(len(p.nbr)):                                   # grouped by owner, ascending inside a group
            run = ghosts[p.recv_start[k] - p.n_owned:p.recv_start[k] - p.n_owned + p.recv_count[k]]
            assert np.all(np.diff(run) > 0)
        covered[p.elem_ids] += 1
        # what I receive from s is exactly what s sends me, in the same order
        for k, s in enumerate(p.nbr):
            q = parts[s]
            kk = list(q.nbr).index(p.rank)
            sent = q.local_nodes[q.send_nodes[q.send_ptr[kk]:q.send_ptr[kk + 1]]]
            recv = p.local_nodes[p.recv_start[k]:p.recv_start[k] + p.recv_count[k]]
            assert np.array_equal(sent, recv)
        assert p.recv_count.sum() == len(ghosts)
    assert covered.min() >= 1                                          # every element lives somewhere


def test_owned_rows_of_local_assembly_equal_global_rows():
    mesh, es, props, fixed, f = _case()
    conn = mesh.cells_dict["line"]
    n = len(mesh.points)
    E, nu = meshgen.E_STEEL, meshgen.NU_STEEL
    K, _ = S.frame_assemble(mesh.points, conn, es, props, E, nu)
    owner = P.box_owner(mesh.points, 3)
    for r in range(6):
        p = P.partition_mesh(conn, n, 3, r % 3, owner=owner if r >= 3 else None)
        Kl, _ = S.frame_assemble(mesh.points[p.local_nodes], p.conn_local, es[p.elem_ids], props, E, nu)
        ld = p.local_dofs(6)
        rows = slice(0, 6 * p.n_owned)
        Kg = K[ld[rows]][:, ld]                                       # global rows/cols in local numbering
        d = (Kl[rows] - Kg)
        # same contributions; scipy's COO->CSR duplicate summation order is unspecified, so the oracle
        # agrees to rounding (the device assembly is checked bit-exact in test_gpu_dist.py)
        assert abs(d).max() <= 1e-15 * abs(Kg).max()
        fl, f_l = P.localize_bc(p, 6, fixed, f, n)
        assert np.array_equal(np.sort(ld[fl]), np.intersect1d(fixed, ld))
        assert np.array_equal(f_l, f[ld])


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _dist_pcg_worker(rank, world, port, out_dir, kind="slabs"):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mesh, es, props, fixed, f = _case()
    conn = mesh.cells_dict["line"]
    n = len(mesh.points)
    E, nu = meshgen.E_STEEL, meshgen.NU_STEEL
    p = P.partition_mesh(conn, n, world, rank, owner=P.box_owner(mesh.points, world) if kind == "boxes" else None)
    Kl, _ = S.frame_assemble(mesh.points[p.local_nodes], p.conn_local, es[p.elem_ids], props, E, nu)
    fixed_l, f_l = P.localize_bc(p, 6, fixed, f, n)
    nloc, no = Kl.shape[0], 6 * p.n_owned
    free = np.ones(nloc, dtype=bool); free[fixed_l] = False
    Pm = sp.diags(free.astype(float))
    A = (Pm @ Kl @ Pm + sp.diags((~free).astype(float))).tocsr()[:no]   # masked operator, owned rows
    dinv = 1.0 / A.diagonal()
    b = np.where(free, f_l, 0.0)[:no]

    def halo(v):                                                        # same pattern as dist_halo_exchange
        reqs, bufs = [], []
        for k, s in enumerate(p.nbr):
            sn = p.send_nodes[p.send_ptr[k]:p.send_ptr[k + 1]]
            sb = torch.from_numpy(np.ascontiguousarray(v.reshape(-1, 6)[sn].reshape(-1)))
            rb = torch.zeros(int(p.recv_count[k]) * 6, dtype=torch.float64)
            reqs += [dist.isend(sb, int(s)), dist.irecv(rb, int(s))]
            bufs.append((k, rb))
        for rq in reqs:
            rq.wait()
        for k, rb in bufs:
            v[6 * p.recv_start[k]:6 * (p.recv_start[k] + p.recv_count[k])] = rb.numpy()

    x = np.zeros(nloc); r = np.zeros(nloc); z = np.zeros(nloc); pp = np.zeros(no); q = np.zeros(no)
    r[:no] = b; z[:no] = dinv * b
    red = np.array([0.0, r[:no] @ z[:no], b @ b])
    gprev = alpha = 1.0
    tol2 = None
    its = 0
    for it in range(5000):
        halo(z)
        s = A @ z
        red[0] = z[:no] @ s
        t = torch.from_numpy(red.copy()); dist.all_reduce(t); red = t.numpy().copy()
        delta, gamma, rr = red
        if it == 0:
            tol2 = (1e-12) ** 2 * rr
        elif rr <= tol2:
            break
        beta = 0.0 if it == 0 else gamma / gprev
        den = delta if it == 0 else delta - beta * gamma / alpha
        alpha_new = gamma / den
        pp = z[:no] + beta * pp
        q = s + beta * q
        x[:no] += alpha_new * pp
        r[:no] -= alpha_new * q
        z[:no] = dinv * r[:no]
        gprev, alpha = gamma, alpha_new
        red[1] = r[:no] @ z[:no]; red[2] = r[:no] @ r[:no]
        its += 1
    np.save(os.path.join(out_dir, f"u_{rank}.npy"), x[:no])
    np.save(os.path.join(out_dir, f"nodes_{rank}.npy"), p.owned_nodes)
    np.save(os.path.join(out_dir, f"its_{rank}.npy"), np.array([its]))
    dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["slabs", "boxes"])
def test_distributed_pcg_gloo_world2(tmp_path, kind):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_dist_pcg_worker, args=(world, _free_port(), str(tmp_path), kind), nprocs=world, join=True)
    mesh, es, props, fixed, f = _case()
    E, nu = meshgen.E_STEEL, meshgen.NU_STEEL
    K, _ = S.frame_assemble(mesh.points, mesh.cells_dict["line"], es, props, E, nu)
    free = np.setdiff1d(np.arange(len(f)), fixed)
    uo, _ = S.solve_static(K, f, fixed, free, method="direct")
    u = np.zeros(len(f))
    for r in range(world):                                             # owned rows of rank r, ascending global node order
        u.reshape(-1, 6)[np.load(tmp_path / f"nodes_{r}.npy")] = np.load(tmp_path / f"u_{r}.npy").reshape(-1, 6)
    assert np.linalg.norm(u - uo) <= 1e-10 * np.linalg.norm(uo)
    assert int(np.load(tmp_path / "its_0.npy")[0]) == int(np.load(tmp_path / "its_1.npy")[0]) > 10
