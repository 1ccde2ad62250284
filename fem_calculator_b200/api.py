"""Thin Python objects over the C-ABI handle.  Array in, array out; all numerics run in
libfemb200.so on the GPU.  The reference-shaped adapters live in ``compat.py``."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


class Handle:
    """Owns one femb_handle (one device, one stream)."""

    def __init__(self, device: int = 0):
        self.lib = L.load()
        self._h = C.c_void_p()
        rc = self.lib.femb_create(int(device), C.byref(self._h))
        if rc != 0:
            raise L.FembError(rc, "femb_create failed: no usable CUDA (sm_100) device — femb200 has no CPU fallback")
        self.device = device

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.femb_destroy(self._h)        # also drops every femb_host_register registration of the handle
            self._h = None
            self.__dict__.pop("_pins", None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != 0:
            msg = self.lib.femb_last_error(self._h)
            raise L.FembError(rc, msg.decode() if msg else "")

    # ---- shared ---------------------------------------------------------------------
    def assemble(self):
        self._check(self.lib.femb_assemble(self._h))

    def get_csr(self, which=L.MAT_K):
        """(indptr int32, indices int32, data float64) of K or M."""
        nr, nz = C.c_int64(), C.c_int64()
        self._check(self.lib.femb_get_csr_size(self._h, which, C.byref(nr), C.byref(nz)))
        indptr = np.zeros(nr.value + 1, dtype=np.int32)
        indices = np.zeros(nz.value, dtype=np.int32)
        data = np.zeros(nz.value, dtype=np.float64)
        self._check(self.lib.femb_get_csr(self._h, which, indptr, indices, data))
        return indptr, indices, data

    def set_bc(self, fixed_dofs, f, u_prescribed=None):
        fixed = np.ascontiguousarray(fixed_dofs, dtype=np.int64)
        f = np.ascontiguousarray(f, dtype=np.float64)
        up = None if u_prescribed is None else np.ascontiguousarray(u_prescribed, dtype=np.float64)
        self._check(self.lib.femb_set_bc(self._h, len(fixed), L.ptr(fixed), f, L.ptr(up)))
        self.ndof = len(f)

    def solve_static(self, method=L.SOLVER_AUTO, precond=L.PRECOND_AUTO, rtol=1e-12, max_iter=200000,
                     check_every=50, minus_f=True, want_u=True, want_reactions=True, profile=False, op=L.OP_AUTO):
        o = L.SolveOpts(method, precond, max_iter, check_every, rtol, int(profile), int(op))
        st = L.Stats()
        u = np.zeros(self.ndof) if want_u else None
        r = np.zeros(self.ndof) if want_reactions else None
        rc = self.lib.femb_solve_static(self._h, C.byref(o), int(minus_f), L.ptr(u), L.ptr(r), C.byref(st))
        self.last_stats = st.as_dict()
        self._check(rc)
        return u, r, self.last_stats

    def apply_k(self, x, op=L.OP_AUTO, masked=False):
        """(K x or K_ff x, operator used) — the product the Krylov loops are built on."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros_like(x)
        used = C.c_int32()
        self._check(self.lib.femb_apply_k(self._h, int(op), int(masked), x, y, C.byref(used)))
        return y, used.value

    def modal(self, k=20, rtol=1e-8, max_iter=5000, block=0, lambda_min=1e-6, op=L.OP_AUTO, precond=L.PRECOND_AUTO,
              accept_rtol=0.0):
        """``accept_rtol`` > rtol: also return a solve that stagnated above rtol when its worst pencil residual is below
        accept_rtol (stats['converged'] == 0, stats['rel_residual'] = what was reached); 0 = strict."""
        o = L.EigOpts(k, block, max_iter, int(op), rtol, lambda_min, int(precond), 0, float(accept_rtol))
        st = L.Stats()
        lam = np.zeros(k)
        phi = np.zeros((k, self.ndof))  # column-major (ndof,k) == row-major (k,ndof)
        nf = C.c_int32()
        rc = self.lib.femb_modal(self._h, C.byref(o), L.ptr(lam), L.ptr(phi), C.byref(nf), C.byref(st))
        self.last_stats = st.as_dict()
        self._check(rc)
        n = nf.value
        return lam[:n].copy(), np.ascontiguousarray(phi[:n].T), self.last_stats

    def timer_start(self):
        self._check(self.lib.femb_timer(self._h, 0, None))

    def timer_stop(self):
        ms = C.c_double()
        self._check(self.lib.femb_timer(self._h, 1, C.byref(ms)))
        return ms.value

    def time_kernel(self, which, warm=3, reps=20):
        ms, by = C.c_double(), C.c_double()
        self._check(self.lib.femb_time_kernel(self._h, which, warm, reps, C.byref(ms), C.byref(by)))
        return ms.value, by.value


class FrameModel(Handle):
    """3-D frame: element generation -> assembly -> BC -> solve -> stress / modal."""

    def set_mesh(self, points, conn, elem_sec, sec_props, E, G, rho=7850.0):
        self.points = np.ascontiguousarray(points, dtype=np.float64)
        self.conn = np.ascontiguousarray(conn, dtype=np.int64)
        es = np.ascontiguousarray(elem_sec, dtype=np.int32)
        sp = np.ascontiguousarray(sec_props, dtype=np.float64).reshape(-1, 8)
        self.n_nodes, self.n_elem = len(self.points), len(self.conn)
        self.ndof = 6 * self.n_nodes
        self._check(self.lib.femb_frame_set_mesh(self._h, self.n_nodes, self.n_elem, self.points, self.conn.reshape(-1),
                                                 es, len(sp), sp.reshape(-1), float(E), float(G), float(rho)))

    def elements(self, want_k=True, want_m=True):
        ke = np.zeros((self.n_elem, 12, 12)) if want_k else None
        me = np.zeros((self.n_elem, 12, 12)) if want_m else None
        self._check(self.lib.femb_frame_elements(self._h, L.ptr(ke), L.ptr(me)))
        return ke, me

    def stress(self, u=None):
        s = np.zeros(self.n_nodes)
        uu = None if u is None else np.ascontiguousarray(u, dtype=np.float64)
        self._check(self.lib.femb_frame_stress(self._h, L.ptr(uu), L.ptr(s)))
        return s

    _PIN_MIN_BYTES = 8 << 20

    def _pin(self, role, arr):
        """Page-lock ``arr`` in place for this handle and keep a reference to it: while the model holds the array its
        address cannot be freed and handed out again, so the registration can never alias other memory.  One array per
        role ('f', 'u'); a different array replaces (and unregisters) the previous one."""
        pins = self.__dict__.setdefault("_pins", {})
        cur = pins.get(role)
        if cur is not None and cur is arr:
            return
        if cur is not None:
            self.lib.femb_host_unregister(self._h, C.c_void_p(cur.ctypes.data))
            pins.pop(role)
        if arr is not None and arr.nbytes >= self._PIN_MIN_BYTES:
            if self.lib.femb_host_register(self._h, C.c_void_p(arr.ctypes.data), arr.nbytes) == 0:
                pins[role] = arr

    def batch_solve(self, xyz, sec_props, E, G, fixed_mask, f, want_u=True, out=None, pin=True):
        """BASELINE config 4: (u (n_models, ndof), stats).  ``out``: a C-contiguous float64 (n_models, ndof) array to
        receive u (pass the same one again to reuse its page-locked registration); ``pin``: page-lock f and u in place
        (buffers of 8 MB or more) — the model keeps them referenced until other arrays come or it is closed."""
        xyz = np.ascontiguousarray(xyz, dtype=np.float64)
        sp = np.ascontiguousarray(sec_props, dtype=np.float64).reshape(-1, 8)
        fm = np.ascontiguousarray(fixed_mask, dtype=np.uint8)
        f = np.ascontiguousarray(f, dtype=np.float64)
        nm, ne = len(sp), len(xyz) - 1
        u = None
        if want_u:
            u = out if out is not None else np.zeros((nm, 6 * len(xyz)))
            if u.shape != (nm, 6 * len(xyz)) or u.dtype != np.float64 or not u.flags.c_contiguous:
                raise ValueError("out must be a C-contiguous float64 array of shape (n_models, ndof)")
        if pin:
            self._pin("f", f)
            self._pin("u", u)
        st = L.Stats()
        rc = self.lib.femb_frame_batch_solve(self._h, nm, ne, xyz.reshape(-1), sp.reshape(-1), float(E), float(G),
                                             fm, f.reshape(-1), L.ptr(u), C.byref(st))
        self.last_stats = st.as_dict()
        self._check(rc)
        return u, self.last_stats


class DistFrameModel(FrameModel):
    """One rank of a row-block partitioned frame (one process per GPU; csrc/dist.cu).

    Every rank passes the same GLOBAL arrays (what the reference holds in memory) plus its rank;
    the local mesh and halo lists come from ``partition.partition_mesh``.  ``unique_id`` is the
    128-byte NCCL id made by rank 0 (``DistFrameModel.unique_id()``) and broadcast by the caller
    (torch.distributed / MPI / a file); world == 1 needs none."""

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        rc = L.load().femb_dist_unique_id(buf)
        if rc != 0:
            raise L.FembError(rc, "femb_dist_unique_id failed (libnccl.so.2 not loadable?)")
        return bytes(buf)

    def setup(self, points, conn, elem_sec, sec_props, E, G, fixed_dofs, f, rank, world, unique_id=None, rho=7850.0,
              all_gather=None, lines=True, line_bundles=768, partition="boxes"):
        """``all_gather(obj) -> [obj of rank 0, ..., obj of rank world-1]`` (e.g. a wrapper of
        torch.distributed.all_gather_object) switches the iteration's exchanges from NCCL to the
        peer-memory kernels (CUDA IPC over NVLink); without it NCCL is used.  ``lines``: hand the rank's
        member-line tables to the library (PRECOND_LINES / AUTO on the partition; needs the peer-memory path).
        ``partition``: "boxes" (coordinate-bisection boxes of equal node count, ``partition.box_owner`` — member lines are
        cut into few long pieces), "slabs" (contiguous ranges of the node order) or an explicit node -> rank array."""
        from . import partition as P
        points = np.asarray(points, dtype=np.float64)
        conn = np.asarray(conn, dtype=np.int64)
        n_nodes = len(points)
        if isinstance(partition, str):
            if partition not in ("boxes", "slabs"):
                raise ValueError("partition must be 'boxes', 'slabs' or a node -> rank array")
            owner = P.box_owner(points, world) if (partition == "boxes" and world > 1) else None
        else:
            owner = np.asarray(partition, dtype=np.int64)
        self.part = part = P.partition_mesh(conn, n_nodes, world, rank, owner=owner)
        if world > 1:
            idbuf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
            self._check(self.lib.femb_dist_init(self._h, int(rank), int(world), idbuf))
        self.set_mesh(points[part.local_nodes], part.conn_local, np.asarray(elem_sec, dtype=np.int32)[part.elem_ids],
                      sec_props, E, G, rho)
        self.assemble()
        fixed_l, f_l = P.localize_bc(part, 6, fixed_dofs, f, n_nodes)
        self.set_bc(fixed_l, f_l)
        self.n_owned_dof = 6 * part.n_owned
        nbr = np.ascontiguousarray(part.nbr, dtype=np.int32)
        sp = np.ascontiguousarray(part.send_ptr, dtype=np.int64)
        sn = np.ascontiguousarray(part.send_nodes, dtype=np.int32)
        rs = np.ascontiguousarray(part.recv_start, dtype=np.int64)
        rcnt = np.ascontiguousarray(part.recv_count, dtype=np.int64)
        self._check(self.lib.femb_dist_set_halo(self._h, part.n_owned, len(nbr), L.ptr(nbr), L.ptr(sp), L.ptr(sn),
                                                L.ptr(rs), L.ptr(rcnt)))
        self._bc_local = (fixed_l, f_l)
        self.lines = False
        if lines and world > 1 and all_gather is not None and world <= 8:
            # symbolic phase of the line preconditioner on the GLOBAL mesh (host; identical on every rank), then the
            # rows of this rank's local nodes
            lb = symbolic_line_bundles(points, conn, line_bundles)
            if lb["n_lines"] > 0 and lb["coverage"] >= 0.5:
                ln = part.local_nodes
                fam_off = np.ascontiguousarray(lb["fam_off"], dtype=np.int32)
                nbl = np.ascontiguousarray(lb["node_bundle"][:, ln], dtype=np.int32)
                nll = np.ascontiguousarray(lb["node_line"][:, ln], dtype=np.int32)
                npl = np.ascontiguousarray(lb["node_pos"][:, ln], dtype=np.int32)
                ndl = np.ascontiguousarray(lb["node_dir"][:, ln, :], dtype=np.float64)
                self._check(self.lib.femb_dist_set_lines(self._h, int(fam_off[3]), L.ptr(fam_off), L.ptr(nbl), L.ptr(nll),
                                                         L.ptr(npl), L.ptr(ndl)))
                self.lines = True
        self.p2p = False
        if world > 1 and all_gather is not None and world <= 8:
            buf = (C.c_uint8 * 128)()
            self._check(self.lib.femb_dist_p2p_export(self._h, buf))
            mine = (bytes(buf), {int(s): int(r) for s, r in zip(part.nbr, part.recv_start)})
            everyone = all_gather(mine)
            blob = b"".join(e[0] for e in everyone)
            # where MY nodes start in neighbour k's ghost numbering = its recv_start for source rank `rank`
            pgs = np.array([everyone[int(s)][1][int(rank)] for s in part.nbr], dtype=np.int64)
            hb = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
            self._check(self.lib.femb_dist_p2p_import(self._h, hb, L.ptr(pgs)))
            self.p2p = True
        return part

    def set_bc_local(self):
        """Re-upload the partition's fixed-DOF list and load vector (the host -> device part of a step)."""
        self.set_bc(*self._bc_local)

    def modal_dist(self, k=20, rtol=1e-8, max_iter=5000, block=0, lambda_min=1e-6, precond=L.PRECOND_AUTO):
        """(lambda (k,), phi_owned (n_owned_dof, k), stats) — collective: every rank calls it.  ``precond``: the
        preconditioner of the inner distributed PCG solves (AUTO / LINES: the line preconditioner on the partition
        when ``setup`` handed the line tables over and the peer-memory exchange is on; else Jacobi)."""
        o = L.EigOpts(k, block, max_iter, 0, rtol, lambda_min, int(precond), 0, 0.0)
        st = L.Stats()
        lam = np.zeros(k)
        phi = np.zeros((k, self.n_owned_dof))
        nf = C.c_int32()
        rc = self.lib.femb_modal(self._h, C.byref(o), L.ptr(lam), L.ptr(phi), C.byref(nf), C.byref(st))
        self.last_stats = st.as_dict()
        self._check(rc)
        n = nf.value
        return lam[:n].copy(), np.ascontiguousarray(phi[:n].T), self.last_stats

    def solve_static_dist(self, precond=L.PRECOND_JACOBI, rtol=1e-12, max_iter=200000, check_every=50,
                          minus_f=True, want_u=True, want_reactions=True):
        """(u_owned, reactions_owned, stats) — owned DOFs only, global order within the slab."""
        o = L.SolveOpts(L.SOLVER_PCG, precond, max_iter, check_every, rtol, 0, 0)
        st = L.Stats()
        u = np.zeros(self.n_owned_dof) if want_u else None
        r = np.zeros(self.n_owned_dof) if want_reactions else None
        rc = self.lib.femb_dist_solve_static(self._h, C.byref(o), int(minus_f), L.ptr(u), L.ptr(r), C.byref(st))
        self.last_stats = st.as_dict()
        self._check(rc)
        return u, r, self.last_stats


class Tet10Model(Handle):
    """Tet10 solid: element generation -> assembly -> BC -> solve -> reactions."""

    def set_mesh(self, points, conn10, E, nu):
        self.points = np.ascontiguousarray(points, dtype=np.float64)
        self.conn = np.ascontiguousarray(conn10, dtype=np.int64)
        self.n_nodes, self.n_elem = len(self.points), len(self.conn)
        self.ndof = 3 * self.n_nodes
        self._check(self.lib.femb_tet10_set_mesh(self._h, self.n_nodes, self.n_elem, self.points,
                                                 self.conn.reshape(-1), float(E), float(nu)))

    def elements(self):
        ke = np.zeros((self.n_elem, 30, 30))
        self._check(self.lib.femb_tet10_elements(self._h, L.ptr(ke)))
        return ke

    @property
    def negative_detj(self):
        return int(self.lib.femb_tet10_negative_detj(self._h))


def symbolic_pattern(n_nodes, conn):
    """Host-only block-CSR pattern (rowptr, colidx) — needs no GPU."""
    lib = L.load()
    conn = np.ascontiguousarray(conn, dtype=np.int64)
    nper = conn.shape[1]
    nb = C.c_int64()
    rc = lib.femb_symbolic_pattern(n_nodes, len(conn), nper, conn.reshape(-1), C.byref(nb), None, None)
    if rc:
        raise L.FembError(rc, "femb_symbolic_pattern")
    rowptr = np.zeros(n_nodes + 1, dtype=np.int32)
    colidx = np.zeros(nb.value, dtype=np.int32)
    rc = lib.femb_symbolic_pattern(n_nodes, len(conn), nper, conn.reshape(-1), C.byref(nb), L.ptr(rowptr), L.ptr(colidx))
    if rc:
        raise L.FembError(rc, "femb_symbolic_pattern")
    return rowptr, colidx


def symbolic_aggregates(points, n_parts):
    """Host-only: the node aggregates of the two-level preconditioner (proportional recursive
    coordinate bisection, csrc/coarse.cpp) — aggregate id per node, needs no GPU."""
    lib = L.load()
    pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
    agg = np.zeros(len(pts), dtype=np.int32)
    rc = lib.femb_symbolic_aggregates(len(pts), L.ptr(pts), int(n_parts), L.ptr(agg))
    if rc:
        raise L.FembError(rc, "femb_symbolic_aggregates")
    return agg


def symbolic_coarse(points, conn, n_parts):
    """Host-only: (agg_of_node, nbr_ptr, nbr, blk_slot) of the two-level preconditioner's symbolic
    phase for a frame mesh (csrc/coarse.cpp) — needs no GPU."""
    lib = L.load()
    pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
    cn = np.ascontiguousarray(conn, dtype=np.int64).reshape(-1, 2)
    n = len(pts)
    agg = np.zeros(n, dtype=np.int32)
    nbr_ptr = np.zeros(int(n_parts) + 1, dtype=np.int32)
    nn = C.c_int64()
    rc = lib.femb_symbolic_coarse(n, len(cn), L.ptr(cn), L.ptr(pts), int(n_parts), L.ptr(agg), L.ptr(nbr_ptr), C.byref(nn), None, None)
    if rc:
        raise L.FembError(rc, "femb_symbolic_coarse")
    rowptr, colidx = symbolic_pattern(n, cn)
    nbr = np.zeros(nn.value, dtype=np.int32)
    slot = np.zeros(len(colidx), dtype=np.int32)
    rc = lib.femb_symbolic_coarse(n, len(cn), L.ptr(cn), L.ptr(pts), int(n_parts), L.ptr(agg), L.ptr(nbr_ptr), C.byref(nn),
                                  L.ptr(nbr), L.ptr(slot))
    if rc:
        raise L.FembError(rc, "femb_symbolic_coarse")
    return agg, nbr_ptr, nbr, slot


def symbolic_lines(points, conn, cos_tol=0.94, min_nodes=3):
    """Host-only: member lines (maximal chains of nearly collinear members, csrc/coarse.cpp) as
    (line_ptr, line_nodes, line_dir (n_lines,3), line_family) — needs no GPU."""
    lib = L.load()
    pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
    cn = np.ascontiguousarray(conn, dtype=np.int64).reshape(-1, 2)
    nl, nn = C.c_int64(), C.c_int64()
    args = (len(pts), len(cn), L.ptr(cn), L.ptr(pts), float(cos_tol), int(min_nodes), C.byref(nl), C.byref(nn))
    rc = lib.femb_symbolic_lines(*args, None, None, None, None)
    if rc:
        raise L.FembError(rc, "femb_symbolic_lines")
    line_ptr = np.zeros(nl.value + 1, dtype=np.int32)
    line_nodes = np.zeros(max(nn.value, 1), dtype=np.int32)
    line_dir = np.zeros((max(nl.value, 1), 3))
    fam = np.zeros(max(nl.value, 1), dtype=np.int32)
    rc = lib.femb_symbolic_lines(*args, L.ptr(line_ptr), L.ptr(line_nodes), L.ptr(line_dir), L.ptr(fam))
    if rc:
        raise L.FembError(rc, "femb_symbolic_lines")
    return line_ptr, line_nodes[:nn.value], line_dir[:nl.value], fam[:nl.value]


def symbolic_line_bundles(points, conn, target_per_family=768):
    """Host-only: symbolic phase of PRECOND_LINES (csrc/coarse.cpp build_line_symbolic) — needs no GPU.
    Returns dict(node_bundle / node_pos / node_line (3,N), node_dir (3,N,3), fam_off (4,), n_lines, n_entries, coverage)."""
    lib = L.load()
    pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
    cn = np.ascontiguousarray(conn, dtype=np.int64).reshape(-1, 2)
    nb = np.full((3, len(pts)), -1, dtype=np.int32)
    pos = np.full((3, len(pts)), -1, dtype=np.int32)
    fam_off = np.zeros(4, dtype=np.int32)
    line = np.full((3, len(pts)), -1, dtype=np.int32)
    ndir = np.zeros((3, len(pts), 3))
    nl, ne, cov = C.c_int64(), C.c_int64(), C.c_double()
    rc = lib.femb_symbolic_line_bundles(len(pts), len(cn), L.ptr(cn), L.ptr(pts), int(target_per_family), L.ptr(nb),
                                        L.ptr(pos), L.ptr(fam_off), C.byref(nl), C.byref(ne), C.byref(cov), L.ptr(line),
                                        L.ptr(ndir))
    if rc:
        raise L.FembError(rc, "femb_symbolic_line_bundles")
    return {"node_bundle": nb, "node_pos": pos, "node_line": line, "node_dir": ndir, "fam_off": fam_off,
            "n_lines": nl.value, "n_entries": ne.value, "coverage": cov.value}
