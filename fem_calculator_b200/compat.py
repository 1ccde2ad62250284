"""Reference-shaped host side: the same names, inputs, outputs and error behaviour as the
reference's Python entry points for the hot path, with every numerical step executed by
libfemb200.so on the GPU.

  BeamAnalysisWindow.run_simulation        BeamSolver.py:345-465  -> run_simulation_b200 /
                                                                     BeamAnalysisB200
  get_timoshenko_stiffness_matrix          BeamSolver.py:646-660  -> BeamAnalysisB200.<same>
  get_lumped_mass_matrix                   BeamSolver.py:662-675  -> BeamAnalysisB200.<same>
  bc_nodes_indexing                        BeamSolver.py:677-686  -> BeamAnalysisB200.<same>
  ForceAnalysis                            ReactionSolver.py:16-232 -> ForceAnalysisB200

Only the group -> node -> DOF bookkeeping (a few numpy index operations that define the
input layout) stays on the host.  There is no CPU fallback: without a CUDA device every
numerical call raises, which the reference's callers surface through their existing
``except Exception`` dialogs (BeamSolver.py:464-465, FEM_main.py:378-382).
"""
from __future__ import annotations

import warnings

import numpy as np

from . import _lib as L
from .api import FrameModel, Tet10Model
from .msh import Mesh, read_msh
from .sections import calculate_section_properties as _closed_form_sections

RHO_LITERAL = 7850  # BeamSolver.py:376 — the UI's density field is ignored by the reference


def _group_nodes(mesh, element_type, name):
    """BeamSolver.py:677-686 / ReactionSolver.py:75-85 (KeyError -> empty array)."""
    try:
        cells = mesh.cells_dict.get(element_type)
        if cells is None:
            return np.array([], dtype=int)
        phys = mesh.cell_data_dict.get("gmsh:physical", {}).get(element_type)
        if phys is None:
            return np.array([], dtype=int)
        target = mesh.field_data[name][0]
        sel = np.asarray(cells)[np.asarray(phys) == target].reshape(-1)
        if sel.size == 0 or sel.min() < 0:
            return np.unique(sel)
        # sorted unique node ids without a sort: mark and collect (the groups of a 1M-DOF frame hold 1e4-1e5 vertices
        # and this runs inside every run_simulation call)
        mark = np.zeros(int(sel.max()) + 1, dtype=bool)
        mark[sel] = True
        return np.flatnonzero(mark).astype(sel.dtype, copy=False)
    except (KeyError, IndexError):
        return np.array([], dtype=int)


def frame_bc_vectors(mesh, bc_data, num_nodes):
    """Up_node (sorted unique fixed DOFs) and the load vector f, BeamSolver.py:395-409."""
    f = np.zeros(6 * num_nodes)
    up = []
    for bc in bc_data:
        nodes = _group_nodes(mesh, "vertex", bc["group"])
        if bc["type"] == "Fix":
            for c, key in enumerate(("fix_x", "fix_y", "fix_z", "fix_rx", "fix_ry", "fix_rz")):
                if bc.get(key):
                    up.append(6 * nodes + c)
        elif bc["type"] == "Force":
            f[6 * nodes + 0] += bc.get("force_x", 0)
            f[6 * nodes + 1] += bc.get("force_y", 0)
            f[6 * nodes + 2] += bc.get("force_z", 0)
    fixed = np.unique(np.concatenate(up)).astype(np.int64) if up else np.zeros(0, dtype=np.int64)
    return fixed, f


def frame_section_table(mesh, section_data, props_fn):
    """Per-element section index + (S,8) table, BeamSolver.py:356-371.  Returns
    (elem_sec, props, missing_group_name or None)."""
    props_map = {s["group"]: props_fn(s["type"], s["params"], s.get("rotate", False)) for s in section_data}
    gid2name = {int(v[0]): k for k, v in mesh.field_data.items()}
    tags = np.asarray(mesh.cell_data_dict["gmsh:physical"]["line"])
    names = list(props_map.keys())
    lut = {}
    if tags.size and tags.dtype.kind in "iu" and tags.min() >= 0 and tags.max() < (1 << 22):
        present = np.flatnonzero(np.bincount(tags))            # distinct tags without sorting half a million entries
    else:
        present = np.unique(tags)
    for t in present:
        name = gid2name.get(int(t))
        if not name or name not in props_map:
            return None, None, name
        lut[int(t)] = names.index(name)
    lut_arr = np.full((max(lut) + 1) if lut else 1, 0, dtype=np.int32)
    for t, sidx in lut.items():
        lut_arr[t] = sidx
    elem_sec = lut_arr[tags].astype(np.int32)
    props = np.asarray([props_map[n] for n in names], dtype=np.float64).reshape(-1, 8)
    return elem_sec, props, None


_warned_closed_form = [False]


def default_props_fn():
    """The section front end used when the caller passes none: the reference's own
    ``calculate_section_properties`` (BeamSolver.py:32, a sectionproperties warping analysis) when the reference
    module and sectionproperties are importable — e.g. inside the patched application — else the closed forms of
    ``sections.py`` with a one-time warning (J and the shear areas of thin-walled shapes differ by a few %)."""
    try:
        import sectionproperties  # noqa: F401
        import BeamSolver
        return BeamSolver.calculate_section_properties
    except Exception:
        if not _warned_closed_form[0]:
            _warned_closed_form[0] = True
            warnings.warn("sectionproperties / BeamSolver not importable: using the closed-form section properties of "
                          "fem_calculator_b200.sections (pass props_fn=calculate_section_properties to use the reference's)",
                          RuntimeWarning, stacklevel=3)
        return _closed_form_sections


def run_simulation_b200(window, k_modes=20, device=0, props_fn=None, solver=L.SOLVER_AUTO, rtol=1e-12,
                        modal_rtol=1e-8, on_error=None):
    """Drop-in body for BeamAnalysisWindow.run_simulation (BeamSolver.py:345-455).

    Reads exactly what the reference reads (window.mesh / points / conn / section_data /
    bc_data / young_input.text() / poisson_input.text()) and writes exactly what it writes
    (window.u, window.smoothed_stresses, window.natural_frequencies [rad/s],
    window.mode_shapes (6N, n_modes)).  ``k_modes`` lowest modes are returned instead of
    all n_f (the report prints 10, the plots use 5: BeamSolver.py:540-547,575).
    ``on_error(title, message)`` stands in for QMessageBox.critical; default raises.
    """
    def err(title, msg):
        if on_error is not None:
            on_error(title, msg)
            return None
        raise RuntimeError(f"{title}: {msg}")

    if not window.mesh:
        return err("Error", "Please load a mesh file first.")
    E = float(window.young_input.text())
    nu = float(window.poisson_input.text())
    G = E / (2 * (1 + nu))
    num_nodes = len(window.points)
    fn = props_fn or default_props_fn()
    elem_sec, props, missing = frame_section_table(window.mesh, window.section_data, fn)
    if elem_sec is None:
        return err("Error", f"Section properties not defined for physical group '{missing}'.")
    fixed, f = frame_bc_vectors(window.mesh, window.bc_data, num_nodes)

    import os, time
    trace = os.environ.get("FEMB_TRACE")
    t_ = [time.perf_counter()]

    def lap(name):
        if trace:
            t_.append(time.perf_counter())
            print(f"[femb trace] {name}: {(t_[-1] - t_[-2]) * 1e3:.1f} ms", flush=True)

    lap("host bookkeeping (section table, BC vectors)")
    # the handle (device, stream, HBM buffers) lives with the window and is reused by later runs:
    # only the first run_simulation pays for context creation and the device allocations
    model = getattr(window, "_femb_model", None)
    if model is None or getattr(model, "_h", None) is None or model.device != device:
        model = FrameModel(device)
        try:
            window._femb_model = model
        except Exception:
            pass
    lap("femb_create / reuse")
    try:
        model.set_mesh(window.points, window.conn, elem_sec, props, E, G, RHO_LITERAL)
        lap("set_mesh")
        model.assemble()
        lap("assemble (symbolic + numeric)")
        model.set_bc(fixed, f)
        lap("set_bc")
        u, reactions, st = model.solve_static(method=solver, rtol=rtol)
        lap("solve_static (+ D2H of u, reactions)")
        window.u = u
        window.reaction_forces = reactions          # extra: K u - f (the reference computes none)
        window.smoothed_stresses = model.stress()
        lap("stress")
        window.solve_stats = st
        if k_modes:
            try:
                lam, phi, mst = model.modal(k=int(k_modes), rtol=modal_rtol)
            except L.FembError as e:
                if e.code != L.FEMB_ERR_NOT_CONVERGED:
                    raise
                # ill-conditioned chains (cond(K) eps > modal_rtol): take what FP64 can resolve, and say so
                lam, phi, mst = model.modal(k=int(k_modes), rtol=modal_rtol, accept_rtol=1e-4)
                warnings.warn(f"modal solve stagnated at a pencil residual of {mst['rel_residual']:.1e} (> modal_rtol = "
                              f"{modal_rtol:.1e}); eigenpairs are returned at that accuracy", RuntimeWarning, stacklevel=2)
            window.natural_frequencies = np.sqrt(lam)      # BeamSolver.py:451
            window.mode_shapes = phi                        # zeros on fixed DOFs, :453-455
            window.modal_stats = mst
    finally:
        if getattr(window, "_femb_model", None) is not model:
            model.close()
        lap("close / keep")
    return window


class _Text:
    def __init__(self, v):
        self._v = v

    def text(self):
        return repr(float(self._v))


class BeamAnalysisB200:
    """Headless stand-in for BeamAnalysisWindow (BeamSolver.py:176) exposing the hot-path
    attributes and helper methods with the reference's names and signatures."""

    def __init__(self, mesh=None, section_data=None, bc_data=None, E=2e11, nu=0.3, device=0, props_fn=None):
        self.mesh = mesh
        self.points = None if mesh is None else mesh.points
        self.conn = None if mesh is None else mesh.cells_dict.get("line")
        self.section_data = list(section_data or [])
        self.bc_data = list(bc_data or [])
        self.young_input, self.poisson_input = _Text(E), _Text(nu)
        self.u = self.smoothed_stresses = self.natural_frequencies = self.mode_shapes = None
        self.device = device
        self.props_fn = props_fn

    def mesh_load(self, mesh_path):
        """BeamSolver.py:207-220 without the file dialog."""
        self.mesh = read_msh(mesh_path)
        self.points = self.mesh.points
        self.conn = self.mesh.cells_dict.get("line")
        if self.conn is None:
            raise ValueError("No 'line' elements in .msh file.")
        self.section_data.clear()
        self.bc_data.clear()

    def run_simulation(self, k_modes=20, **kw):
        return run_simulation_b200(self, k_modes=k_modes, device=self.device, props_fn=self.props_fn, **kw)

    def close(self):
        """Release the GPU handle kept between runs."""
        for name in ("_femb_model", "_elem_model"):
            m = getattr(self, name, None)
            if m is not None:
                m.close()
                setattr(self, name, None)

    def bc_nodes_indexing(self, element_type, bc_name):
        return _group_nodes(self.mesh, element_type, bc_name)

    # ---- result consumers (the step after the path): the two tables of the reference's report as arrays / CSV
    def nodal_result_table(self):
        """Rows of the 'Nodal displacement and stress results' table (BeamSolver.py:530-538) as an
        (N, 8) float array: node id, x, y, z [m], u_x, u_y, u_z [m], smoothed stress [MPa]."""
        if self.u is None or self.smoothed_stresses is None:
            raise RuntimeError("run_simulation first")
        pts = np.asarray(self.points, dtype=np.float64)
        disp = np.asarray(self.u, dtype=np.float64).reshape(-1, 6)[:, :3]
        return np.column_stack([np.arange(len(pts)), pts, disp, np.asarray(self.smoothed_stresses) / 1e6])

    def modal_result_table(self, n_modes=10):
        """Rows of the 'Modal Analysis Results' table (BeamSolver.py:540-547): mode number, frequency in
        rad/s and in Hz for the first ``n_modes`` (the report prints 10) — (m, 3) float array."""
        if self.natural_frequencies is None:
            return np.zeros((0, 3))
        w = np.asarray(self.natural_frequencies, dtype=np.float64)[:n_modes]
        return np.column_stack([np.arange(1, len(w) + 1), w, w / (2 * np.pi)])

    def write_result_tables(self, prefix):
        """The two tables as CSV with the reference's column titles and number formats
        (``<prefix>_nodes.csv``, ``<prefix>_modes.csv``); returns the paths written."""
        paths = []
        nodal = self.nodal_result_table()
        with open(f"{prefix}_nodes.csv", "w") as fh:
            fh.write("Node ID,X (m),Y (m),Z (m),Disp X (m),Disp Y (m),Disp Z (m),Stress (MPa)\n")
            for r in nodal:
                fh.write(f"{int(r[0])},{r[1]:.4f},{r[2]:.4f},{r[3]:.4f},{r[4]:.4e},{r[5]:.4e},{r[6]:.4e},{r[7]:.4f}\n")
        paths.append(f"{prefix}_nodes.csv")
        modal = self.modal_result_table()
        if len(modal):
            with open(f"{prefix}_modes.csv", "w") as fh:
                fh.write("Mode,Frequency (rad/s),Frequency (Hz)\n")
                for r in modal:
                    fh.write(f"{int(r[0])},{r[1]:.4f},{r[2]:.4f}\n")
            paths.append(f"{prefix}_modes.csv")
        return paths

    def qr_algorithm(self, A, max_iter=1000, tol=1e-9):
        """BeamSolver.py:467-481 — the unshifted QR iteration on A = inv(M_ff) K_ff, same signature and return value
        (eigenvalues ascending, the matching columns of the accumulated Q), evaluated on the GPU.  Kept for callers
        that use the helper directly; run_simulation does not go through it (femb_modal solves the symmetric pencil
        for the lowest modes).  This helper is dense O(n^3) per sweep by definition, so it runs on library kernels
        (torch.linalg.qr / matmul in FP64 on the handle's device) rather than on hand-written ones; the stopping test
        is the reference's np.allclose(diag, diag_new, atol=tol) with numpy's default rtol = 1e-5.  No CPU fallback."""
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("qr_algorithm: no CUDA device (femb200 has no CPU fallback)")
        dev = torch.device("cuda", self.device)
        Ak = torch.as_tensor(np.asarray(A, dtype=np.float64), device=dev).clone()
        n = Ak.shape[0]
        V = torch.eye(n, dtype=torch.float64, device=dev)
        for _ in range(int(max_iter)):
            Q, R = torch.linalg.qr(Ak)
            A_new = R @ Q
            V = V @ Q
            d0, d1 = torch.diagonal(Ak), torch.diagonal(A_new)
            done = bool(torch.all(torch.abs(d0 - d1) <= tol + 1e-5 * torch.abs(d1)))
            Ak = A_new
            if done:
                break
        eigenvalues = torch.diagonal(Ak)
        idx = torch.argsort(eigenvalues)
        return eigenvalues[idx].cpu().numpy(), V[:, idx].cpu().numpy()

    def _one_element(self, L_, E, G, props, rho, want_k, want_m):
        # one small handle kept for the helper calls (creating a CUDA handle per 12x12 matrix cost ~100 ms a call)
        m = getattr(self, "_elem_model", None)
        if m is None or getattr(m, "_h", None) is None:
            m = self._elem_model = FrameModel(self.device)
        pts = np.array([[0.0, 0.0, 0.0], [float(L_), 0.0, 0.0]])
        m.set_mesh(pts, np.array([[0, 1]]), np.zeros(1, dtype=np.int32), np.asarray(props, dtype=np.float64), E, G, rho)
        return m.elements(want_k, want_m)

    def get_timoshenko_stiffness_matrix(self, L, E, G, A, I_x, I_y, J, kappa_y, kappa_z):
        """BeamSolver.py:646 — local 12x12 stiffness, evaluated by the CUDA element kernel on
        an x-aligned element (lambda = I, so global == local)."""
        ke, _ = self._one_element(L, E, G, [A, I_x, I_y, J, kappa_y, kappa_z, 0.0, 0.0], 0.0, True, False)
        return ke[0]

    def get_lumped_mass_matrix(self, L, A, Ix, Iy, J, rho):
        """BeamSolver.py:662."""
        _, me = self._one_element(L, 1.0, 1.0, [A, Ix, Iy, J, 0.0, 0.0, 0.0, 0.0], rho, False, True)
        return me[0]


class ForceAnalysisB200:
    """ForceAnalysis (ReactionSolver.py:16) with the numerical methods on the GPU.  Same
    constructor, attributes and method names; ``msh_file`` may also be an in-memory mesh."""

    def __init__(self, msh_file, force_data, fix_data, E, v, device=0):
        self.msh_file = msh_file
        self.force_data = force_data
        self.fix_data = fix_data
        self.E = E
        self.v = v
        self.pd = 3
        self.points = None
        self.tetra10_conn = None
        self.num_nodes = 0
        self.K = None
        self.u = None
        self.f = None
        self.reaction_forces = None
        self.fixed_nodes_info = []
        self.negative_detJ_count = 0
        self.applied_forces_info = []
        self.device = device
        self._model = None
        self._read_mesh()

    def _read_mesh(self):
        """ReactionSolver.py:59-73."""
        self.mesh = read_msh(self.msh_file) if isinstance(self.msh_file, str) else self.msh_file
        self.points = self.mesh.points
        self.num_nodes = len(self.points)
        self.tetra10_conn = self.mesh.cells_dict.get("tetra10")
        if self.tetra10_conn is None:
            raise ValueError("메쉬 파일에 'tetra10' 요소가 없습니다.")  # same message as :68
        self.diri_nodes = self._get_nodes_from_physical_group("Diri_BCs", "vertex")
        self.neumann_nodes = self._get_nodes_from_physical_group("Neumann_BCs", "vertex")

    def _get_nodes_from_physical_group(self, group_name, cell_type):
        return _group_nodes(self.mesh, cell_type, group_name)

    def _ensure_model(self):
        if self._model is None:
            self._model = Tet10Model(self.device)
            self._model.set_mesh(self.points, self.tetra10_conn, self.E, self.v)
        return self._model

    def assemble_stiffness_matrix(self, export_csr=True):
        """ReactionSolver.py:115-152.  K stays resident in HBM; ``export_csr`` also copies
        it back as a scipy CSR matrix into ``self.K`` like the reference (structural
        pattern: exact zeros are kept)."""
        m = self._ensure_model()
        m.assemble()
        self.negative_detJ_count = m.negative_detj
        if export_csr:
            indptr, indices, data = m.get_csr(L.MAT_K)
            try:
                import scipy.sparse as sp
                n = self.pd * self.num_nodes
                self.K = sp.csr_matrix((data, indices, indptr), shape=(n, n))
            except ImportError:
                self.K = (indptr, indices, data)

    def _nearest_in_group(self, group_nodes, targets):
        """Index (into the mesh) of the group node closest to every target point: one distance table for all points."""
        pts = np.asarray(self.points, dtype=np.float64)[group_nodes]
        d2 = ((pts[None, :, :] - np.asarray(targets, dtype=np.float64).reshape(-1, 1, 3)) ** 2).sum(axis=2)
        return np.asarray(group_nodes)[np.argmin(d2, axis=1)]

    def apply_boundary_conditions(self):
        """Same inputs and outputs as ReactionSolver.py:154-194 (f, fixed_dofs, active_dofs, fixed_nodes_info,
        applied_forces_info): every fix / force point is snapped to the nearest node of its physical group, a DOF is
        fixed iff its flag is 0.  Vectorised: one distance table per group instead of a loop over the points."""
        total_dof = self.pd * self.num_nodes
        fix_pts = np.array([[d["pos_x"], d["pos_y"], d["pos_z"]] for d in self.fix_data], dtype=np.float64).reshape(-1, 3)
        fix_free = np.array([[d["fix_x"] == 0, d["fix_y"] == 0, d["fix_z"] == 0] for d in self.fix_data], dtype=bool).reshape(-1, 3)
        fix_nodes = self._nearest_in_group(self.diri_nodes, fix_pts) if len(fix_pts) else np.zeros(0, dtype=np.int64)
        dof_table = 3 * fix_nodes[:, None] + np.arange(3)[None, :]
        self.fixed_dofs = np.unique(dof_table[fix_free])
        self.fixed_nodes_info = [{"node_idx": n, "pos": self.points[n], "dofs": [int(v) for v in row[keep]]}
                                 for n, row, keep in zip(fix_nodes, dof_table, fix_free)]
        frc_pts = np.array([[d["force_x_pstn"], d["force_y_pstn"], d["force_z_pstn"]] for d in self.force_data], dtype=np.float64).reshape(-1, 3)
        frc_vec = np.array([[d["force_x"], d["force_y"], d["force_z"]] for d in self.force_data], dtype=np.float64).reshape(-1, 3)
        frc_nodes = self._nearest_in_group(self.neumann_nodes, frc_pts) if len(frc_pts) else np.zeros(0, dtype=np.int64)
        f3 = np.zeros((self.num_nodes, 3))
        np.add.at(f3, frc_nodes, frc_vec)
        self.f = f3.reshape(total_dof)
        self.applied_forces_info = [{"node_idx": n, "pos": self.points[n], "force_vec": v} for n, v in zip(frc_nodes, frc_vec)]
        free = np.ones(total_dof, dtype=bool)
        free[self.fixed_dofs.astype(np.int64)] = False
        self.active_dofs = np.flatnonzero(free)
        self._ensure_model().set_bc(self.fixed_dofs.astype(np.int64), self.f)

    def solve(self, method=L.SOLVER_AUTO, rtol=1e-12):
        """ReactionSolver.py:196-205: u on active DOFs, reaction_forces = K_full @ u."""
        m = self._ensure_model()
        self.u, self.reaction_forces, self.solve_stats = m.solve_static(method=method, rtol=rtol, minus_f=False)

    def reaction_table(self):
        """(node ids (m,), reactions (m, 3)) at the fix points, in fix_data order — the rows of the reference's
        reaction print-out (ReactionSolver.py:207-224) and report table as arrays."""
        nodes = np.array([info["node_idx"] for info in self.fixed_nodes_info], dtype=np.int64)
        R = np.asarray(self.reaction_forces, dtype=np.float64).reshape(-1, 3)[nodes] if len(nodes) else np.zeros((0, 3))
        return nodes, R

    def print_reactions(self):
        """The reaction / equilibrium print-out of ReactionSolver.py:207-224 (same lines and number formats), built
        from reaction_table()."""
        if self.reaction_forces is None:
            return
        nodes, R = self.reaction_table()
        out = ["", "--- Reaction Forces ---"]
        out += [f"  Node {n} (Fix Point {k + 1}): Rx={r[0]:.4e}, Ry={r[1]:.4e}, Rz={r[2]:.4e} N" for k, (n, r) in enumerate(zip(nodes, R))]
        applied = np.array([[d["force_x"], d["force_y"], d["force_z"]] for d in self.force_data], dtype=np.float64).reshape(-1, 3).sum(axis=0)
        out += ["", "--- Force Equilibrium Check ---", f"  Sum of Applied Forces (Fx, Fy, Fz): {applied}",
                f"  Sum of Reaction Forces (Rx, Ry, Rz): {-R.sum(axis=0)}"]
        print("\n".join(out))

    def run_simulation(self):
        """ReactionSolver.py:226-232 minus the docx report (out of scope)."""
        self.assemble_stiffness_matrix()
        self.apply_boundary_conditions()
        self.solve()
        self.print_reactions()

    def close(self):
        if self._model is not None:
            self._model.close()
            self._model = None


def accelerate_force_analysis(reference_cls):
    """``ForceAnalysis`` of the untouched application (ReactionSolver.py:16) with its three numerical methods moved
    to the GPU: a subclass that INHERITS the reference's own _read_mesh, apply_boundary_conditions, print_reactions,
    plot and generate_report and overrides only assemble_stiffness_matrix / solve (plus a hook that hands the BC
    vectors the reference computed to the device).  ``FEM_main.py:14`` then reads
    ``ForceAnalysis = accelerate_force_analysis(ForceAnalysis)``."""

    class ForceAnalysisGPU(reference_cls):
        device = 0
        _model = None

        _ensure_model = ForceAnalysisB200._ensure_model
        assemble_stiffness_matrix = ForceAnalysisB200.assemble_stiffness_matrix
        close = ForceAnalysisB200.close

        def apply_boundary_conditions(self):
            super().apply_boundary_conditions()                       # the reference's own bookkeeping
            self._ensure_model().set_bc(np.asarray(self.fixed_dofs, dtype=np.int64), self.f)

        solve = ForceAnalysisB200.solve

    ForceAnalysisGPU.__name__ = reference_cls.__name__ + "GPU"
    return ForceAnalysisGPU
