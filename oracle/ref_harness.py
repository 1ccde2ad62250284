"""TEST INFRASTRUCTURE ONLY — never imported by the product path.

Headless driver for the UNMODIFIED reference (``/root/reference/BeamSolver.py`` and
``ReactionSolver.py``).  The reference cannot be imported as-is in this container
(PyQt5, meshio, matplotlib, sectionproperties, pyvista, tkinter are absent and
``uic.loadUiType('Beam_analysis.ui')`` runs at import, BeamSolver.py:3-25), so this
module injects inert stand-ins for those GUI / IO modules into ``sys.modules`` and
then executes the reference source verbatim.  Nothing numerical is stubbed: the
reference's own numpy / scipy arithmetic runs (BeamSolver.py:345-481, 646-686;
ReactionSolver.py:87-205).

It exists to (a) validate ``oracle/ref_sparse.py`` against the real reference and
(b) generate the committed golden vectors under ``tests/golden/``
(``oracle/make_golden.py``).  ``/root/reference`` does not exist on the GPU box, so
nothing that runs there may import this file; callers must check ``available()``.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np

REFERENCE_DIR = os.environ.get("FEMB_REFERENCE_DIR", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "BeamSolver.py"))


class _Anything:
    """Inert object: any attribute / call / item access yields another _Anything."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()

    def __getitem__(self, k):
        return _Anything()

    def __iter__(self):
        return iter(())

    def __or__(self, other):
        return self

    def __bool__(self):
        return False


class _QMessageBox:
    """Records dialogs instead of showing them; question() answers No so that
    run_simulation skips create_report (BeamSolver.py:461)."""

    Yes, No = 1, 0
    log: list = []

    @classmethod
    def _rec(cls, kind, args):
        cls.log.append((kind,) + tuple(str(a) for a in args[1:]))

    @classmethod
    def critical(cls, *a, **k):
        cls._rec("critical", a)

    @classmethod
    def warning(cls, *a, **k):
        cls._rec("warning", a)

    @classmethod
    def information(cls, *a, **k):
        cls._rec("information", a)

    @classmethod
    def question(cls, *a, **k):
        cls._rec("question", a)
        return cls.No


def _stub_module(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)

    def _fallback(attr):
        if attr.startswith("__"):
            raise AttributeError(attr)
        return _Anything

    m.__getattr__ = _fallback  # PEP 562
    return m


_STUB_NAMES = [
    "PyQt5", "PyQt5.uic", "PyQt5.QtWidgets", "PyQt5.QtCore", "PyQt5.QtGui",
    "matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d",
    "meshio", "tkinter", "tkinter.filedialog", "pyvista",
    "sectionproperties", "sectionproperties.pre", "sectionproperties.pre.library",
    "sectionproperties.pre.library.steel_sections",
    "sectionproperties.pre.library.primitive_sections",
    "sectionproperties.pre.pre", "sectionproperties.analysis",
    "sectionproperties.analysis.section",
    "docx", "docx.shared", "docx.enum", "docx.enum.text", "gmsh",
]


class _QDialogBase:
    def __init__(self, *a, **k):
        pass


def _install_stubs():
    saved = {n: sys.modules.get(n) for n in _STUB_NAMES}
    for n in _STUB_NAMES:
        sys.modules[n] = _stub_module(n)
    for n in _STUB_NAMES:  # `import a.b.c as x` resolves through parent attributes
        if "." in n:
            parent, child = n.rsplit(".", 1)
            setattr(sys.modules[parent], child, sys.modules[n])
    uic = sys.modules["PyQt5.uic"]
    uic.loadUiType = lambda *a, **k: (type("Ui_Dialog", (), {}), _QDialogBase)
    sys.modules["PyQt5"].uic = uic
    qtw = sys.modules["PyQt5.QtWidgets"]
    qtw.QMessageBox = _QMessageBox
    qtw.QDialog = _QDialogBase
    qtw.QWidget = _QDialogBase
    # ReactionSolver.py:9-14 — make python-docx "missing" so DOCX_AVAILABLE=False
    for n in ("docx", "docx.shared", "docx.enum", "docx.enum.text"):
        del sys.modules[n]
    return saved


def _restore(saved):
    for n, m in saved.items():
        if m is None:
            sys.modules.pop(n, None)
        else:
            sys.modules[n] = m


_cache: dict = {}


def _load(fname: str):
    """Execute an unmodified reference source file under the stub modules."""
    if fname in _cache:
        return _cache[fname]
    if not available():
        raise RuntimeError(f"reference not present at {REFERENCE_DIR}")
    saved = _install_stubs()

    class _BlockDocx:
        # a meta-path finder that makes `import docx` raise ImportError
        @staticmethod
        def find_spec(name, path=None, target=None):
            if name == "docx" or name.startswith("docx."):
                raise ImportError("docx blocked by oracle harness")
            return None

    sys.meta_path.insert(0, _BlockDocx)
    try:
        path = os.path.join(REFERENCE_DIR, fname)
        spec = importlib.util.spec_from_file_location("_femref_" + fname[:-3], path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.meta_path.remove(_BlockDocx)
        _restore(saved)
    _cache[fname] = mod
    return mod


class MeshRecord:
    """meshio-shaped in-memory mesh: exactly the attributes the reference reads
    (BeamSolver.py:212-217,357-358,677-686; ReactionSolver.py:62-85)."""

    def __init__(self, points, cells_dict, field_data, cell_physical):
        self.points = np.asarray(points, dtype=float)
        self.cells_dict = {k: np.asarray(v) for k, v in cells_dict.items()}
        self.field_data = {k: np.asarray(v) for k, v in field_data.items()}
        self.cell_data_dict = {"gmsh:physical": {k: np.asarray(v) for k, v in cell_physical.items()}}


class _LineEdit:
    def __init__(self, value):
        self._v = value

    def text(self):
        return repr(float(self._v))


def run_beam_reference(mesh: MeshRecord, section_props: dict, bc_data: list, E: float, nu: float,
                       with_modal: bool = True):
    """Run BeamAnalysisWindow.run_simulation (BeamSolver.py:345-457) verbatim.

    ``section_props``: group name -> 8-tuple (A, I_x, I_y, J, kappa_y, kappa_z,
    c_y_max, c_z_max), i.e. the return record of calculate_section_properties
    (BeamSolver.py:79) — sectionproperties itself is out of scope / not installed,
    so the module-level function is replaced by a lookup.

    Returns dict(u, smoothed_stresses, natural_frequencies, mode_shapes, dialogs).
    With ``with_modal=False`` the O(n^3)x1000 qr_algorithm (BeamSolver.py:467) is
    replaced by a stub returning no modes so the static part can run at larger N.
    """
    mod = _load("BeamSolver.py")
    win = object.__new__(mod.BeamAnalysisWindow)
    win.mesh = mesh
    win.points = mesh.points
    win.conn = mesh.cells_dict["line"]
    win.section_data = [{"group": g, "type": "__lookup__", "params": {"group": g}, "rotate": False}
                        for g in section_props]
    win.bc_data = bc_data
    win.young_input = _LineEdit(E)
    win.poisson_input = _LineEdit(nu)
    win.u = win.smoothed_stresses = win.natural_frequencies = win.mode_shapes = None
    win._generate_plots_for_report = lambda: None
    orig_csp = mod.calculate_section_properties
    mod.calculate_section_properties = lambda t, p, r=False: tuple(section_props[p["group"]])
    if not with_modal:
        win.qr_algorithm = lambda A, max_iter=1000, tol=1e-9: (np.zeros(0), np.zeros((A.shape[0], 0)))
    _QMessageBox.log = []
    try:
        win.run_simulation()
    finally:
        mod.calculate_section_properties = orig_csp
    dialogs = list(_QMessageBox.log)
    crit = [d for d in dialogs if d[0] == "critical"]
    if crit:
        raise RuntimeError(f"reference raised a critical dialog: {crit}")
    return {
        "u": win.u,
        "smoothed_stresses": win.smoothed_stresses,
        "natural_frequencies": win.natural_frequencies,
        "mode_shapes": win.mode_shapes,
        "dialogs": dialogs,
    }


def beam_reference_helpers():
    """The reference's own helper methods, bound to a bare window object."""
    mod = _load("BeamSolver.py")
    win = object.__new__(mod.BeamAnalysisWindow)
    return win.get_timoshenko_stiffness_matrix, win.get_lumped_mass_matrix, win.qr_algorithm


def beam_reference_matrices(mesh: MeshRecord, section_props: dict, E: float, nu: float, rho: float = 7850):
    """Dense K, M exactly as the loop at BeamSolver.py:360-393 builds them.  That loop
    is inline in run_simulation (no callable), so it is replayed here through the
    reference's own element helpers and numpy calls in the same order."""
    kfun, mfun, _ = beam_reference_helpers()
    pts = mesh.points
    conn = mesh.cells_dict["line"]
    n = len(pts)
    G = E / (2 * (1 + nu))
    gid = {v[0]: k for k, v in mesh.field_data.items()}
    tags = mesh.cell_data_dict["gmsh:physical"]["line"]
    K = np.zeros((6 * n, 6 * n))
    M = np.zeros((6 * n, 6 * n))
    eps = 1e-6
    for i, el in enumerate(conn):
        A, I_x, I_y, J, ky, kz, _, _ = section_props[gid[tags[i]]]
        p1, p2 = pts[el[0]], pts[el[1]]
        L = np.linalg.norm(p2 - p1)
        k_ = kfun(L, E, G, A, I_x, I_y, J, ky, kz)
        m_ = mfun(L, A, I_x, I_y, J, rho)
        Cxx, Cyx, Czx = (p2 - p1) / L
        if Cxx**2 + Cyx**2 < eps**2:
            lam = np.array([[0., 0., 1. if Czx > 0 else -1.], [0., 1., 0.], [-1. if Czx > 0 else 1., 0., 0.]])
        else:
            D = np.sqrt(Cxx**2 + Cyx**2)
            lam = np.array([[Cxx, Cyx, Czx], [-Cyx / D, Cxx / D, 0], [-Cxx * Czx / D, -Cyx * Czx / D, D]])
        R = np.kron(np.eye(4, dtype=float), lam)
        kl = R.T @ k_ @ R
        ml = R.T @ m_ @ R
        for j, Jn in enumerate(el):
            for l, Ln in enumerate(el):
                K[6 * Jn:6 * Jn + 6, 6 * Ln:6 * Ln + 6] += kl[6 * j:6 * j + 6, 6 * l:6 * l + 6]
                M[6 * Jn:6 * Jn + 6, 6 * Ln:6 * Ln + 6] += ml[6 * j:6 * j + 6, 6 * l:6 * l + 6]
    return K, M


def run_tet10_reference(mesh: MeshRecord, force_data: list, fix_data: list, E: float, nu: float):
    """Run ForceAnalysis assemble/apply/solve (ReactionSolver.py:115-205) verbatim on an
    in-memory mesh (``meshio.read`` is the only thing replaced)."""
    mod = _load("ReactionSolver.py")
    orig = mod.meshio.read
    mod.meshio.read = lambda path: mesh
    import contextlib
    import io
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            fa = mod.ForceAnalysis("<in-memory>", force_data, fix_data, E, nu)
            fa.assemble_stiffness_matrix()
            fa.apply_boundary_conditions()
            fa.solve()
    finally:
        mod.meshio.read = orig
    return fa
