"""Synthetic meshes for the BASELINE.json configurations.

gmsh and meshio are not available, so the five benchmark configurations and the
secondary Tet10 box are generated here in the meshio-shaped record (``msh.Mesh``)
that the reference reads — node numbering, element order and physical groups are
part of the input contract (DOF = 6*node + c, BeamSolver.py:354,360), never
renumbered downstream.  Everything is deterministic; jitter uses
``np.random.default_rng(20261018)``.
"""
from __future__ import annotations

import numpy as np

from .msh import Mesh
from .sections import calculate_section_properties

E_STEEL, NU_STEEL, RHO_STEEL = 2.0e11, 0.3, 7850.0
SEED = 20261018


def _fix_bc(group, x=True, y=True, z=True, rx=True, ry=True, rz=True):
    return {"type": "Fix", "group": group, "fix_x": x, "fix_y": y, "fix_z": z,
            "fix_rx": rx, "fix_ry": ry, "fix_rz": rz}


def _force_bc(group, fx=0.0, fy=0.0, fz=0.0):
    return {"type": "Force", "group": group, "force_x": fx, "force_y": fy, "force_z": fz}


def chain_mesh(n_el: int, length: float, groups: dict, axis=(1.0, 0.0, 0.0)) -> Mesh:
    """Straight chain of n_el line elements, nodes 0..n_el in order along ``axis``.
    ``groups``: physical point groups, name -> list of node indices."""
    t = np.linspace(0.0, length, n_el + 1)
    ax = np.asarray(axis, dtype=float)
    ax = ax / np.linalg.norm(ax)
    pts = t[:, None] * ax[None, :]
    conn = np.stack([np.arange(n_el), np.arange(1, n_el + 1)], axis=1)
    field, vcells, vtags = {}, [], []
    tag = 1
    for name, nodes in groups.items():
        field[name] = [tag, 0]
        for n in nodes:
            vcells.append([n])
            vtags.append(tag)
        tag += 1
    field["beam"] = [tag, 1]
    return Mesh(pts, {"vertex": np.asarray(vcells, dtype=np.int64).reshape(-1, 1), "line": conn},
                field, {"vertex": vtags, "line": np.full(n_el, tag)})


def cantilever_case(n_el: int = 2, length: float = 2.0, P: float = -1000.0):
    """C1-like cantilever: rectangular d=0.1, b=0.05, all six DOFs fixed at node 0,
    tip load (0, P, 0) — the setup of the shipped ``cantilever_beam`` file."""
    mesh = chain_mesh(n_el, length, {"fix": [0], "load_y": [n_el]})
    sec = [{"group": "beam", "type": "rectangular section", "params": {"d": 0.1, "b": 0.05}, "rotate": False}]
    bc = [_fix_bc("fix"), _force_bc("load_y", fy=P)]
    return mesh, sec, bc


def simply_supported_case(n_el: int = 10000, length: float = 10.0):
    """C2: simply supported I-section beam, Euler-Bernoulli (kappa=0 ->
    BeamSolver.py:647-648 fallback).  pin: x,y,z,rx at node 0; roller: y,z,rx at the
    far end.  A mid-span load makes the static problem non-trivial."""
    mesh = chain_mesh(n_el, length, {"pin": [0], "roller": [n_el], "mid": [n_el // 2]})
    sec = [{"group": "beam", "type": "I section",
            "params": {"d": 0.2, "b": 0.1, "t_f": 0.0085, "t_w": 0.0056, "r": 0.0}, "rotate": False}]
    bc = [_fix_bc("pin", ry=False, rz=False), _fix_bc("roller", x=False, ry=False, rz=False),
          _force_bc("mid", fy=-1000.0, fz=-500.0)]
    return mesh, sec, bc


def euler_bernoulli(props):
    """8-tuple with shear coefficients zeroed -> Euler-Bernoulli element."""
    A, Ix, Iy, J, _, _, cy, cz = props
    return (A, Ix, Iy, J, 0.0, 0.0, cy, cz)


LATTICE_SECTIONS = [
    {"group": "beams_x", "type": "hollow box section", "params": {"d": 0.1, "b": 0.1, "t": 0.005, "r_out": 0.0}, "rotate": False},
    {"group": "beams_y", "type": "C section", "params": {"d": 0.1, "b": 0.05, "t_f": 0.005, "t_w": 0.005, "r": 0.0}, "rotate": False},
    {"group": "beams_z", "type": "L section", "params": {"d": 0.075, "b": 0.075, "t": 0.006, "r_r": 0.0, "r_t": 0.0}, "rotate": False},
]


def lattice_frame_case(nx: int, ny: int, nz: int, pitch: float = 1.0, jitter: float = 0.0,
                       load=(100.0, 0.0, -1000.0)):
    """C3 / C5: nx*ny*nz-node space frame, node index = (ix*ny + iy)*nz + iz (z
    fastest), members along x / y / z in three physical groups with box / C / L
    sections.  ``base`` = all z=0 nodes (fixed x6), ``top`` = all top-level nodes
    (each gets the full load vector — the reference's 'distributed' load,
    BeamSolver.py:406-407).  ``jitter`` (fraction of pitch) perturbs every node with
    the seeded RNG so the rotation matrices are generic (dense 6x6 blocks); with
    jitter=0 z-members take the vertical branch (BeamSolver.py:380-381)."""
    ix, iy, iz = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    pts = np.stack([ix, iy, iz], axis=-1).reshape(-1, 3).astype(np.float64) * pitch
    if jitter:
        rng = np.random.default_rng(SEED)
        pts = pts + rng.uniform(-jitter * pitch, jitter * pitch, size=pts.shape)
    nid = np.arange(nx * ny * nz, dtype=np.int64).reshape(nx, ny, nz)
    ex = np.stack([nid[:-1].ravel(), nid[1:].ravel()], axis=1)
    ey = np.stack([nid[:, :-1].ravel(), nid[:, 1:].ravel()], axis=1)
    ez = np.stack([nid[:, :, :-1].ravel(), nid[:, :, 1:].ravel()], axis=1)
    conn = np.concatenate([ex, ey, ez], axis=0)
    ltags = np.concatenate([np.full(len(ex), 3), np.full(len(ey), 4), np.full(len(ez), 5)])
    base = nid[:, :, 0].ravel()
    top = nid[:, :, -1].ravel()
    vcells = np.concatenate([base, top]).reshape(-1, 1)
    vtags = np.concatenate([np.full(len(base), 1), np.full(len(top), 2)])
    field = {"base": [1, 0], "top": [2, 0], "beams_x": [3, 1], "beams_y": [4, 1], "beams_z": [5, 1]}
    mesh = Mesh(pts, {"vertex": vcells, "line": conn}, field, {"vertex": vtags, "line": ltags})
    bc = [_fix_bc("base"), _force_bc("top", *load)]
    return mesh, [dict(s) for s in LATTICE_SECTIONS], bc


def batch_cantilever_params(n_models: int, seed: int = SEED):
    """C4: per-model parameters for the batched load-case sweep: rect
    section d~U[0.08,0.3], b~U[0.04,0.15]; tip F_y~U[-5e3,-5e2]; uniform nodal
    F_z~U[-50,-5].  Returns dict of (n_models,) arrays."""
    rng = np.random.default_rng(seed)
    return {
        "d": rng.uniform(0.08, 0.3, n_models),
        "b": rng.uniform(0.04, 0.15, n_models),
        "tip_fy": rng.uniform(-5e3, -5e2, n_models),
        "nodal_fz": rng.uniform(-50.0, -5.0, n_models),
    }


def section_table(mesh: Mesh, section_data: list, props_fn=calculate_section_properties):
    """Resolve the per-element section record the way BeamSolver.py:356-371 does:
    line physical tag -> group name -> 8-tuple.  Returns (elem_sec (E,) int32 index
    into props, props (S,8) float64).  Raises KeyError for an element whose group has
    no section (the reference aborts with a dialog, BeamSolver.py:367-369)."""
    props_map = {s["group"]: props_fn(s["type"], s["params"], s.get("rotate", False)) for s in section_data}
    gid2name = {int(v[0]): k for k, v in mesh.field_data.items()}
    tags = mesh.cell_data_dict["gmsh:physical"]["line"]
    names = list(props_map.keys())
    tag2sec = {}
    for t in np.unique(tags):
        name = gid2name.get(int(t))
        if not name or name not in props_map:
            raise KeyError(f"Section properties not defined for physical group '{name}'.")
        tag2sec[int(t)] = names.index(name)
    lut = np.full(int(max(tag2sec)) + 1, -1, dtype=np.int32)
    for t, s in tag2sec.items():
        lut[t] = s
    props = np.asarray([props_map[n] for n in names], dtype=np.float64).reshape(-1, 8)
    return lut[tags].astype(np.int32), props


# ----------------------------------------------------------------------------- Tet10

_KUHN_PERMS = [(0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0)]
_PERM_ODD = [False, True, True, False, False, True]
_EDGES = [(0, 1), (1, 2), (0, 2), (0, 3), (1, 3), (2, 3)]  # meshio/VTK tetra10 mid-edge order


def tet10_box_case(nx: int, ny: int, nz: int, dims=(0.8, 0.2, 0.8), force_data=None, fix_data=None):
    """Structured box of nx*ny*nz hexes, each split into 6 positively oriented Tet10
    (Kuhn), mid-edge nodes in meshio order (ReactionSolver.py:102-110).  Physical
    groups as gmsh_creation.py:67-71 writes them: ``box`` (volume), ``Neumann_BCs`` /
    ``Diri_BCs`` (points at the force / fix positions, snapped to the nearest corner
    node).  Defaults are the GUI's (FEM_main.py:115-127)."""
    if force_data is None:
        force_data = [{"force_x": 0.0, "force_y": 3000.0, "force_z": 0.0,
                       "force_x_pstn": 0.4, "force_y_pstn": 0.2, "force_z_pstn": 0.4}]
    if fix_data is None:
        fix_data = [{"pos_x": x, "pos_y": 0.0, "pos_z": z, "fix_x": 0, "fix_y": 0, "fix_z": 0}
                    for x in (0.0, 0.8) for z in (0.0, 0.8)]
    ci, cj, ck = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    origin = 2 * np.stack([ci, cj, ck], axis=-1).reshape(-1, 1, 1, 3)  # fine-grid coords
    tets = np.zeros((6, 4, 3), dtype=np.int64)
    for t, perm in enumerate(_KUHN_PERMS):
        v = np.zeros(3, dtype=np.int64)
        verts = [v.copy()]
        for ax in perm:
            v[ax] += 2
            verts.append(v.copy())
        if _PERM_ODD[t]:
            verts[2], verts[3] = verts[3], verts[2]
        tets[t] = np.asarray(verts)
    corners = origin + tets[None]                                  # (nh, 6, 4, 3)
    corners = corners.reshape(-1, 4, 3)
    mids = np.stack([(corners[:, a] + corners[:, b]) // 2 for a, b in _EDGES], axis=1)
    fine = np.concatenate([corners, mids], axis=1)                 # (ne, 10, 3)
    fy, fz = 2 * ny + 1, 2 * nz + 1
    key = (fine[..., 0] * fy + fine[..., 1]) * fz + fine[..., 2]
    uniq, inv = np.unique(key.ravel(), return_inverse=True)
    conn = inv.reshape(-1, 10)
    fx_ = uniq // (fy * fz)
    fy_ = (uniq // fz) % fy
    fz_ = uniq % fz
    pts = np.stack([fx_ * dims[0] / (2 * nx), fy_ * dims[1] / (2 * ny), fz_ * dims[2] / (2 * nz)], axis=1)
    is_corner = (fx_ % 2 == 0) & (fy_ % 2 == 0) & (fz_ % 2 == 0)
    corner_ids = np.nonzero(is_corner)[0]

    def snap(p):
        d = np.linalg.norm(pts[corner_ids] - np.asarray(p, dtype=float), axis=1)
        return int(corner_ids[np.argmin(d)])

    neu = [snap((f["force_x_pstn"], f["force_y_pstn"], f["force_z_pstn"])) for f in force_data]
    dir_ = [snap((f["pos_x"], f["pos_y"], f["pos_z"])) for f in fix_data]
    vcells = np.asarray(neu + dir_, dtype=np.int64).reshape(-1, 1)
    vtags = [2] * len(neu) + [3] * len(dir_)
    field = {"box": [1, 3], "Neumann_BCs": [2, 0], "Diri_BCs": [3, 0]}
    mesh = Mesh(pts, {"vertex": vcells, "tetra10": conn}, field,
                {"vertex": vtags, "tetra10": np.full(len(conn), 1)})
    return mesh, force_data, fix_data
