"""Minimal Gmsh MSH 4.1 ASCII reader / writer producing the meshio-shaped record the
reference's solvers consume.

The reference reads meshes through ``meshio.read`` and then touches exactly four
attributes: ``points`` (N,3) float64 in file order, ``cells_dict[type]`` (0-based
point indices), ``field_data[name] = [physical_tag, dim]`` and
``cell_data_dict['gmsh:physical'][type]`` (BeamSolver.py:212-217,357-358,677-686;
ReactionSolver.py:62-85).  meshio is not installed here, so this module supplies the
same record for element types 15 (vertex), 1 (line) and 11 (tetra10).

meshio conventions reproduced:
  * node tags are renumbered to 0-based indices in order of appearance in $Nodes;
  * an entity's first physical tag becomes the 'gmsh:physical' value of its cells,
    0 when the entity has none;
  * gmsh's tetra10 node order differs from meshio/VTK in the last two mid-edge
    nodes: meshio swaps local nodes 8 and 9 on read (and back on write).
"""
from __future__ import annotations

import numpy as np

_GMSH_TYPES = {15: ("vertex", 1), 1: ("line", 2), 11: ("tetra10", 10)}
_TYPE_CODES = {"vertex": 15, "line": 1, "tetra10": 11}
_DIM_OF = {"vertex": 0, "line": 1, "tetra10": 3}


class Mesh:
    """In-memory mesh with the attribute layout of ``meshio.Mesh`` that the
    reference relies on."""

    def __init__(self, points, cells_dict, field_data=None, cell_physical=None):
        self.points = np.ascontiguousarray(points, dtype=np.float64)
        self.cells_dict = {k: np.ascontiguousarray(v, dtype=np.int64) for k, v in cells_dict.items()}
        self.field_data = {k: np.asarray(v, dtype=np.int64) for k, v in (field_data or {}).items()}
        phys = cell_physical or {}
        self.cell_data_dict = {"gmsh:physical": {k: np.asarray(v, dtype=np.int64) for k, v in phys.items()}}

    def group_nodes(self, group: str, cell_type: str = "vertex") -> np.ndarray:
        """Node indices of a physical group — BeamSolver.py:677-686 /
        ReactionSolver.py:75-85 semantics (sorted unique; empty if absent)."""
        cells = self.cells_dict.get(cell_type)
        phys = self.cell_data_dict["gmsh:physical"].get(cell_type)
        if cells is None or phys is None or group not in self.field_data:
            return np.zeros(0, dtype=np.int64)
        tag = int(self.field_data[group][0])
        return np.unique(cells[phys == tag].ravel())


def _sections(text: str) -> dict:
    out, name, buf = {}, None, []
    for raw in text.splitlines():
        line = raw.strip()
        if not line:
            continue
        if line.startswith("$End"):
            out[name] = buf
            name, buf = None, []
        elif line.startswith("$"):
            name, buf = line[1:], []
        elif name is not None:
            buf.append(line)
    return out


def read_msh(path: str) -> Mesh:
    with open(path, "r") as fh:
        sec = _sections(fh.read())
    ver = sec["MeshFormat"][0].split()
    if ver[0] != "4.1" or int(ver[1]) != 0:
        # 4.0 lays out $Entities / $Nodes / $Elements differently: refuse it instead of misreading tags and coordinates
        raise ValueError(f"only MSH 4.1 ASCII is supported, got header {ver}")

    field_data = {}
    for ln in sec.get("PhysicalNames", [])[1:]:
        dim, tag, name = ln.split(maxsplit=2)
        field_data[name.strip().strip('"')] = np.array([int(tag), int(dim)], dtype=np.int64)

    # entity (dim, tag) -> first physical tag
    ent_phys = {}
    ents = sec.get("Entities")
    if ents:
        npnt, ncur, nsur, nvol = (int(x) for x in ents[0].split())
        row = 1
        for dim, count in enumerate((npnt, ncur, nsur, nvol)):
            for _ in range(count):
                tok = ents[row].split()
                row += 1
                tag = int(tok[0])
                off = 4 if dim == 0 else 7  # xyz | bounding box
                nphys = int(tok[off])
                if nphys:
                    ent_phys[(dim, tag)] = int(tok[off + 1])

    nodes = sec["Nodes"]
    nblocks, nnodes = int(nodes[0].split()[0]), int(nodes[0].split()[1])
    points = np.zeros((nnodes, 3))
    tag2idx = {}
    row, k = 1, 0
    for _ in range(nblocks):
        _, _, parametric, nb = (int(x) for x in nodes[row].split())
        row += 1
        tags = [int(nodes[row + i]) for i in range(nb)]
        row += nb
        for i in range(nb):
            points[k] = [float(x) for x in nodes[row + i].split()[:3]]
            tag2idx[tags[i]] = k
            k += 1
        row += nb

    els = sec["Elements"]
    nblocks = int(els[0].split()[0])
    cells, phys = {}, {}
    row = 1
    for _ in range(nblocks):
        edim, etag, etype, nb = (int(x) for x in els[row].split())
        row += 1
        if etype not in _GMSH_TYPES:
            row += nb
            continue
        name, nn = _GMSH_TYPES[etype]
        for i in range(nb):
            tok = els[row + i].split()
            cells.setdefault(name, []).append([tag2idx[int(t)] for t in tok[1:1 + nn]])
            phys.setdefault(name, []).append(ent_phys.get((edim, etag), 0))
        row += nb
    cells_dict = {k: np.asarray(v, dtype=np.int64) for k, v in cells.items()}
    if "tetra10" in cells_dict:  # gmsh -> meshio/VTK order
        cells_dict["tetra10"] = cells_dict["tetra10"][:, [0, 1, 2, 3, 4, 5, 6, 7, 9, 8]]
    return Mesh(points, cells_dict, field_data, phys)


def write_msh(path: str, mesh: Mesh) -> None:
    """Write a Mesh as MSH 4.1 ASCII: one entity per (dim, physical tag), node tags =
    index+1 in one node block, so that ``read_msh(write_msh(m))`` reproduces
    ``points``/``cells_dict``/physical tags exactly (floats via repr)."""
    phys = mesh.cell_data_dict["gmsh:physical"]
    lines = ["$MeshFormat", "4.1 0 8", "$EndMeshFormat"]
    if mesh.field_data:
        lines.append("$PhysicalNames")
        lines.append(str(len(mesh.field_data)))
        for name, (tag, dim) in mesh.field_data.items():
            lines.append(f'{int(dim)} {int(tag)} "{name}"')
        lines.append("$EndPhysicalNames")
    ents = {0: [], 1: [], 2: [], 3: []}
    blocks = []
    for ctype, conn in mesh.cells_dict.items():
        dim = _DIM_OF[ctype]
        tags = phys.get(ctype, np.zeros(len(conn), dtype=np.int64))
        for ptag in sorted(set(int(t) for t in tags)):
            etag = len(ents[dim]) + 1
            ents[dim].append((etag, ptag))
            sel = conn[tags == ptag]
            if ctype == "tetra10":
                sel = sel[:, [0, 1, 2, 3, 4, 5, 6, 7, 9, 8]]
            blocks.append((dim, etag, _TYPE_CODES[ctype], sel))
    lines.append("$Entities")
    lines.append(" ".join(str(len(ents[d])) for d in range(4)))
    for dim in range(4):
        for etag, ptag in ents[dim]:
            box = "0 0 0" if dim == 0 else "0 0 0 0 0 0"
            ph = f"1 {ptag}" if ptag else "0"
            lines.append(f"{etag} {box} {ph}" + ("" if dim == 0 else " 0"))
    lines.append("$EndEntities")
    n = len(mesh.points)
    lines.append("$Nodes")
    lines.append(f"1 {n} 1 {n}")
    lines.append(f"3 1 0 {n}")
    lines.extend(str(i + 1) for i in range(n))
    lines.extend(" ".join(repr(float(c)) for c in p) for p in mesh.points)
    lines.append("$EndNodes")
    total = sum(len(b[3]) for b in blocks)
    lines.append("$Elements")
    lines.append(f"{len(blocks)} {total} 1 {total}")
    eid = 1
    for dim, etag, code, sel in blocks:
        lines.append(f"{dim} {etag} {code} {len(sel)}")
        for rowv in sel:
            lines.append(f"{eid} " + " ".join(str(int(v) + 1) for v in rowv))
            eid += 1
    lines.append("$EndElements")
    with open(path, "w") as fh:
        fh.write("\n".join(lines) + "\n")
