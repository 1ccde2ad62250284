"""Executable specification (numpy loops, CPU) of the line coarse space planned for round 2 (DESIGN.md
section 8), written against the DATA STRUCTURES the kernels will use rather than against sparse-matrix
algebra, so that the CUDA version is a transcription and this file is its parity oracle:

  lines        femb_symbolic_lines (csrc/coarse.cpp): line_ptr / line_nodes / line_dir / line_family
  node_lines   CSR node -> (line, position in line)               [prolongation gathers through it]
  segments     every line is cut into chunks of SEG consecutive nodes; seg_first[line] + pos // SEG
  K            block CSR (rowptr, colidx, 6x6 blocks) of the UN-eliminated matrix + the free-DOF mask

  galerkin()   per member direction f: Kf[a, b] = sum over blocks (i in a, j in b) of
               (m_i * t_a)^T K_ij[0:3, 0:3] (m_j * t_b)        one owner (line a) per row: no atomics
               and the segment diagonal Dseg[s] (same sum restricted to i, j in segment s)
  apply()      z = omega D^-1 r + sum_f P_f Kf^-1 P_f^T r + P_seg Dseg^-1 P_seg^T r
               restriction: one fixed-order sum per line / segment; prolongation: gather per node

tests/test_host_logic.py::test_line_precond_spec_matches_matrix_form checks it against P^T A P.
"""
import numpy as np

SEG = 8


def node_lines(n_nodes, line_ptr, line_nodes):
    cnt = np.bincount(line_nodes, minlength=n_nodes)
    ptr = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    lines = np.zeros(ptr[-1], dtype=np.int64)
    pos = np.zeros(ptr[-1], dtype=np.int64)
    cur = ptr[:-1].copy()
    for k in range(len(line_ptr) - 1):
        for p, v in enumerate(line_nodes[line_ptr[k]:line_ptr[k + 1]]):
            lines[cur[v]] = k
            pos[cur[v]] = p
            cur[v] += 1
    return ptr, lines, pos


def segments(line_ptr):
    length = np.diff(line_ptr)
    nseg = -(-length // SEG)
    seg_first = np.concatenate([[0], np.cumsum(nseg)]).astype(np.int64)
    return seg_first[:-1], int(seg_first[-1])


def galerkin(rowptr, colidx, blocks, mask, line_ptr, line_nodes, line_dir, line_family):
    """(list of three dense Kf, row index of every line inside its family, Dseg)"""
    n_nodes = len(rowptr) - 1
    nl_ptr, nl_line, nl_pos = node_lines(n_nodes, line_ptr, line_nodes)
    seg_first, n_seg = segments(line_ptr)
    fam_rows = [np.flatnonzero(line_family == f) for f in range(3)]
    row_in_family = np.zeros(len(line_family), dtype=np.int64)
    for f in range(3):
        row_in_family[fam_rows[f]] = np.arange(len(fam_rows[f]))
    Kf = [np.zeros((len(fam_rows[f]), len(fam_rows[f]))) for f in range(3)]
    Dseg = np.zeros(n_seg)
    m3 = mask.reshape(-1, 6)[:, :3].astype(float)
    for a in range(len(line_family)):                      # one owner per row of Kf: the line itself
        f = line_family[a]
        ta = line_dir[a]
        for pa, i in enumerate(line_nodes[line_ptr[a]:line_ptr[a + 1]]):
            ti = m3[i] * ta
            for blk in range(rowptr[i], rowptr[i + 1]):
                j = colidx[blk]
                w = ti @ blocks[blk][:3, :3]               # row vector (m_i t_a)^T K_ij[0:3, 0:3]
                for q in range(nl_ptr[j], nl_ptr[j + 1]):
                    b = nl_line[q]
                    if line_family[b] != f:
                        continue
                    v = w @ (m3[j] * line_dir[b])
                    Kf[f][row_in_family[a], row_in_family[b]] += v
                    if b == a and nl_pos[q] // SEG == pa // SEG:
                        Dseg[seg_first[a] + pa // SEG] += v
    for f in range(3):                                     # lines with no free DOF (e.g. inside a fixed base): identity
        dead = np.diag(Kf[f]) <= 0.0
        Kf[f][dead, dead] = 1.0
    return Kf, row_in_family, Dseg


def apply(r, dinv, omega, mask, line_ptr, line_nodes, line_dir, line_family, Kf_inv, row_in_family, Dseg):
    n_nodes = len(r) // 6
    seg_first, n_seg = segments(line_ptr)
    m3 = mask.reshape(-1, 6)[:, :3].astype(float)
    r3 = r.reshape(-1, 6)[:, :3]
    rl = np.zeros(len(line_family))
    rs = np.zeros(n_seg)
    for a in range(len(line_family)):                      # restriction: fixed order along the line
        for pa, i in enumerate(line_nodes[line_ptr[a]:line_ptr[a + 1]]):
            v = (m3[i] * line_dir[a]) @ r3[i]
            rl[a] += v
            rs[seg_first[a] + pa // SEG] += v
    yl = np.zeros_like(rl)
    for f in range(3):
        rows = np.flatnonzero(line_family == f)
        if len(rows):
            yl[rows[np.argsort(row_in_family[rows])]] = Kf_inv[f] @ rl[rows[np.argsort(row_in_family[rows])]]
    ys = np.where(Dseg > 0, rs / np.where(Dseg > 0, Dseg, 1.0), 0.0)
    z = omega * dinv * r
    z3 = z.reshape(-1, 6)[:, :3]
    nl_ptr, nl_line, nl_pos = node_lines(n_nodes, line_ptr, line_nodes)
    for i in range(n_nodes):                               # prolongation: gather per node
        for q in range(nl_ptr[i], nl_ptr[i + 1]):
            a = nl_line[q]
            z3[i] += m3[i] * line_dir[a] * (yl[a] + ys[seg_first[a] + nl_pos[q] // SEG])
    return z
