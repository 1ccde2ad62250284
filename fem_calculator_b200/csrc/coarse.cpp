// Host-side symbolic part of the two-level preconditioner (twolevel.cu): node aggregates by
// proportional recursive coordinate bisection, the aggregate adjacency induced by the block pattern,
// and for every stored block of K the slot of its coarse block inside its row aggregate's
// neighbour list.  Topology + one coordinate snapshot only; everything numeric (centroids, the
// Galerkin matrix, its inverse) is rebuilt on the device for every assembled K.
//
// The coarse space is the span of the six rigid-body modes of every aggregate (three translations,
// three rotations about the aggregate centroid) — the kernel of every unconstrained frame element
// (BeamSolver.py:646-660 has zero row sums against rigid motions), which is exactly the part of the
// spectrum a Jacobi-preconditioned CG on K_ff (BeamSolver.py:417) resolves slowest.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <vector>

#include "common.cuh"

namespace femb {

// agg[i] in [0, n_parts): every part gets floor/ceil(n/n_parts) nodes; parts are numbered in
// recursion order (left subtree first), so neighbouring ids are neighbouring boxes.
void build_aggregates(int64_t n_nodes, const double* xyz, int n_parts, std::vector<int32_t>& agg) {
  agg.assign((size_t)n_nodes, 0);
  if (n_parts <= 1 || n_nodes == 0) return;
  std::vector<int32_t> idx((size_t)n_nodes);
  std::iota(idx.begin(), idx.end(), 0);
  struct Job { int64_t lo, hi; int parts, base; };
  std::vector<Job> stack;
  stack.push_back({0, n_nodes, n_parts, 0});
  while (!stack.empty()) {
    const Job j = stack.back();
    stack.pop_back();
    if (j.parts == 1 || j.hi - j.lo <= 1) {
      for (int64_t k = j.lo; k < j.hi; ++k) agg[idx[k]] = j.base;
      continue;
    }
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
    for (int64_t k = j.lo; k < j.hi; ++k)
      for (int d = 0; d < 3; ++d) {
        const double v = xyz[3 * (size_t)idx[k] + d];
        mn[d] = std::min(mn[d], v);
        mx[d] = std::max(mx[d], v);
      }
    int ax = 0;
    for (int d = 1; d < 3; ++d)
      if (mx[d] - mn[d] > mx[ax] - mn[ax]) ax = d;
    const int pl = j.parts / 2;
    int64_t nl = ((j.hi - j.lo) * pl) / j.parts;
    nl = std::max<int64_t>(1, std::min<int64_t>(nl, j.hi - j.lo - 1));
    auto less = [&](int32_t a, int32_t b) {
      const double va = xyz[3 * (size_t)a + ax], vb = xyz[3 * (size_t)b + ax];
      return va != vb ? va < vb : a < b;
    };
    std::nth_element(idx.begin() + j.lo, idx.begin() + j.lo + nl, idx.begin() + j.hi, less);
    stack.push_back({j.lo + nl, j.hi, j.parts - pl, j.base + pl});
    stack.push_back({j.lo, j.lo + nl, pl, j.base});
  }
}

// CoarseSym from the block pattern: node lists (ascending node id), neighbour lists (ascending
// aggregate id, self included — every node has its diagonal block) and the per-block slot.
void build_coarse_symbolic(const Symbolic& S, const std::vector<int32_t>& agg, int n_agg, CoarseSym& C) {
  const int64_t N = S.n_nodes;
  C = CoarseSym();
  C.n_agg = n_agg;
  C.node_agg = agg;
  C.agg_ptr.assign((size_t)n_agg + 1, 0);
  for (int64_t i = 0; i < N; ++i) C.agg_ptr[agg[i] + 1]++;
  for (int a = 0; a < n_agg; ++a) C.agg_ptr[a + 1] += C.agg_ptr[a];
  C.agg_nodes.resize((size_t)N);
  {
    std::vector<int32_t> cur(C.agg_ptr.begin(), C.agg_ptr.end() - 1);
    for (int64_t i = 0; i < N; ++i) C.agg_nodes[cur[agg[i]]++] = (int32_t)i;
  }
  C.nbr_ptr.assign((size_t)n_agg + 1, 0);
  C.nbr.clear();
  std::vector<int32_t> tmp;
  for (int a = 0; a < n_agg; ++a) {
    tmp.clear();
    for (int32_t k = C.agg_ptr[a]; k < C.agg_ptr[a + 1]; ++k) {
      const int32_t i = C.agg_nodes[k];
      for (int32_t b = S.rowptr[i]; b < S.rowptr[i + 1]; ++b) tmp.push_back(agg[S.colidx[b]]);
    }
    std::sort(tmp.begin(), tmp.end());
    tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
    C.nbr.insert(C.nbr.end(), tmp.begin(), tmp.end());
    C.nbr_ptr[a + 1] = (int32_t)C.nbr.size();
    C.max_nbr = std::max<int32_t>(C.max_nbr, (int32_t)tmp.size());
  }
  C.blk_slot.resize((size_t)S.nnzb);
  for (int64_t i = 0; i < N; ++i) {
    const int a = agg[i];
    const int32_t* lo = C.nbr.data() + C.nbr_ptr[a];
    const int32_t* hi = C.nbr.data() + C.nbr_ptr[a + 1];
    for (int32_t b = S.rowptr[i]; b < S.rowptr[i + 1]; ++b)
      C.blk_slot[b] = (int32_t)(std::lower_bound(lo, hi, agg[S.colidx[b]]) - lo);
  }
}

// Member lines: maximal chains of (nearly) collinear frame members.  Groundwork for the next coarse
// space: a row of collinear members moving along its own axis is a near-null
// vector of D^-1 K_ff (it costs only the bending energy of the crossing members), and one axial
// translation mode per line takes those out of the Krylov iteration (tests/prototypes/coarse_space_study.py,
// tests/test_host_logic.py::test_member_line_modes_cut_the_iteration_count).
//   * at every node the incident member ends are paired greedily, most anti-parallel pair first, when
//     the cosine of their outward directions is below -cos_tol (ties broken by element id: deterministic);
//   * the pairings link elements into chains; every element belongs to exactly one chain; closed rings
//     are cut at their lowest element;
//   * chains with fewer than min_nodes nodes are dropped.
// Output: node lists ordered along the line, the unit end-to-end direction (for a closed ring: of its
// first member) and the "family" = index of the dominant direction component (the dense coarse solve
// is done per family).
void build_member_lines(int64_t n_nodes, int64_t n_elem, const int32_t* conn, const double* xyz, double cos_tol,
                        int min_nodes, std::vector<int32_t>& line_ptr, std::vector<int32_t>& line_nodes,
                        std::vector<double>& line_dir, std::vector<int32_t>& line_family) {
  line_ptr.assign(1, 0);
  line_nodes.clear();
  line_dir.clear();
  line_family.clear();
  // incident element ends per node
  std::vector<int32_t> ptr((size_t)n_nodes + 1, 0);
  for (int64_t e = 0; e < n_elem; ++e) { ptr[conn[2 * e] + 1]++; ptr[conn[2 * e + 1] + 1]++; }
  for (int64_t i = 0; i < n_nodes; ++i) ptr[i + 1] += ptr[i];
  std::vector<int32_t> inc((size_t)2 * n_elem);
  {
    std::vector<int32_t> cur(ptr.begin(), ptr.end() - 1);
    for (int64_t e = 0; e < n_elem; ++e)
      for (int a = 0; a < 2; ++a) inc[cur[conn[2 * e + a]]++] = (int32_t)(2 * e + a);   // code = 2 e + end
  }
  // link[2 e + end] = code of the member end that continues element e through that end's node, or -1
  std::vector<int32_t> link((size_t)2 * n_elem, -1);
  struct Cand { double dot; int32_t a, b; };
  std::vector<Cand> cand;
  std::vector<double> dir;
  for (int64_t v = 0; v < n_nodes; ++v) {
    const int k = ptr[v + 1] - ptr[v];
    if (k < 2) continue;
    dir.assign((size_t)3 * k, 0.0);
    for (int q = 0; q < k; ++q) {
      const int32_t code = inc[ptr[v] + q];
      const int32_t other = conn[2 * (code >> 1) + (1 - (code & 1))];
      double d[3], nrm = 0.0;
      for (int c = 0; c < 3; ++c) { d[c] = xyz[3 * (size_t)other + c] - xyz[3 * (size_t)v + c]; nrm += d[c] * d[c]; }
      nrm = nrm > 0.0 ? 1.0 / std::sqrt(nrm) : 0.0;
      for (int c = 0; c < 3; ++c) dir[3 * q + c] = d[c] * nrm;
    }
    cand.clear();
    for (int p = 0; p < k; ++p)
      for (int q = p + 1; q < k; ++q) {
        const double dot = dir[3 * p] * dir[3 * q] + dir[3 * p + 1] * dir[3 * q + 1] + dir[3 * p + 2] * dir[3 * q + 2];
        if (dot < -cos_tol) cand.push_back({dot, inc[ptr[v] + p], inc[ptr[v] + q]});
      }
    std::sort(cand.begin(), cand.end(), [](const Cand& x, const Cand& y) {
      return x.dot != y.dot ? x.dot < y.dot : (x.a != y.a ? x.a < y.a : x.b < y.b);
    });
    for (const Cand& c : cand)
      if (link[c.a] < 0 && link[c.b] < 0 && (c.a >> 1) != (c.b >> 1)) { link[c.a] = c.b; link[c.b] = c.a; }
  }
  // walk the chains: start at elements with a free end (lowest id first), then the closed rings
  std::vector<uint8_t> seen((size_t)n_elem, 0);
  std::vector<int32_t> nodes;
  auto walk = [&](int32_t e0, int start_end) {
    // leave e0 through the end opposite to start_end, i.e. start_end is the chain's first node
    nodes.clear();
    int32_t e = e0;
    int in_end = start_end;
    nodes.push_back(conn[2 * e + in_end]);
    while (true) {
      seen[e] = 1;
      const int out_end = 1 - in_end;
      nodes.push_back(conn[2 * e + out_end]);
      const int32_t nxt = link[2 * e + out_end];
      if (nxt < 0 || seen[nxt >> 1]) break;
      e = nxt >> 1;
      in_end = nxt & 1;
    }
    if ((int)nodes.size() < min_nodes) return;
    double d[3], nrm = 0.0;
    const int32_t a = nodes.front(), b = (nodes.back() != nodes.front()) ? nodes.back() : nodes[1];
    for (int c = 0; c < 3; ++c) { d[c] = xyz[3 * (size_t)b + c] - xyz[3 * (size_t)a + c]; nrm += d[c] * d[c]; }
    nrm = nrm > 0.0 ? 1.0 / std::sqrt(nrm) : 0.0;
    int fam = 0;
    for (int c = 0; c < 3; ++c) { d[c] *= nrm; if (std::fabs(d[c]) > std::fabs(d[fam])) fam = c; }
    if (d[fam] < 0.0) for (int c = 0; c < 3; ++c) d[c] = -d[c];      // orientation-independent sign
    line_nodes.insert(line_nodes.end(), nodes.begin(), nodes.end());
    line_ptr.push_back((int32_t)line_nodes.size());
    line_dir.insert(line_dir.end(), d, d + 3);
    line_family.push_back(fam);
  };
  for (int64_t e = 0; e < n_elem; ++e) {
    if (seen[e]) continue;
    if (link[2 * e] < 0) walk((int32_t)e, 0);
    else if (link[2 * e + 1] < 0) walk((int32_t)e, 1);
  }
  for (int64_t e = 0; e < n_elem; ++e)
    if (!seen[e]) walk((int32_t)e, 0);       // closed ring
}


// Symbolic phase of the line preconditioner (lines.cu).  Why lines: after diagonal scaling the slow modes of a
// frame are rows of collinear members moving along their own axis — such a motion costs only the bending
// energy of the crossing members (12 EI / L^3) while the diagonal carries EA / L.  The preconditioner
//     M^-1 = D^-1 + sum_lines Q_a (Q_a^T A Q_a)^-1 Q_a^T + sum_families P_f (P_f^T A P_f)^-1 P_f^T
// solves the axial chain of every member line exactly (a tridiagonal system per line) and couples the lines
// through one axial translation mode per BUNDLE of neighbouring lines (dense inverse per family, a few
// hundred unknowns).  Here: member lines (build_member_lines) cut to kLnMaxLen nodes, families = classes of
// node-disjoint lines (dominant direction component; a line that shares a node with an earlier line of its
// class moves to another class or is dropped), bundles = recursive coordinate bisection of the line
// midpoints into at most target_per_family groups per family.
void build_line_symbolic(const Symbolic& S, const int32_t* conn, const double* xyz, int target_per_family, LineSym& L) {
  L = LineSym();
  const int64_t N = S.n_nodes;
  std::vector<int32_t> lp, ln, lf;
  std::vector<double> ld;
  build_member_lines(N, S.n_elem, conn, xyz, 0.94, 3, lp, ln, ld, lf);
  // cut long lines; a closed ring repeats its first node at the end: drop the repeat
  struct Raw { int32_t lo, hi, fam; };
  std::vector<Raw> raw;
  for (size_t a = 0; a + 1 < lp.size(); ++a) {
    int32_t lo = lp[a], hi = lp[a + 1];
    if (hi - lo >= 2 && ln[lo] == ln[hi - 1]) --hi;
    const int32_t len = hi - lo;
    const int32_t pieces = (len + kLnMaxLen - 1) / kLnMaxLen;
    for (int32_t q = 0; q < pieces; ++q) {
      const int32_t a0 = lo + (int32_t)((int64_t)len * q / pieces), a1 = lo + (int32_t)((int64_t)len * (q + 1) / pieces);
      if (a1 - a0 >= 2) raw.push_back({a0, a1, lf[a]});
    }
  }
  // families: node-disjoint classes
  std::vector<uint8_t> used((size_t)kLnMaxFam * N, 0);
  std::vector<std::vector<int32_t>> fam_lines(kLnMaxFam);
  for (size_t a = 0; a < raw.size(); ++a) {
    int chosen = -1;
    for (int t = 0; t < kLnMaxFam && chosen < 0; ++t) {
      const int f = (raw[a].fam + t) % kLnMaxFam;
      bool free_ = true;
      for (int32_t k = raw[a].lo; k < raw[a].hi && free_; ++k) free_ = !used[(size_t)f * N + ln[k]];
      if (free_) chosen = f;
    }
    if (chosen < 0) continue;
    for (int32_t k = raw[a].lo; k < raw[a].hi; ++k) used[(size_t)chosen * N + ln[k]] = 1;
    fam_lines[chosen].push_back((int32_t)a);
  }
  // bundles per family: RCB of the line midpoints
  L.node_bundle.assign((size_t)kLnMaxFam * N, -1);
  L.node_ent.assign((size_t)kLnMaxFam * N, -1);
  L.node_line.assign((size_t)kLnMaxFam * N, -1);
  L.line_ptr.assign(1, 0);
  L.bundle_ptr.assign(1, 0);
  int32_t coarse0 = 0;
  for (int f = 0; f < kLnMaxFam; ++f) {
    const std::vector<int32_t>& lines = fam_lines[f];
    L.fam_off[f] = coarse0;
    const int nl = (int)lines.size();
    if (nl == 0) continue;
    const int nb = std::max(1, std::min(nl, target_per_family));
    std::vector<double> mid((size_t)nl * 3);
    for (int q = 0; q < nl; ++q) {
      const Raw& r = raw[lines[q]];
      for (int c = 0; c < 3; ++c) mid[3 * (size_t)q + c] = 0.5 * (xyz[3 * (size_t)ln[r.lo] + c] + xyz[3 * (size_t)ln[r.hi - 1] + c]);
    }
    std::vector<int32_t> part;
    build_aggregates(nl, mid.data(), nb, part);
    std::vector<int32_t> order((size_t)nl);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int32_t x, int32_t y) { return part[x] < part[y]; });
    int32_t prev = -1;
    for (int q = 0; q < nl; ++q) {
      const int32_t pq = part[order[q]];
      const Raw& r = raw[lines[order[q]]];
      if (pq != prev) {                               // RCB parts are never empty for nb <= nl
        if (prev >= 0) L.bundle_ptr.push_back((int32_t)L.line_bundle.size());
        prev = pq;
      }
      const int32_t cidx = coarse0 + pq;
      L.line_bundle.push_back(cidx);
      for (int32_t k = r.lo; k < r.hi; ++k) {
        const int32_t node = ln[k];
        const int32_t e = (int32_t)L.ent_node.size();
        L.ent_node.push_back(node);
        L.ent_blk_diag.push_back(S.diag_blk[node]);
        int32_t nxt = -1;
        if (k + 1 < r.hi) {
          const int32_t* b0 = S.colidx.data() + S.rowptr[node];
          const int32_t* b1 = S.colidx.data() + S.rowptr[node + 1];
          const int32_t* it = std::lower_bound(b0, b1, ln[k + 1]);
          if (it != b1 && *it == ln[k + 1]) nxt = (int32_t)(it - S.colidx.data());
        }
        L.ent_blk_next.push_back(nxt);
        L.node_bundle[(size_t)f * N + node] = cidx;
        L.node_ent[(size_t)f * N + node] = e;
        L.node_line[(size_t)f * N + node] = (int32_t)L.line_bundle.size() - 1;
      }
      L.line_ptr.push_back((int32_t)L.ent_node.size());
    }
    L.bundle_ptr.push_back((int32_t)L.line_bundle.size());
    coarse0 += nb;
  }
  L.fam_off[kLnMaxFam] = coarse0;
  L.n_coarse = coarse0;
  L.n_lines = (int32_t)L.line_bundle.size();
  L.n_entries = (int64_t)L.ent_node.size();
  int64_t covered = 0;
  for (int64_t i = 0; i < N; ++i) {
    bool any = false;
    for (int f = 0; f < kLnMaxFam; ++f) any = any || L.node_ent[(size_t)f * N + i] >= 0;
    covered += any;
  }
  L.coverage = N > 0 ? (double)covered / (double)N : 0.0;
}


void build_line_symbolic_local(const Symbolic& S, int64_t n_owned, int32_t n_coarse, const int32_t* fam_off,
                               const int32_t* node_bundle, const int32_t* node_line, const int32_t* node_pos,
                               const double* node_dir, LineSym& L) {
  L = LineSym();
  const int64_t N = S.n_nodes;
  L.n_coarse = n_coarse;
  for (int f = 0; f <= kLnMaxFam; ++f) L.fam_off[f] = fam_off[f];
  L.node_bundle.assign(node_bundle, node_bundle + (size_t)kLnMaxFam * N);
  L.node_dir.assign(node_dir, node_dir + (size_t)kLnMaxFam * N * 3);
  L.line_ptr.assign(1, 0);
  L.bundle_ptr.assign(1, 0);
  struct Item { int32_t bundle, line, pos, node; };
  std::vector<Item> items;
  for (int f = 0; f < kLnMaxFam; ++f) {
    items.clear();
    for (int64_t i = 0; i < n_owned; ++i) {
      const size_t fn = (size_t)f * N + i;
      if (node_bundle[fn] >= 0) items.push_back({node_bundle[fn], node_line[fn], node_pos[fn], (int32_t)i});
    }
    std::sort(items.begin(), items.end(), [](const Item& a, const Item& b) {
      return a.bundle != b.bundle ? a.bundle < b.bundle : (a.line != b.line ? a.line < b.line : a.pos < b.pos);
    });
    for (size_t k = 0; k < items.size();) {
      // one piece: consecutive positions of one global line, all owned
      size_t k1 = k + 1;
      while (k1 < items.size() && items[k1].line == items[k].line && items[k1].pos == items[k1 - 1].pos + 1 &&
             (int)(k1 - k) < kLnMaxLen) ++k1;
      const int32_t bundle = items[k].bundle;
      if (L.bundle_ids.empty() || L.bundle_ids.back() != bundle) {
        if (!L.bundle_ids.empty()) L.bundle_ptr.push_back((int32_t)L.line_bundle.size());
        L.bundle_ids.push_back(bundle);
      }
      L.line_bundle.push_back(bundle);
      for (size_t q = k; q < k1; ++q) {
        const int32_t node = items[q].node;
        L.ent_node.push_back(node);
        L.ent_blk_diag.push_back(S.diag_blk[node]);
        int32_t nxt = -1;
        if (q + 1 < k1) {
          const int32_t* b0 = S.colidx.data() + S.rowptr[node];
          const int32_t* b1 = S.colidx.data() + S.rowptr[node + 1];
          const int32_t* it = std::lower_bound(b0, b1, items[q + 1].node);
          if (it != b1 && *it == items[q + 1].node) nxt = (int32_t)(it - S.colidx.data());
        }
        L.ent_blk_next.push_back(nxt);
      }
      L.line_ptr.push_back((int32_t)L.ent_node.size());
      k = k1;
    }
  }
  if (!L.bundle_ids.empty()) L.bundle_ptr.push_back((int32_t)L.line_bundle.size());
  L.n_lines = (int32_t)L.line_bundle.size();
  L.n_entries = (int64_t)L.ent_node.size();
  L.coverage = 1.0;
}

}  // namespace femb

extern "C" int femb_symbolic_aggregates(int64_t n_nodes, const double* xyz, int32_t n_parts,
                                        int32_t* agg_of_node) {
  if (n_nodes < 0 || n_parts < 1 || (!xyz && n_nodes > 0) || (!agg_of_node && n_nodes > 0)) return FEMB_ERR_ARG;
  std::vector<int32_t> agg;
  femb::build_aggregates(n_nodes, xyz, n_parts, agg);
  if (n_nodes) std::memcpy(agg_of_node, agg.data(), (size_t)n_nodes * sizeof(int32_t));
  return FEMB_OK;
}

// Host-only view of the whole coarse symbolic phase for the CPU test-suite: aggregates, aggregate
// adjacency and per-block slots of a frame mesh (2 nodes per element).  Two calls: with nbr == NULL
// and blk_slot == NULL only agg_of_node, nbr_ptr and *n_nbr are filled.
extern "C" int femb_symbolic_coarse(int64_t n_nodes, int64_t n_elem, const int64_t* conn, const double* xyz,
                                    int32_t n_parts, int32_t* agg_of_node, int32_t* nbr_ptr, int64_t* n_nbr,
                                    int32_t* nbr, int32_t* blk_slot) {
  if (n_nodes < 0 || n_elem < 0 || n_parts < 1 || (!conn && n_elem > 0) || (!xyz && n_nodes > 0) || !n_nbr) return FEMB_ERR_ARG;
  std::vector<int32_t> c32((size_t)n_elem * 2);
  for (size_t i = 0; i < c32.size(); ++i) {
    if (conn[i] < 0 || conn[i] >= n_nodes) return FEMB_ERR_ARG;
    c32[i] = (int32_t)conn[i];
  }
  femb::Symbolic S;
  femb::build_symbolic(n_nodes, n_elem, 2, 6, c32.data(), 1 << 30, 1 << 30, S);
  std::vector<int32_t> agg;
  femb::build_aggregates(n_nodes, xyz, n_parts, agg);
  femb::CoarseSym C;
  femb::build_coarse_symbolic(S, agg, n_parts, C);
  if (agg_of_node && n_nodes) std::memcpy(agg_of_node, agg.data(), (size_t)n_nodes * sizeof(int32_t));
  if (nbr_ptr) std::memcpy(nbr_ptr, C.nbr_ptr.data(), C.nbr_ptr.size() * sizeof(int32_t));
  *n_nbr = (int64_t)C.nbr.size();
  if (nbr && !C.nbr.empty()) std::memcpy(nbr, C.nbr.data(), C.nbr.size() * sizeof(int32_t));
  if (blk_slot && !C.blk_slot.empty()) std::memcpy(blk_slot, C.blk_slot.data(), C.blk_slot.size() * sizeof(int32_t));
  return FEMB_OK;
}

// Host-only: member lines of a frame mesh (see build_member_lines).  Two calls: line_nodes == NULL
// returns the sizes in *n_lines / *n_line_nodes.
extern "C" int femb_symbolic_lines(int64_t n_nodes, int64_t n_elem, const int64_t* conn, const double* xyz,
                                   double cos_tol, int32_t min_nodes, int64_t* n_lines, int64_t* n_line_nodes,
                                   int32_t* line_ptr, int32_t* line_nodes, double* line_dir, int32_t* line_family) {
  if (n_nodes < 0 || n_elem < 0 || (!conn && n_elem > 0) || (!xyz && n_nodes > 0) || !n_lines || !n_line_nodes) return FEMB_ERR_ARG;
  if (!(cos_tol > 0.0 && cos_tol < 1.0) || min_nodes < 2) return FEMB_ERR_ARG;
  std::vector<int32_t> c32((size_t)n_elem * 2);
  for (size_t i = 0; i < c32.size(); ++i) {
    if (conn[i] < 0 || conn[i] >= n_nodes) return FEMB_ERR_ARG;
    c32[i] = (int32_t)conn[i];
  }
  std::vector<int32_t> lp, ln, lf;
  std::vector<double> ld;
  femb::build_member_lines(n_nodes, n_elem, c32.data(), xyz, cos_tol, min_nodes, lp, ln, ld, lf);
  *n_lines = (int64_t)lf.size();
  *n_line_nodes = (int64_t)ln.size();
  if (line_nodes) {
    if (!line_ptr || !line_dir || !line_family) return FEMB_ERR_ARG;
    std::memcpy(line_ptr, lp.data(), lp.size() * sizeof(int32_t));
    if (!ln.empty()) std::memcpy(line_nodes, ln.data(), ln.size() * sizeof(int32_t));
    if (!ld.empty()) std::memcpy(line_dir, ld.data(), ld.size() * sizeof(double));
    if (!lf.empty()) std::memcpy(line_family, lf.data(), lf.size() * sizeof(int32_t));
  }
  return FEMB_OK;
}

// Host-only view of the symbolic phase of the line preconditioner (csrc/lines.cu) for the CPU test-suite.
// node_bundle: (3, n_nodes) coarse index of the node's line in each family (-1: none); node_pos: (3, n_nodes)
// position of the node inside the sorted entry list (entries of one line are consecutive, lines of one bundle
// too); fam_off: (4) coarse index range per family.
extern "C" int femb_symbolic_line_bundles(int64_t n_nodes, int64_t n_elem, const int64_t* conn, const double* xyz,
                                          int32_t target_per_family, int32_t* node_bundle, int32_t* node_pos,
                                          int32_t* fam_off, int64_t* n_lines, int64_t* n_entries, double* coverage,
                                          int32_t* node_line, double* node_dir) {
  if (n_nodes < 0 || n_elem < 0 || (!conn && n_elem > 0) || (!xyz && n_nodes > 0) || target_per_family < 1) return FEMB_ERR_ARG;
  std::vector<int32_t> c32((size_t)n_elem * 2);
  for (size_t i = 0; i < c32.size(); ++i) {
    if (conn[i] < 0 || conn[i] >= n_nodes) return FEMB_ERR_ARG;
    c32[i] = (int32_t)conn[i];
  }
  femb::Symbolic S;
  femb::build_symbolic(n_nodes, n_elem, 2, 6, c32.data(), 1 << 30, 1 << 30, S);
  femb::LineSym L;
  femb::build_line_symbolic(S, c32.data(), xyz, target_per_family, L);
  if (node_bundle && !L.node_bundle.empty()) std::memcpy(node_bundle, L.node_bundle.data(), L.node_bundle.size() * sizeof(int32_t));
  if (node_pos && !L.node_ent.empty()) std::memcpy(node_pos, L.node_ent.data(), L.node_ent.size() * sizeof(int32_t));
  if (fam_off) std::memcpy(fam_off, L.fam_off, sizeof(L.fam_off));
  if (node_line && !L.node_line.empty()) std::memcpy(node_line, L.node_line.data(), L.node_line.size() * sizeof(int32_t));
  if (node_dir) {
    // unit end-to-end direction of every line, copied to its nodes (zero where the node has no line in the family)
    std::vector<double> ldir((size_t)L.n_lines * 3, 0.0);
    for (int32_t a = 0; a < L.n_lines; ++a) {
      const int32_t i0 = L.ent_node[L.line_ptr[a]], i1 = L.ent_node[L.line_ptr[a + 1] - 1];
      double d[3], n2 = 0.0;
      for (int c = 0; c < 3; ++c) { d[c] = xyz[3 * (size_t)i1 + c] - xyz[3 * (size_t)i0 + c]; n2 += d[c] * d[c]; }
      const double inv = n2 > 0.0 ? 1.0 / std::sqrt(n2) : 0.0;
      for (int c = 0; c < 3; ++c) ldir[3 * (size_t)a + c] = d[c] * inv;
    }
    for (int f = 0; f < femb::kLnMaxFam; ++f)
      for (int64_t i = 0; i < n_nodes; ++i) {
        const int32_t a = L.node_line[(size_t)f * n_nodes + i];
        for (int c = 0; c < 3; ++c) node_dir[((size_t)f * n_nodes + i) * 3 + c] = a >= 0 ? ldir[3 * (size_t)a + c] : 0.0;
      }
  }
  if (n_lines) *n_lines = L.n_lines;
  if (n_entries) *n_entries = L.n_entries;
  if (coverage) *coverage = L.coverage;
  return FEMB_OK;
}
