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
