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
