// Host-side symbolic analysis: block-CSR pattern over node adjacency, per-block contribution
// lists in a fixed (element-ascending) order — the "precomputed sorted scatter map" that
// makes the numeric assembly deterministic without float atomics — and the node tiles the
// assembly kernel walks.  Sparse replacement for the implicit pattern of the reference's
// dense scatter (BeamSolver.py:390-393) and of scipy's lil -> csr (ReactionSolver.py:148-151).
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <thread>
#include <vector>

#include "common.cuh"

namespace femb {

namespace {
struct Entry {
  int32_t col;
  uint32_t code;  // e*nper^2 + a*nper + b ; 0xFFFFFFFF = structural diagonal placeholder
};
inline bool entry_less(const Entry& x, const Entry& y) {
  return x.col != y.col ? x.col < y.col : x.code < y.code;
}
constexpr uint32_t kNoCode = 0xFFFFFFFFu;
}  // namespace

void build_symbolic(int64_t n_nodes, int64_t n_elem, int nper, int bs, const int32_t* conn,
                    int tile_max_blocks, int tile_max_contrib, Symbolic& S) {
  const bool trace = getenv("FEMB_TRACE") != nullptr;
  auto t_last = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!trace) return;
    auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[femb trace]   symbolic %s: %.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t_last).count());
    t_last = now;
  };
  S = Symbolic();
  S.bs = bs;
  S.nper = nper;
  S.n_nodes = n_nodes;
  S.n_elem = n_elem;
  const int64_t N = n_nodes;

  // 1. counting sort of (row, col, code) entries by row; every node also gets a placeholder
  //    diagonal entry so that block (i,i) exists even for points no element touches
  //    (the reference keeps their all-zero rows: BeamSolver.py:354,360).
  std::vector<int64_t> start(N + 1, 0);
  for (int64_t i = 0; i < N; ++i) start[i + 1] = 1;
  for (int64_t e = 0; e < n_elem; ++e)
    for (int a = 0; a < nper; ++a) start[conn[e * nper + a] + 1] += nper;
  for (int64_t i = 0; i < N; ++i) start[i + 1] += start[i];
  std::vector<Entry> ent(start[N]);
  std::vector<int64_t> cur(start.begin(), start.end() - 1);
  for (int64_t i = 0; i < N; ++i) ent[cur[i]++] = Entry{(int32_t)i, kNoCode};
  const uint32_t np2 = (uint32_t)(nper * nper);
  for (int64_t e = 0; e < n_elem; ++e) {
    for (int a = 0; a < nper; ++a) {
      const int32_t row = conn[e * nper + a];
      for (int b = 0; b < nper; ++b)
        ent[cur[row]++] = Entry{conn[e * nper + b], (uint32_t)e * np2 + (uint32_t)(a * nper + b)};
    }
  }

  lap("1 bucket entries by row");
  // 2. sort each row's entries by (col, code) — rows are independent, so split across threads
  unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  if (N < 4096) nt = 1;
  std::vector<int32_t> nblk_row(N, 0);
  auto work = [&](int64_t lo, int64_t hi) {
    for (int64_t i = lo; i < hi; ++i) {
      Entry* b = ent.data() + start[i];
      Entry* e = ent.data() + start[i + 1];
      std::sort(b, e, entry_less);
      int32_t nb = 0;
      int32_t prev = -1;
      for (Entry* p = b; p < e; ++p)
        if (p->col != prev) {
          ++nb;
          prev = p->col;
        }
      nblk_row[i] = nb;
    }
  };
  if (nt == 1) {
    work(0, N);
  } else {
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t) th.emplace_back(work, N * t / nt, N * (t + 1) / nt);
    for (auto& t : th) t.join();
  }

  lap("2 sort rows");
  // 3. block rows, columns, contribution lists
  S.rowptr.assign(N + 1, 0);
  for (int64_t i = 0; i < N; ++i) S.rowptr[i + 1] = S.rowptr[i] + nblk_row[i];
  S.nnzb = S.rowptr[N];
  S.colidx.resize(S.nnzb);
  S.blk_row.resize(S.nnzb);
  S.diag_blk.resize(N);
  S.contrib_ptr.assign(S.nnzb + 1, 0);
  S.n_contrib = (int64_t)ent.size() - N;
  S.contrib.resize(S.n_contrib);
  S.contrib_blk.resize(S.n_contrib);
  int64_t c = 0;
  for (int64_t i = 0; i < N; ++i) {
    int32_t blk = S.rowptr[i] - 1;
    int32_t prev = -1;
    for (int64_t k = start[i]; k < start[i + 1]; ++k) {
      const Entry& en = ent[k];
      if (en.col != prev) {
        ++blk;
        prev = en.col;
        S.colidx[blk] = en.col;
        S.blk_row[blk] = (int32_t)i;
        S.contrib_ptr[blk] = (int32_t)c;
        if (en.col == (int32_t)i) S.diag_blk[i] = blk;
      }
      if (en.code != kNoCode) {
        S.contrib[c] = en.code;
        S.contrib_blk[c] = blk;
        ++c;
      }
    }
  }
  S.contrib_ptr[S.nnzb] = (int32_t)c;

  lap("3 blocks + contribution lists");
  // 4. assembly tiles: greedy runs of consecutive block rows within the kernel's capacities
  S.tile_max_blocks = tile_max_blocks;
  S.tile_max_contrib = tile_max_contrib;
  S.tile_ptr.clear();
  S.tile_ptr.push_back(0);
  int64_t nb = 0, nc = 0;
  for (int64_t i = 0; i < N; ++i) {
    const int64_t rb = S.rowptr[i + 1] - S.rowptr[i];
    const int64_t rc = S.contrib_ptr[S.rowptr[i + 1]] - S.contrib_ptr[S.rowptr[i]];
    if (i > S.tile_ptr.back() && (nb + rb > tile_max_blocks || nc + rc > tile_max_contrib)) {
      S.tile_ptr.push_back((int32_t)i);
      nb = nc = 0;
    }
    nb += rb;
    nc += rc;
  }
  S.tile_ptr.push_back((int32_t)N);

  lap("4 tiles");
  // 4b. frame pair view (one thread of the pair kernel per (node, incident element end)):
  //     the diagonal block's list (codes e*4 + a*3, element-ascending) enumerates the pairs;
  //     each pair also owns the off-diagonal contribution (e, a, 1-a) to block (node, other).
  S.pairs_ok = false;
  if (nper == 2) {
    bool ok = true;
    S.pair_ptr.assign(N + 1, 0);
    for (int64_t i = 0; i < N; ++i) {
      const int32_t d = S.diag_blk[i];
      S.pair_ptr[i + 1] = S.pair_ptr[i] + (S.contrib_ptr[d + 1] - S.contrib_ptr[d]);
    }
    const int64_t np = S.pair_ptr[N];
    S.pair_code.resize(np); S.pair_blk.resize(np);
    for (int64_t i = 0; i < N && ok; ++i) {
      const int32_t d = S.diag_blk[i];
      int64_t p = S.pair_ptr[i];
      for (int32_t cc = S.contrib_ptr[d]; cc < S.contrib_ptr[d + 1] && ok; ++cc, ++p) {
        const uint32_t code = S.contrib[cc];
        const uint32_t e = code >> 2, a = (code >> 1) & 1u, b = code & 1u;
        if (a != b) { ok = false; break; }                       // self-loop element
        const int32_t other = conn[e * 2 + (1 - a)];
        if (other == (int32_t)i) { ok = false; break; }
        const int32_t* cb = S.colidx.data() + S.rowptr[i];
        const int32_t* ce = S.colidx.data() + S.rowptr[i + 1];
        const int32_t blk = (int32_t)(std::lower_bound(cb, ce, other) - S.colidx.data());
        const uint32_t ocode = e * 4u + a * 2u + (1u - a);
        const uint32_t* ob = S.contrib.data() + S.contrib_ptr[blk];
        const uint32_t* oe = S.contrib.data() + S.contrib_ptr[blk + 1];
        const int64_t rank = std::lower_bound(ob, oe, ocode) - ob;
        // duplicate members between the same two nodes (several contributions to one
        // off-diagonal block) and hubs with more than 127 members use the generic kernel
        if (rank != 0 || (oe - ob) != 1 || p - S.pair_ptr[i] > 127) { ok = false; break; }
        S.pair_code[p] = e * 2u + a;
        S.pair_blk[p] = blk;
      }
    }
    S.pairs_ok = ok;
    if (ok) {
      S.pair_tile_ptr.clear();
      S.pair_tile_ptr.push_back(0);
      int64_t tb = 0, tp = 0;
      for (int64_t i = 0; i < N; ++i) {
        const int64_t rp = S.pair_ptr[i + 1] - S.pair_ptr[i];
        // pair-kernel tiles: <= 128 pairs (one per thread) and <= 128 nodes
        if (i > S.pair_tile_ptr.back() && (tb + 1 > 128 || tp + rp > 128)) {
          S.pair_tile_ptr.push_back((int32_t)i);
          tb = tp = 0;
        }
        tb += 1;
        tp += rp;
      }
      S.pair_tile_ptr.push_back((int32_t)N);
    }
  }

  lap("4b pair view");
  // 5. chain detection (frames): every node has <= 2 distinct neighbours, no cycles.
  //    chain_order lists nodes path by path from an end point; isolated nodes last-in-place.
  S.is_chain = false;
  if (nper == 2 && N >= 1) {
    bool ok = true;
    for (int64_t i = 0; i < N && ok; ++i) ok = (nblk_row[i] - 1) <= 2;
    if (ok) {
      std::vector<char> seen(N, 0);
      S.chain_order.clear();
      S.chain_order.reserve(N);
      auto neighbours = [&](int32_t i, int32_t out[2]) {
        int n = 0;
        for (int32_t b = S.rowptr[i]; b < S.rowptr[i + 1]; ++b)
          if (S.colidx[b] != i) out[n++] = S.colidx[b];
        return n;
      };
      for (int64_t s = 0; s < N; ++s) {  // start walks from end points (degree <= 1)
        if (seen[s]) continue;
        int32_t nb2[2];
        if (neighbours((int32_t)s, nb2) > 1) continue;
        int32_t prev = -1, curn = (int32_t)s;
        while (curn >= 0 && !seen[curn]) {
          seen[curn] = 1;
          S.chain_order.push_back(curn);
          int32_t nn[2];
          const int k = neighbours(curn, nn);
          int32_t next = -1;
          for (int q = 0; q < k; ++q)
            if (nn[q] != prev && !seen[nn[q]]) next = nn[q];
          prev = curn;
          curn = next;
        }
      }
      S.is_chain = ((int64_t)S.chain_order.size() == N);  // false if a closed ring exists
      if (!S.is_chain) S.chain_order.clear();
    }
  }
  lap("5 chain detection");
}

}  // namespace femb

extern "C" int femb_symbolic_pattern(int64_t n_nodes, int64_t n_elem, int32_t nodes_per_elem,
                                     const int64_t* conn, int64_t* n_blocks, int32_t* rowptr,
                                     int32_t* colidx) {
  if (n_nodes < 0 || n_elem < 0 || nodes_per_elem < 1 || (!conn && n_elem > 0) || !n_blocks)
    return FEMB_ERR_ARG;
  std::vector<int32_t> c32((size_t)(n_elem * nodes_per_elem));
  for (size_t i = 0; i < c32.size(); ++i) {
    if (conn[i] < 0 || conn[i] >= n_nodes) return FEMB_ERR_ARG;
    c32[i] = (int32_t)conn[i];
  }
  femb::Symbolic S;
  femb::build_symbolic(n_nodes, n_elem, nodes_per_elem, nodes_per_elem == 2 ? 6 : 3, c32.data(),
                       1 << 30, 1 << 30, S);
  *n_blocks = S.nnzb;
  if (rowptr) std::memcpy(rowptr, S.rowptr.data(), S.rowptr.size() * sizeof(int32_t));
  if (colidx) std::memcpy(colidx, S.colidx.data(), S.colidx.size() * sizeof(int32_t));
  return FEMB_OK;
}
