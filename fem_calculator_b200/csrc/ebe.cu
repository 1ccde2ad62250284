// Matrix-free ("element-by-element") form of the frame operator y = K_ff x for the Krylov loops.
//
// The assembled BSR SpMV (solver.cu) streams 8*36 bytes per block — 359 MB per product at 1M DOF,
// three times the L2 — and sits at ~75 % of the measured HBM bandwidth: it cannot get faster
// without moving fewer bytes.  A frame element's four 6x6 blocks, however, are a closed form of
// 19 numbers (the direction-cosine rows t, n1, n2 and ten stiffness magnitudes,
// BeamSolver.py:378-388, 646-660) that follow from two coordinate triples and a section row.  This
// kernel therefore never reads K: one thread per (node, incident element end) — the same 16-byte
// pair records the assembly kernel uses — rebuilds the element record in registers, projects the
// two end displacements on (t, n1, n2), applies the ten magnitudes and rotates the 6-vector back:
//     y_node += K_e[a][a] x_node + K_e[a][1-a] x_other            (BeamSolver.py:387, 390-393)
// ~75 FP64 FMAs + the record instead of 2 x 288 bytes of matrix.  The per-node sum over its pairs
// runs in list (element-ascending) order out of shared memory — no float atomics, bit-reproducible.
// Traffic per product: 16 B/pair + 16 B/node + tiles + xyz + x + y + mask = ~40 MB at 1M DOF, all of
// it L2-resident across iterations together with the CG vectors, so the update kernel speeds up too.
//
// The assembled K is still produced (Jacobi diagonal, reactions r = K u - f, CSR export); meshes the
// pair view cannot describe (duplicate members, hub nodes), Tet10 and the row-block distributed
// solver keep the BSR operator.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "ebe.cuh"
#include "elements.cuh"
#include "pcg_common.cuh"

namespace femb {

struct EbeDev {
  const int4* rec;        // (n_pairs) {node, other, blk, sec | a<<24 | pos<<25}
  const int4* node_rec;   // (n_nodes) {first pair, pair count, diagonal block, 0}
  const int4* tiles;      // (n_tiles) {first node, node count, first pair, pair count}
  int n_tiles;
};

// NB = 1: plain vectors x[g], y[g].  NB = 4: four vectors interleaved by right-hand side,
// x[g*4 + q] (the multi-RHS PCG layout of solver.cu).
// DOT: per-vector (x, y) through the ordered grid reduction into scal[0..NB); `done` (may be null)
// is the early-exit flag of the PCG that owns the launch.
template <int NB, int THREADS, bool MASKED, bool DOT, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
frame_ebe_kernel(const FrameParams P, const EbeDev pat, const uint8_t* __restrict__ free_mask,
                 const double* __restrict__ x, double* __restrict__ y, double* partials, int pstride,
                 double* scal, int* ticket, const int* done) {
  constexpr int W = 6 * NB;          // values per pair / per node
  constexpr int STRIDE = W + 1;      // odd stride in doubles: conflict-free phase-1 stores
  __shared__ double s_c[THREADS * STRIDE];
  __shared__ int2 s_node[THREADS];
  if (DOT && done && *done) return;
  const int tid = threadIdx.x;
  const int G = gridDim.x;
  const int4 zero4 = make_int4(0, 0, 0, 0);
  double dot = 0.0;                  // NB = 4: this thread only ever sees vector q = tid & 3
  int t = blockIdx.x;
  int4 td = (t < pat.n_tiles) ? __ldg(pat.tiles + t) : zero4;
  int4 td1 = (t + G < pat.n_tiles) ? __ldg(pat.tiles + t + G) : zero4;
  int4 rec = (tid < td.w) ? __ldg(pat.rec + td.z + tid) : zero4;
  FrameIn in;
  {
    const int a = (rec.w >> 24) & 1;
    frame_load(P, a ? rec.y : rec.x, a ? rec.x : rec.y, rec.w & 0xFFFFFF, in);   // inactive lanes: node 0 / row 0
  }
  for (; t < pat.n_tiles; t += G) {
    const int4 td2 = (t + 2 * G < pat.n_tiles) ? __ldg(pat.tiles + t + 2 * G) : zero4;
    const int4 rec1 = (tid < td1.w) ? __ldg(pat.rec + td1.z + tid) : zero4;
    const int4 nrec = (tid < td.y) ? __ldg(pat.node_rec + td.x + tid) : zero4;
    // ---- phase 1: one pair per thread -> its 6 (x NB) contributions to the node's rows
    if (tid < td.w) {
      const int a = (rec.w >> 24) & 1;
      KRec k;
      krec_from(P, in, a, k);
      double* mine = s_c + tid * STRIDE;
      if (NB == 1) {
        double ua[6], uo[6], o6[6];
        load6(x, rec.x, ua);
        load6(x, rec.y, uo);
        ebe_apply(k, ua, uo, o6);
#pragma unroll
        for (int c = 0; c < 6; ++c) mine[c] = o6[c];
      } else {
        const double2* pa = reinterpret_cast<const double2*>(x + (size_t)rec.x * W);
        const double2* po = reinterpret_cast<const double2*>(x + (size_t)rec.y * W);
#pragma unroll
        for (int half = 0; half < NB / 2; ++half) {   // two right-hand sides at a time (16-byte gathers)
          double ua0[6], ua1[6], uo0[6], uo1[6], o0[6], o1[6];
#pragma unroll
          for (int c = 0; c < 6; ++c) {
            const double2 va = __ldg(pa + c * (NB / 2) + half), vo = __ldg(po + c * (NB / 2) + half);
            ua0[c] = va.x; ua1[c] = va.y; uo0[c] = vo.x; uo1[c] = vo.y;
          }
          ebe_apply(k, ua0, uo0, o0);
          ebe_apply(k, ua1, uo1, o1);
#pragma unroll
          for (int c = 0; c < 6; ++c) { mine[c * NB + 2 * half] = o0[c]; mine[c * NB + 2 * half + 1] = o1[c]; }
        }
      }
    }
    s_node[tid] = make_int2(nrec.x - td.z, nrec.y);
    __syncthreads();
    // ---- the next tile's coordinates / section rows travel while phase 2 runs
    {
      const int a = (rec1.w >> 24) & 1;
      frame_load(P, a ? rec1.y : rec1.x, a ? rec1.x : rec1.y, rec1.w & 0xFFFFFF, in);
    }
    // ---- phase 2: ordered per-node sums, BC mask, store, (x, y) partial
    for (int o = tid; o < td.y * W; o += THREADS) {
      const int ns = o / W, q = o - ns * W;
      const int2 nr = s_node[ns];
      double v = 0.0;
      for (int j = nr.x; j < nr.x + nr.y; ++j) v += s_c[j * STRIDE + q];
      const size_t go = (size_t)td.x * W + o;
      double xg = 0.0;
      if (MASKED || DOT) xg = x[go];
      if (MASKED && !free_mask[NB == 1 ? go : go / NB]) v = xg;
      y[go] = v;
      if (DOT) dot += xg * v;
    }
    __syncthreads();   // s_c / s_node are rewritten by the next tile
    td = td1; td1 = td2; rec = rec1;
  }
  if (DOT) {
    double mine[NB], tot[NB];
#pragma unroll
    for (int q = 0; q < NB; ++q) mine[q] = (NB == 1 || (tid & (NB - 1)) == q) ? dot : 0.0;
    if (grid_reduce<THREADS, NB>(mine, partials, pstride, ticket, tot)) {
      if (tid == 0)
        for (int q = 0; q < NB; ++q) scal[q] = tot[q];
    }
  }
}

// ---- node-gather form: no tiles, no shared memory, no barriers -----------------------------
// LPN = (NBT / NB) * T lanes share one node: lane (qg, part) applies the pairs part, part+T, ... of
// the node's list to the NB right-hand sides [qg*NB, qg*NB+NB) and keeps the 6*NB sums in
// registers; the T partial sums are combined with a fixed xor-shuffle tree (deterministic), lane
// part 0 masks, stores and accumulates (x, y).  The node's own coordinates and x entries are
// loaded once per node instead of once per pair.
template <int NBT, int NB>
__device__ __forceinline__ void load_u(const double* __restrict__ x, int node, int qg, double (&u)[NB][6]) {
  if constexpr (NBT == 1) {
    load6(x, node, u[0]);
  } else if constexpr (NB == 1) {
#pragma unroll
    for (int c = 0; c < 6; ++c) u[0][c] = __ldg(x + ((size_t)node * 6 + c) * NBT + qg);
  } else if constexpr (NB == 2) {
    const double2* p = reinterpret_cast<const double2*>(x + (size_t)node * 6 * NBT + 2 * qg);
#pragma unroll
    for (int c = 0; c < 6; ++c) { const double2 v = __ldg(p + c * (NBT / 2)); u[0][c] = v.x; u[1][c] = v.y; }
  } else {
    static_assert(NB == 4 && NBT == 4, "vector grouping");
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      double4 v;
      asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w)
          : "l"(x + ((size_t)node * 6 + c) * 4));
      u[0][c] = v.x; u[1][c] = v.y; u[2][c] = v.z; u[3][c] = v.w;
    }
  }
}

template <int NBT, int NB>
__device__ __forceinline__ void store_u(double* __restrict__ y, int node, int qg, const double (&u)[NB][6]) {
  if constexpr (NBT == 1) {
    double2* p = reinterpret_cast<double2*>(y + (size_t)node * 6);
    p[0] = make_double2(u[0][0], u[0][1]); p[1] = make_double2(u[0][2], u[0][3]); p[2] = make_double2(u[0][4], u[0][5]);
  } else if constexpr (NB == 1) {
#pragma unroll
    for (int c = 0; c < 6; ++c) y[((size_t)node * 6 + c) * NBT + qg] = u[0][c];
  } else if constexpr (NB == 2) {
    double2* p = reinterpret_cast<double2*>(y + (size_t)node * 6 * NBT + 2 * qg);
#pragma unroll
    for (int c = 0; c < 6; ++c) p[c * (NBT / 2)] = make_double2(u[0][c], u[1][c]);
  } else {
#pragma unroll
    for (int c = 0; c < 6; ++c)
      *reinterpret_cast<double4*>(y + ((size_t)node * 6 + c) * 4) = make_double4(u[0][c], u[1][c], u[2][c], u[3][c]);
  }
}

template <int NBT, int NB, int T, bool MASKED, bool DOT, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
frame_ebe_node_kernel(const FrameParams P, const int4* __restrict__ pair_rec, const int4* __restrict__ node_rec,
                      int n_nodes, const uint8_t* __restrict__ free_mask, const double* __restrict__ x,
                      double* __restrict__ y, double* partials, int pstride, double* scal, int* ticket,
                      const int* done) {
  constexpr int QS = NBT / NB;
  constexpr int LPN = QS * T;
  constexpr int NPC = THREADS / LPN;   // nodes per CTA and grid-stride step
  static_assert(32 % LPN == 0 && THREADS % 32 == 0, "a node's lanes share a warp");
  pdl_wait();
  if (DOT && done && *done) return;
  const int lin = threadIdx.x % LPN;
  const int qg = lin / T, part = lin % T;
  double dot[NB];
#pragma unroll
  for (int q = 0; q < NB; ++q) dot[q] = 0.0;
  for (int base = blockIdx.x * NPC; base < n_nodes; base += gridDim.x * NPC) {
    const int node = base + threadIdx.x / LPN;
    const bool active = node < n_nodes;
    int first = 0, count = 0;
    double px = 0.0, py = 0.0, pz = 0.0;
    double ua[NB][6], acc[NB][6];
#pragma unroll
    for (int q = 0; q < NB; ++q)
#pragma unroll
      for (int c = 0; c < 6; ++c) { ua[q][c] = 0.0; acc[q][c] = 0.0; }
    if (active) {
      const int4 nr = __ldg(node_rec + node);
      first = nr.x; count = nr.y;
      const double* pp = P.xyz + 3 * (size_t)node;
      px = __ldg(pp); py = __ldg(pp + 1); pz = __ldg(pp + 2);
      load_u<NBT, NB>(x, node, qg, ua);
    }
#pragma unroll 2
    for (int j = part; j < count; j += T) {
      const int4 rec = __ldg(pair_rec + first + j);
      const int a = (rec.w >> 24) & 1;
      const double* po = P.xyz + 3 * (size_t)rec.y;
      const double ox = __ldg(po), oy = __ldg(po + 1), oz = __ldg(po + 2);
      const double* sp = P.sec_props + 8 * (size_t)(rec.w & 0xFFFFFF);
      FrameIn in;
      in.dx = a ? px - ox : ox - px; in.dy = a ? py - oy : oy - py; in.dz = a ? pz - oz : oz - pz;
      in.A = __ldg(sp); in.Ix = __ldg(sp + 1); in.Iy = __ldg(sp + 2); in.J = __ldg(sp + 3);
      in.ky = __ldg(sp + 4); in.kz = __ldg(sp + 5);
      double uo[NB][6];
      load_u<NBT, NB>(x, rec.y, qg, uo);
      KRec k;
      krec_from(P, in, a, k);
#pragma unroll
      for (int q = 0; q < NB; ++q) {
        double o6[6];
        ebe_apply(k, ua[q], uo[q], o6);
#pragma unroll
        for (int c = 0; c < 6; ++c) acc[q][c] += o6[c];
      }
    }
    if (T > 1) {
#pragma unroll
      for (int off = 1; off < T; off <<= 1)
#pragma unroll
        for (int q = 0; q < NB; ++q)
#pragma unroll
          for (int c = 0; c < 6; ++c) acc[q][c] += __shfl_xor_sync(0xffffffffu, acc[q][c], off);
    }
    if (active && part == 0) {
      if (MASKED) {
        const uint8_t* fm = free_mask + (size_t)node * 6;
#pragma unroll
        for (int c = 0; c < 6; ++c)
          if (!fm[c]) {
#pragma unroll
            for (int q = 0; q < NB; ++q) acc[q][c] = ua[q][c];
          }
      }
      store_u<NBT, NB>(y, node, qg, acc);
      if (DOT) {
#pragma unroll
        for (int q = 0; q < NB; ++q)
#pragma unroll
          for (int c = 0; c < 6; ++c) dot[q] += ua[q][c] * acc[q][c];
      }
    }
  }
  pdl_trigger();
  if (DOT) {
    double mine[NBT], tot[NBT];
#pragma unroll
    for (int q = 0; q < NBT; ++q) mine[q] = 0.0;
#pragma unroll
    for (int q = 0; q < NB; ++q) {
#pragma unroll
      for (int g = 0; g < QS; ++g)
        if (g == qg) mine[g * NB + q] = dot[q];
    }
    if (grid_reduce<THREADS, NBT>(mine, partials, pstride, ticket, tot)) {
      if (threadIdx.x == 0)
        for (int q = 0; q < NBT; ++q) scal[q] = tot[q];
    }
  }
}

// ---- node-gather form, software-pipelined ---------------------------------------------------
// Same mapping as frame_ebe_node_kernel, but (1) every CTA owns one contiguous, equally sized range
// of nodes (all SMs finish together; neighbouring nodes share gathers in L1), (2) the pair loop is
// a three-stage pipeline per lane — pair record two pairs ahead, the other end's coordinates /
// section row / x entries one pair ahead, arithmetic on the current pair — and (3) the next node's
// record is fetched while the current node is processed, so no dependent chain of global loads sits
// in front of the FP64 work.
template <int NBT, int NB>
struct PairOps {
  double ox, oy, oz;
  double uo[NB][6];
  int a, sec;        // the section row is read at use (a handful of rows: L1 hits)
};

template <int NBT, int NB>
__device__ __forceinline__ void load_pair_ops(const FrameParams& P, const double* __restrict__ x, const int4& rec, int qg,
                                              PairOps<NBT, NB>& o) {
  o.a = (rec.w >> 24) & 1;
  const double* po = P.xyz + 3 * (size_t)rec.y;
  o.ox = __ldg(po); o.oy = __ldg(po + 1); o.oz = __ldg(po + 2);
  o.sec = rec.w & 0xFFFFFF;
  load_u<NBT, NB>(x, rec.y, qg, o.uo);
}

template <int NBT, int NB, int T, bool MASKED, bool DOT, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
frame_ebe_node2_kernel(const FrameParams P, const int4* __restrict__ pair_rec, const int4* __restrict__ node_rec,
                       int n_nodes, const uint8_t* __restrict__ free_mask, const double* __restrict__ x,
                       double* __restrict__ y, double* partials, int pstride, double* scal, int* ticket,
                       const int* done) {
  constexpr int QS = NBT / NB;
  constexpr int LPN = QS * T;
  constexpr int NPC = THREADS / LPN;
  static_assert(32 % LPN == 0 && THREADS % 32 == 0, "a node's lanes share a warp");
  if (DOT && done && *done) return;
  const int lin = threadIdx.x % LPN;
  const int qg = lin / T, part = lin % T;
  const int4 zero4 = make_int4(0, 0, 0, 0);
  double dot[NB];
#pragma unroll
  for (int q = 0; q < NB; ++q) dot[q] = 0.0;
  const int per = (n_nodes + (int)gridDim.x - 1) / (int)gridDim.x;
  const int lo = blockIdx.x * per;
  const int hi = min(n_nodes, lo + per);
  int node = lo + threadIdx.x / LPN;
  int4 nr = (node < hi) ? __ldg(node_rec + node) : zero4;
  for (int base = lo; base < hi; base += NPC, node += NPC) {
    const bool active = node < hi;
    const int first = nr.x, count = active ? nr.y : 0;
    int j = part;
    int4 rc = (j < count) ? __ldg(pair_rec + first + j) : zero4;
    int4 rn = (j + T < count) ? __ldg(pair_rec + first + j + T) : zero4;
    nr = (node + NPC < hi) ? __ldg(node_rec + node + NPC) : zero4;      // next node of this lane
    double px = 0.0, py = 0.0, pz = 0.0;
    double ua[NB][6], acc[NB][6];
#pragma unroll
    for (int q = 0; q < NB; ++q)
#pragma unroll
      for (int c = 0; c < 6; ++c) { ua[q][c] = 0.0; acc[q][c] = 0.0; }
    if (active) {
      const double* pp = P.xyz + 3 * (size_t)node;
      px = __ldg(pp); py = __ldg(pp + 1); pz = __ldg(pp + 2);
      load_u<NBT, NB>(x, node, qg, ua);
    }
    PairOps<NBT, NB> cur;
    load_pair_ops<NBT, NB>(P, x, rc, qg, cur);
#pragma unroll 1
    for (; j < count; j += T) {
      PairOps<NBT, NB> nxt;
      load_pair_ops<NBT, NB>(P, x, rn, qg, nxt);          // past the end: rn = 0 -> node 0 / row 0 (valid, unused)
      rn = (j + 2 * T < count) ? __ldg(pair_rec + first + j + 2 * T) : zero4;
      FrameIn in;
      in.dx = cur.a ? px - cur.ox : cur.ox - px;
      in.dy = cur.a ? py - cur.oy : cur.oy - py;
      in.dz = cur.a ? pz - cur.oz : cur.oz - pz;
      const double* sp = P.sec_props + 8 * (size_t)cur.sec;
      in.A = __ldg(sp); in.Ix = __ldg(sp + 1); in.Iy = __ldg(sp + 2); in.J = __ldg(sp + 3);
      in.ky = __ldg(sp + 4); in.kz = __ldg(sp + 5);
      KRec k;
      krec_from(P, in, cur.a, k);
#pragma unroll
      for (int q = 0; q < NB; ++q) {
        double o6[6];
        ebe_apply(k, ua[q], cur.uo[q], o6);
#pragma unroll
        for (int c = 0; c < 6; ++c) acc[q][c] += o6[c];
      }
      cur = nxt;
    }
    if (T > 1) {
#pragma unroll
      for (int off = 1; off < T; off <<= 1)
#pragma unroll
        for (int q = 0; q < NB; ++q)
#pragma unroll
          for (int c = 0; c < 6; ++c) acc[q][c] += __shfl_xor_sync(0xffffffffu, acc[q][c], off);
    }
    if (active && part == 0) {
      if (MASKED) {
        const uint8_t* fm = free_mask + (size_t)node * 6;
#pragma unroll
        for (int c = 0; c < 6; ++c)
          if (!fm[c]) {
#pragma unroll
            for (int q = 0; q < NB; ++q) acc[q][c] = ua[q][c];
          }
      }
      store_u<NBT, NB>(y, node, qg, acc);
      if (DOT) {
#pragma unroll
        for (int q = 0; q < NB; ++q)
#pragma unroll
          for (int c = 0; c < 6; ++c) dot[q] += ua[q][c] * acc[q][c];
      }
    }
  }
  if (DOT) {
    double mine[NBT], tot[NBT];
#pragma unroll
    for (int q = 0; q < NBT; ++q) mine[q] = 0.0;
#pragma unroll
    for (int q = 0; q < NB; ++q) {
#pragma unroll
      for (int g = 0; g < QS; ++g)
        if (g == qg) mine[g * NB + q] = dot[q];
    }
    if (grid_reduce<THREADS, NBT>(mine, partials, pstride, ticket, tot)) {
      if (threadIdx.x == 0)
        for (int q = 0; q < NBT; ++q) scal[q] = tot[q];
    }
  }
}

// ---- stored element records ------------------------------------------------------------------
// Middle ground between reading K (720 B per element and product) and rebuilding the element
// record per pair (~100 FP64 instructions, two rsqrt, two divisions): after each assembly one
// small kernel stores, per element, the 15 numbers R^T k R is made of — t (3), the two non-zero
// components of n1 (n1.z = 0 in both branches of BeamSolver.py:380-384; n2 = t x n1) and the ten
// stiffness magnitudes — in a 128-byte record.  The operator then reads 128 B per element instead of
// 720 B (5.6x fewer bytes than the BSR SpMV) and is HBM-bound again, but at 64 MB per product.
__global__ void ebe_build_records_kernel(const FrameParams P, int64_t n_elem, double* __restrict__ rec) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_elem) return;
  FrameRec R;
  frame_record(P, (uint32_t)e, R);
  double4* o = reinterpret_cast<double4*>(rec + (size_t)e * 16);
  o[0] = make_double4(R.t[0], R.t[1], R.t[2], R.n1[0]);
  o[1] = make_double4(R.n1[1], R.ax, R.tor, R.k11z);
  o[2] = make_double4(R.k11y, R.k12z, R.k12y, R.k23z);
  o[3] = make_double4(R.k23y, R.k22z - R.k23z, R.k22y - R.k23y, 0.0);
}

__device__ __forceinline__ double4 ldg256(const double* p) {
  double4 v;
  asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
  return v;
}

struct KRecS {   // stored form, n1 = (n1x, n1y, 0)
  double tx, ty, tz, n1x, n1y, n2x, n2y, n2z;
  double ax, tor, k11z, k11y, c12z, c12y, k23z, k23y, d22z, d22y;
};

__device__ __forceinline__ void krec_load(const double* __restrict__ rec, int e, int a, KRecS& k) {
  const double* r = rec + (size_t)e * 16;
  const double4 v0 = ldg256(r), v1 = ldg256(r + 4), v2 = ldg256(r + 8), v3 = ldg256(r + 12);
  const double sa = a ? -1.0 : 1.0;
  k.tx = v0.x; k.ty = v0.y; k.tz = v0.z; k.n1x = v0.w; k.n1y = v1.x;
  k.n2x = -k.tz * k.n1y; k.n2y = k.tz * k.n1x; k.n2z = k.tx * k.n1y - k.ty * k.n1x;   // n2 = t x n1
  k.ax = v1.y; k.tor = v1.z; k.k11z = v1.w; k.k11y = v2.x;
  k.c12z = sa * v2.y; k.c12y = sa * v2.z; k.k23z = v2.w; k.k23y = v3.x; k.d22z = v3.y; k.d22y = v3.z;
}

// same closed form as ebe_apply with n1.z = 0 folded in
__device__ __forceinline__ void ebe_apply_s(const KRecS& k, const double* ua, const double* uo, double* acc) {
  const double ddx = ua[0] - uo[0], ddy = ua[1] - uo[1], ddz = ua[2] - uo[2];
  const double tdx = ua[3] - uo[3], tdy = ua[4] - uo[4], tdz = ua[5] - uo[5];
  const double tsx = ua[3] + uo[3], tsy = ua[4] + uo[4], tsz = ua[5] + uo[5];
  const double dt = k.tx * ddx + k.ty * ddy + k.tz * ddz;
  const double d1 = k.n1x * ddx + k.n1y * ddy;
  const double d2 = k.n2x * ddx + k.n2y * ddy + k.n2z * ddz;
  const double tt = k.tx * tdx + k.ty * tdy + k.tz * tdz;
  const double s1 = k.n1x * tsx + k.n1y * tsy;
  const double s2 = k.n2x * tsx + k.n2y * tsy + k.n2z * tsz;
  const double a1 = k.n1x * ua[3] + k.n1y * ua[4];
  const double a2 = k.n2x * ua[3] + k.n2y * ua[4] + k.n2z * ua[5];
  const double ft = k.ax * dt;
  const double f1 = k.k11z * d1 + k.c12z * s2;
  const double f2 = k.k11y * d2 - k.c12y * s1;
  const double mt = k.tor * tt;
  const double m1 = k.k23y * s1 + k.d22y * a1 - k.c12y * d2;
  const double m2 = k.k23z * s2 + k.d22z * a2 + k.c12z * d1;
  acc[0] += ft * k.tx + f1 * k.n1x + f2 * k.n2x;
  acc[1] += ft * k.ty + f1 * k.n1y + f2 * k.n2y;
  acc[2] += ft * k.tz + f2 * k.n2z;
  acc[3] += mt * k.tx + m1 * k.n1x + m2 * k.n2x;
  acc[4] += mt * k.ty + m1 * k.n1y + m2 * k.n2y;
  acc[5] += mt * k.tz + m2 * k.n2z;
}

template <int NBT, int NB, int T, bool MASKED, bool DOT, int THREADS, int MINB, int UNR>
__global__ void __launch_bounds__(THREADS, MINB)
frame_ebe_rec_kernel(const double* __restrict__ rec, const int2* __restrict__ pairs, const int4* __restrict__ node_rec,
                     int n_nodes, const uint8_t* __restrict__ free_mask, const double* __restrict__ x,
                     double* __restrict__ y, double* partials, int pstride, double* scal, int* ticket,
                     const int* done) {
  constexpr int QS = NBT / NB;
  constexpr int LPN = QS * T;
  constexpr int NPC = THREADS / LPN;
  static_assert(32 % LPN == 0 && THREADS % 32 == 0, "a node's lanes share a warp");
  if (DOT && done && *done) return;
  const int lin = threadIdx.x % LPN;
  const int qg = lin / T, part = lin % T;
  const int4 zero4 = make_int4(0, 0, 0, 0);
  double dot[NB];
#pragma unroll
  for (int q = 0; q < NB; ++q) dot[q] = 0.0;
  const int per = (n_nodes + (int)gridDim.x - 1) / (int)gridDim.x;   // contiguous, equally sized node ranges
  const int lo = blockIdx.x * per;
  const int hi = min(n_nodes, lo + per);
  int node = lo + threadIdx.x / LPN;
  int4 nr = (node < hi) ? __ldg(node_rec + node) : zero4;
  for (int base = lo; base < hi; base += NPC, node += NPC) {
    const bool active = node < hi;
    const int first = nr.x, count = active ? nr.y : 0;
    nr = (node + NPC < hi) ? __ldg(node_rec + node + NPC) : zero4;    // next node of this lane
    double ua[NB][6], acc[NB][6];
#pragma unroll
    for (int q = 0; q < NB; ++q)
#pragma unroll
      for (int c = 0; c < 6; ++c) { ua[q][c] = 0.0; acc[q][c] = 0.0; }
    if (active) load_u<NBT, NB>(x, node, qg, ua);
#pragma unroll UNR
    for (int j = part; j < count; j += T) {
      const int2 pr = __ldg(pairs + first + j);
      KRecS k;
      krec_load(rec, pr.y >> 1, pr.y & 1, k);
      double uo[NB][6];
      load_u<NBT, NB>(x, pr.x, qg, uo);
#pragma unroll
      for (int q = 0; q < NB; ++q) ebe_apply_s(k, ua[q], uo[q], acc[q]);
    }
    if (T > 1) {
#pragma unroll
      for (int off = 1; off < T; off <<= 1)
#pragma unroll
        for (int q = 0; q < NB; ++q)
#pragma unroll
          for (int c = 0; c < 6; ++c) acc[q][c] += __shfl_xor_sync(0xffffffffu, acc[q][c], off);
    }
    if (active && part == 0) {
      if (MASKED) {
        const uint8_t* fm = free_mask + (size_t)node * 6;
#pragma unroll
        for (int c = 0; c < 6; ++c)
          if (!fm[c]) {
#pragma unroll
            for (int q = 0; q < NB; ++q) acc[q][c] = ua[q][c];
          }
      }
      store_u<NBT, NB>(y, node, qg, acc);
      if (DOT) {
#pragma unroll
        for (int q = 0; q < NB; ++q)
#pragma unroll
          for (int c = 0; c < 6; ++c) dot[q] += ua[q][c] * acc[q][c];
      }
    }
  }
  if (DOT) {
    double mine[NBT], tot[NBT];
#pragma unroll
    for (int q = 0; q < NBT; ++q) mine[q] = 0.0;
#pragma unroll
    for (int q = 0; q < NB; ++q) {
#pragma unroll
      for (int g = 0; g < QS; ++g)
        if (g == qg) mine[g * NB + q] = dot[q];
    }
    if (grid_reduce<THREADS, NBT>(mine, partials, pstride, ticket, tot)) {
      if (threadIdx.x == 0)
        for (int q = 0; q < NBT; ++q) scal[q] = tot[q];
    }
  }
}

static FrameParams ebe_params(const femb_handle* h) {
  FrameParams P;
  P.xyz = h->xyz.p; P.conn = h->conn.p; P.elem_sec = h->elem_sec.p; P.sec_props = h->sec_props.p;
  P.E = h->E; P.G = h->G; P.rho = h->rho;
  return P;
}

// operator choice of a Krylov solve: opts.reserved (FEMB_OP_*), overridable with FEMB_OPERATOR=bsr|ebe
bool ebe_available(const femb_handle* h) {
  return h->kind == Kind::Frame && h->pairs_dev_ok && !dist_active(h);
}

bool ebe_selected(const femb_handle* h, int op) {
  static int env = -1;
  if (env < 0) {
    const char* e = getenv("FEMB_OPERATOR");
    env = 0;
    if (e && (e[0] == 'b' || e[0] == 'B')) env = FEMB_OP_BSR;
    if (e && (e[0] == 'e' || e[0] == 'E')) env = FEMB_OP_EBE;
  }
  if (env && op != FEMB_OP_EBE_FUSED) op = env;
  if (op == FEMB_OP_BSR) return false;
  return ebe_available(h);
}

constexpr int kEbeThreads = 128;

double ebe_bytes(const femb_handle* h, int nb) {
  const Symbolic& S = h->sym;
  const double n_tiles = (double)S.pair_tile_ptr.size() - 1.0;
  // pair + node + tile records, coordinates, x read once, y written, BC mask
  return 16.0 * (double)S.pair_code.size() + 16.0 * h->n_nodes + 16.0 * n_tiles + 24.0 * h->n_nodes +
         16.0 * nb * h->ndof + 1.0 * h->ndof;
}

template <int NBT, int NB, int T, int THREADS, int MINB>
static void launch_node_variant(femb_handle* h, const FrameParams& P, const double* x, double* y, bool masked,
                                double* dot_partials, double* scal_out, int* ticket, const int* done, int per_sm) {
  constexpr int NPC = THREADS / ((NBT / NB) * T);
  const int n_nodes = (int)h->n_nodes;
  const int need = (n_nodes + NPC - 1) / NPC;
  const int grid = std::max(1, std::min(need, h->num_sms * std::min(per_sm > 0 ? per_sm : MINB, 8)));
  const int pstride = h->num_sms * 8;
  const int4* pr = reinterpret_cast<const int4*>(h->pair_rec.p);
  const int4* nr = reinterpret_cast<const int4*>(h->pair_node_rec.p);
#define EBEN(M, D) launch_pdl(frame_ebe_node_kernel<NBT, NB, T, M, D, THREADS, MINB>, grid, THREADS, h->stream, \
    P, pr, nr, n_nodes, (const uint8_t*)h->free_mask.p, x, y, dot_partials, pstride, scal_out, ticket, done)
  if (masked && dot_partials) EBEN(true, true);
  else if (masked) EBEN(true, false);
  else EBEN(false, false);
#undef EBEN
}

template <int NBT, int NB, int T, int THREADS, int MINB>
static void launch_node2_variant(femb_handle* h, const FrameParams& P, const double* x, double* y, bool masked,
                                 double* dot_partials, double* scal_out, int* ticket, const int* done, int per_sm) {
  constexpr int NPC = THREADS / ((NBT / NB) * T);
  const int n_nodes = (int)h->n_nodes;
  const int need = (n_nodes + NPC - 1) / NPC;
  const int grid = std::max(1, std::min(need, h->num_sms * std::min(per_sm > 0 ? per_sm : MINB, 8)));
  const int pstride = h->num_sms * 8;
  const int4* pr = reinterpret_cast<const int4*>(h->pair_rec.p);
  const int4* nr = reinterpret_cast<const int4*>(h->pair_node_rec.p);
#define EBEN(M, D) frame_ebe_node2_kernel<NBT, NB, T, M, D, THREADS, MINB><<<grid, THREADS, 0, h->stream>>>( \
    P, pr, nr, n_nodes, h->free_mask.p, x, y, dot_partials, pstride, scal_out, ticket, done)
  if (masked && dot_partials) EBEN(true, true);
  else if (masked) EBEN(true, false);
  else EBEN(false, false);
#undef EBEN
}

static int ensure_ebe_records(femb_handle* h) {
  if (h->ebe_rec_valid) return FEMB_OK;
  FEMB_CUDA(h, h->ebe_rec.ensure((size_t)h->n_elem * 16));
  ebe_build_records_kernel<<<(unsigned)((h->n_elem + 127) / 128), 128, 0, h->stream>>>(ebe_params(h), h->n_elem, h->ebe_rec.p);
  h->launches++;
  FEMB_CUDA(h, cudaGetLastError());
  h->ebe_rec_valid = true;
  return FEMB_OK;
}

template <int NBT, int NB, int T, int THREADS, int MINB, int UNR>
static void launch_rec_variant(femb_handle* h, const double* x, double* y, bool masked, double* dot_partials,
                               double* scal_out, int* ticket, const int* done, int per_sm) {
  constexpr int NPC = THREADS / ((NBT / NB) * T);
  const int n_nodes = (int)h->n_nodes;
  const int need = (n_nodes + NPC - 1) / NPC;
  const int grid = std::max(1, std::min(need, h->num_sms * std::min(per_sm > 0 ? per_sm : MINB, 8)));
  const int pstride = h->num_sms * 8;
  const int2* pr = reinterpret_cast<const int2*>(h->ebe_pair.p);
  const int4* nr = reinterpret_cast<const int4*>(h->pair_node_rec.p);
#define EBER(M, D) frame_ebe_rec_kernel<NBT, NB, T, M, D, THREADS, MINB, UNR><<<grid, THREADS, 0, h->stream>>>( \
    h->ebe_rec.p, pr, nr, n_nodes, h->free_mask.p, x, y, dot_partials, pstride, scal_out, ticket, done)
  if (masked && dot_partials) EBER(true, true);
  else if (masked) EBER(true, false);
  else EBER(false, false);
#undef EBER
}

// y = K_ff x (masked) or K x; nb = 1 (plain) or 4 (interleaved).  dot_partials != null: (x_q, y_q)
// -> scal_out[q], ticket = reduction ticket slot, done = early-exit flag (may be null).
// FEMB_EBE_VARIANT: 0 = tile kernel (pairs in shared memory); node-gather kernels: nb=1: 1 -> T=1,
// 2 -> T=2, 3 -> T=4;  nb=4: 1 -> 4 vectors/lane, 2 -> 2 vectors/lane, 3 -> 2 vectors/lane T=2,
// 4 -> 1 vector/lane.
int launch_ebe(femb_handle* h, const double* x, double* y, int nb, bool masked, double* dot_partials,
               double* scal_out, int* ticket, const int* done) {
  const Symbolic& S = h->sym;
  const int n_tiles = (int)S.pair_tile_ptr.size() - 1;
  if (n_tiles <= 0) return FEMB_OK;
  EbeDev pp{reinterpret_cast<const int4*>(h->pair_rec.p), reinterpret_cast<const int4*>(h->pair_node_rec.p),
            reinterpret_cast<const int4*>(h->pair_tiles.p), n_tiles};
  const FrameParams P = ebe_params(h);
  const int pstride = h->num_sms * 8;
  static int ctas = -1, variant1 = -1, variant4 = -1;
  if (ctas < 0) {
    const char* e = getenv("FEMB_EBE_CTAS"); ctas = e ? atoi(e) : 0;
    e = getenv("FEMB_EBE_VARIANT"); variant1 = e ? atoi(e) : 5;
    e = getenv("FEMB_EBE_VARIANT4"); variant4 = e ? atoi(e) : 5;
  }
  const bool dot = dot_partials != nullptr;
#define NODEV(NBT, NB, T, MINB) launch_node_variant<NBT, NB, T, 128, MINB>(h, P, x, y, masked, dot_partials, scal_out, ticket, done, ctas)
#define NODE2(NBT, NB, T, MINB) launch_node2_variant<NBT, NB, T, 128, MINB>(h, P, x, y, masked, dot_partials, scal_out, ticket, done, ctas)
#define RECV(NBT, NB, T, MINB, UNR) launch_rec_variant<NBT, NB, T, 128, MINB, UNR>(h, x, y, masked, dot_partials, scal_out, ticket, done, ctas)
  if ((nb == 1 && variant1 >= 20) || (nb == 4 && variant4 >= 20)) {
    int rc = ensure_ebe_records(h);
    if (rc) return rc;
  }
  if (nb == 1 && variant1 > 0) {
    switch (variant1) {
      case 21: RECV(1, 1, 1, 6, 2); break;
      case 22: RECV(1, 1, 2, 6, 2); break;
      case 23: RECV(1, 1, 1, 8, 1); break;
      case 24: RECV(1, 1, 2, 8, 1); break;
      case 25: RECV(1, 1, 2, 5, 3); break;
      case 26: RECV(1, 1, 4, 6, 2); break;
      case 11: NODE2(1, 1, 1, 4); break;
      case 12: NODE2(1, 1, 2, 4); break;
      case 13: NODE2(1, 1, 1, 5); break;
      case 14: NODE2(1, 1, 2, 5); break;
      case 15: NODE2(1, 1, 1, 6); break;
      case 16: NODE2(1, 1, 2, 6); break;
      case 1: NODEV(1, 1, 1, 5); break;
      case 2: NODEV(1, 1, 2, 5); break;
      case 3: NODEV(1, 1, 4, 5); break;
      case 4: NODEV(1, 1, 1, 4); break;
      case 5: NODEV(1, 1, 2, 4); break;
      case 6: NODEV(1, 1, 2, 6); break;
      default: NODEV(1, 1, 1, 6); break;
    }
  } else if (nb == 4 && variant4 > 0) {
    switch (variant4) {
      case 21: RECV(4, 2, 1, 4, 2); break;
      case 22: RECV(4, 2, 2, 4, 2); break;
      case 23: RECV(4, 4, 1, 3, 1); break;
      case 24: RECV(4, 4, 2, 3, 1); break;
      case 25: RECV(4, 2, 1, 5, 1); break;
      case 26: RECV(4, 2, 2, 5, 1); break;
      case 11: NODE2(4, 2, 1, 3); break;
      case 12: NODE2(4, 2, 2, 3); break;
      case 13: NODE2(4, 2, 1, 4); break;
      case 14: NODE2(4, 2, 2, 4); break;
      case 15: NODE2(4, 4, 2, 2); break;
      case 16: NODE2(4, 1, 1, 5); break;
      case 1: NODEV(4, 4, 1, 2); break;
      case 2: NODEV(4, 2, 1, 4); break;
      case 3: NODEV(4, 2, 2, 4); break;
      case 4: NODEV(4, 1, 1, 5); break;
      case 5: NODEV(4, 2, 1, 3); break;
      case 6: NODEV(4, 2, 2, 3); break;
      default: NODEV(4, 4, 2, 2); break;
    }
#undef NODEV
#undef NODE2
#undef RECV
  } else if (nb == 1) {
    const int per_sm = ctas > 0 ? std::min(ctas, 8) : 6;
    const int grid = std::min(n_tiles, h->num_sms * per_sm);
#define EBE1(M, D) frame_ebe_kernel<1, kEbeThreads, M, D, 6><<<grid, kEbeThreads, 0, h->stream>>>( \
    P, pp, h->free_mask.p, x, y, dot_partials, pstride, scal_out, ticket, done)
    if (masked && dot) EBE1(true, true);
    else if (masked) EBE1(true, false);
    else EBE1(false, false);
#undef EBE1
  } else if (nb == 4) {
    const int per_sm = ctas > 0 ? std::min(ctas, 8) : 3;
    const int grid = std::min(n_tiles, h->num_sms * per_sm);
#define EBE4(M, D) frame_ebe_kernel<4, kEbeThreads, M, D, 3><<<grid, kEbeThreads, 0, h->stream>>>( \
    P, pp, h->free_mask.p, x, y, dot_partials, pstride, scal_out, ticket, done)
    if (masked && dot) EBE4(true, true);
    else if (masked) EBE4(true, false);
    else EBE4(false, false);
#undef EBE4
  } else {
    return fail(h, FEMB_ERR_ARG, "matrix-free operator: 1 or 4 vectors");
  }
  h->launches++;
  FEMB_CUDA(h, cudaGetLastError());
  return FEMB_OK;
}

}  // namespace femb
