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
#include "elements.cuh"
#include "pcg_common.cuh"

namespace femb {

struct EbeDev {
  const int4* rec;        // (n_pairs) {node, other, blk, sec | a<<24 | pos<<25}
  const int4* node_rec;   // (n_nodes) {first pair, pair count, diagonal block, 0}
  const int4* tiles;      // (n_tiles) {first node, node count, first pair, pair count}
  int n_tiles;
};

// stiffness part of the element record (the lumped-mass terms of FrameRec are dead code here)
struct KRec {
  double t[3], n1[3], n2[3];
  double ax, tor, k11z, k11y, c12z, c12y, k23z, k23y, d22z, d22y;   // c12 = s_a*k12, d22 = k22 - k23
};

__device__ __forceinline__ void krec_from(const FrameParams& P, const FrameIn& in, int a, KRec& k) {
  FrameRec R;
  frame_record_from(P, in, R);
  const double sa = a ? -1.0 : 1.0;
#pragma unroll
  for (int i = 0; i < 3; ++i) { k.t[i] = R.t[i]; k.n1[i] = R.n1[i]; k.n2[i] = R.n2[i]; }
  k.ax = R.ax; k.tor = R.tor; k.k11z = R.k11z; k.k11y = R.k11y;
  k.c12z = sa * R.k12z; k.c12y = sa * R.k12y;
  k.k23z = R.k23z; k.k23y = R.k23y;
  k.d22z = R.k22z - R.k23z; k.d22y = R.k22y - R.k23y;
}

__device__ __forceinline__ double dot3(const double* a, double x, double y, double z) {
  return a[0] * x + a[1] * y + a[2] * z;
}

// out[0..5] = K_e[a][a] ua + K_e[a][1-a] uo in global axes.  With d = translations, th = rotations
// and the projections on (t, n1, n2) the two block rows of elements.cuh collapse to
//   force : ax (t.dd) t + [k11z (n1.dd) + c12z (n2.ts)] n1 + [k11y (n2.dd) - c12y (n1.ts)] n2
//   moment: tor (t.td) t + [-c12y (n2.dd) + k23y (n1.ts) + d22y (n1.ta)] n1
//                        + [ c12z (n1.dd) + k23z (n2.ts) + d22z (n2.ta)] n2
// where dd = d_a - d_o, td = th_a - th_o, ts = th_a + th_o, ta = th_a.
__device__ __forceinline__ void ebe_apply(const KRec& k, const double* ua, const double* uo, double* out) {
  const double ddx = ua[0] - uo[0], ddy = ua[1] - uo[1], ddz = ua[2] - uo[2];
  const double tdx = ua[3] - uo[3], tdy = ua[4] - uo[4], tdz = ua[5] - uo[5];
  const double tsx = ua[3] + uo[3], tsy = ua[4] + uo[4], tsz = ua[5] + uo[5];
  const double dt = dot3(k.t, ddx, ddy, ddz), d1 = dot3(k.n1, ddx, ddy, ddz), d2 = dot3(k.n2, ddx, ddy, ddz);
  const double tt = dot3(k.t, tdx, tdy, tdz);
  const double s1 = dot3(k.n1, tsx, tsy, tsz), s2 = dot3(k.n2, tsx, tsy, tsz);
  const double a1 = dot3(k.n1, ua[3], ua[4], ua[5]), a2 = dot3(k.n2, ua[3], ua[4], ua[5]);
  const double ft = k.ax * dt;
  const double f1 = k.k11z * d1 + k.c12z * s2;
  const double f2 = k.k11y * d2 - k.c12y * s1;
  const double mt = k.tor * tt;
  const double m1 = k.k23y * s1 + k.d22y * a1 - k.c12y * d2;
  const double m2 = k.k23z * s2 + k.d22z * a2 + k.c12z * d1;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    out[i] = ft * k.t[i] + f1 * k.n1[i] + f2 * k.n2[i];
    out[3 + i] = mt * k.t[i] + m1 * k.n1[i] + m2 * k.n2[i];
  }
}

__device__ __forceinline__ void load6(const double* __restrict__ x, int node, double* u) {
  const double2* p = reinterpret_cast<const double2*>(x + (size_t)node * 6);
  const double2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
  u[0] = a.x; u[1] = a.y; u[2] = b.x; u[3] = b.y; u[4] = c.x; u[5] = c.y;
}

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
  if (env) op = env;
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

// y = K_ff x (masked) or K x; nb = 1 (plain) or 4 (interleaved).  dot_partials != null: (x_q, y_q)
// -> scal_out[q], ticket = reduction ticket slot, done = early-exit flag (may be null).
int launch_ebe(femb_handle* h, const double* x, double* y, int nb, bool masked, double* dot_partials,
               double* scal_out, int* ticket, const int* done) {
  const Symbolic& S = h->sym;
  const int n_tiles = (int)S.pair_tile_ptr.size() - 1;
  if (n_tiles <= 0) return FEMB_OK;
  EbeDev pp{reinterpret_cast<const int4*>(h->pair_rec.p), reinterpret_cast<const int4*>(h->pair_node_rec.p),
            reinterpret_cast<const int4*>(h->pair_tiles.p), n_tiles};
  const FrameParams P = ebe_params(h);
  const int pstride = h->num_sms * 8;
  static int ctas = -1;
  if (ctas < 0) { const char* e = getenv("FEMB_EBE_CTAS"); ctas = e ? atoi(e) : 0; }
  const bool dot = dot_partials != nullptr;
  if (nb == 1) {
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
