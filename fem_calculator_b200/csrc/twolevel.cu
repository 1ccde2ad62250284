// Two-level (Jacobi + rigid-body coarse space) preconditioner for the matrix-free frame PCG.
//
// Why: with the matrix-free operator an iteration of the Jacobi-PCG costs 41 us at 1M DOF, so the
// static solve K_ff u = f (BeamSolver.py:417) is bound by the iteration COUNT — 6,931 at 56^3 nodes,
// growing linearly with the lattice edge.  The modes Jacobi cannot see are the smooth ones, and for a
// frame those are locally rigid-body motions (every unconstrained element of BeamSolver.py:646-660
// has the six rigid motions in its kernel).  So the nodes are grouped into aggregates (coarse.cpp)
// and the additive two-level preconditioner
//        M^-1 r = omega D^-1 r + P (P^T A P)^-1 P^T r,      A = the masked operator K_ff + I_fixed,
// is used, P = six rigid-body modes per aggregate (translations, rotations about the centroid) with
// the rows of fixed DOFs zeroed (the Jacobi term carries a weight omega = 2).  1,324 instead of 6,931
// iterations at 1M DOF with 444 aggregates.
//
// Pieces (all deterministic — fixed-order sums, no float atomics):
//   tl_centroid_kernel        centroid of every aggregate (per assembled K: coordinates may change)
//   tl_coarse_assemble_kernel Kc = P^T A P from the assembled BSR K: one CTA per coarse block row, one
//                             thread per entry of a coarse 6x6 block, work items = (neighbour slot, node
//                             chunk), chunk partials added in chunk order
//   coarse_invert (direct.cu) Kc^-1 explicitly, on the DMMA Cholesky kernels
//   tl_update_kernel          one CTA per aggregate: the Chronopoulos-Gear vector update of its nodes
//                             fused with the restriction rc = P^T r of the new residual
//   tl_coarse_z_kernel        one CTA per aggregate: its six rows of y = Kc^-1 rc, then
//                             z = omega D^-1 r + P y for its nodes and the (r, z) partial
// The iteration is  operator (ebe.cu, linked reductions) -> tl_update -> tl_coarse_z : three kernels,
// the reductions travel as published partial sums exactly as in the linked Jacobi-PCG (pcg_common.cuh).
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "pcg_common.cuh"

namespace femb {

constexpr int kTlThreads = 384;       // update / coarse-z CTAs (one per aggregate); one thread per (node, DOF pair)
constexpr int kTlAsmGroups = 28;      // coarse-assembly CTA: 28 slot groups x 36 entries = 1008 threads
constexpr int kTlMaxAgg = 1024;       // one published partial per aggregate (partials arrays hold num_sms * 8)
constexpr double kTlOmega = 2.0;      // weight of the Jacobi term in the additive preconditioner (FEMB_TL_OMEGA):
                                      // 1M DOF, 444 aggregates: omega 1 / 2 / 3 -> 1372 / 1324 / 1342 iterations
constexpr double kTlRidge = 1e-8;     // relative ridge on diag(Kc): keeps the factorisation positive when the
                                      // free DOFs of an aggregate do not carry all six rigid-body modes

struct TlDev {
  const int32_t* agg_ptr;
  const int32_t* agg_nodes;
  const double* centroid;   // (n_agg,3)
  const double* xyz;
  const uint8_t* free_mask;
  const double* inv;        // (n_pad, n_pad)
  double* rc;               // (n_pad)
  int n_agg, n_pad;
  double omega;             // weight of the Jacobi term: z = omega D^-1 r + P Kc^-1 P^T r
};

__global__ void tl_centroid_kernel(const int32_t* __restrict__ agg_ptr, const int32_t* __restrict__ agg_nodes,
                                   const double* __restrict__ xyz, double* __restrict__ centroid, int n_agg) {
  __shared__ double s_part[3 * kTlThreads / 32];
  const int I = blockIdx.x;
  const int first = agg_ptr[I], cnt = agg_ptr[I + 1] - first;
  double v[3] = {0.0, 0.0, 0.0};
  for (int k = threadIdx.x; k < cnt; k += kTlThreads) {
    const double* p = xyz + 3 * (size_t)agg_nodes[first + k];
    v[0] += p[0]; v[1] += p[1]; v[2] += p[2];
  }
  block_sum_all<kTlThreads, 3>(v, s_part);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int d = 0; d < 3; ++d) centroid[3 * (size_t)I + d] = cnt > 0 ? v[d] / cnt : 0.0;
  }
}

// column m of the rigid-body block T = [I, -[rho]x; 0, I] of a node at offset rho from the centroid:
// m < 3 translation e_m; m >= 3 rotation about axis m-3: u = e x rho, theta = e.  Masked by the node's
// free-DOF flags (rows of fixed DOFs are zero in P).
__device__ __forceinline__ void rbm_column(int m, double rx, double ry, double rz, const uint8_t* fm, double* t) {
#pragma unroll
  for (int a = 0; a < 6; ++a) t[a] = 0.0;
  if (m < 3) {
#pragma unroll
    for (int a = 0; a < 3; ++a) t[a] = (a == m) ? 1.0 : 0.0;
  } else if (m == 3) {
    t[1] = -rz; t[2] = ry; t[3] = 1.0;
  } else if (m == 4) {
    t[0] = rz; t[2] = -rx; t[4] = 1.0;
  } else {
    t[0] = -ry; t[1] = rx; t[5] = 1.0;
  }
#pragma unroll
  for (int a = 0; a < 6; ++a)
    if (!fm[a]) t[a] = 0.0;
}

// Kc(I, J) = sum over stored blocks (i in I, j in J) of T_i^T D_i K_ij D_j T_j, written into the top
// left corner of the augmented work matrix (row-major, leading dimension ld).  One CTA per coarse
// block row I, 28 groups of 36 threads (one thread per entry (r, c) of a coarse 6x6 block).  A work
// item is (slot s of I's neighbour list, chunk of kTlAsmChunk nodes of I): its group walks the block
// rows of the chunk in storage order, adds the blocks of slot s and parks the partial sum in `scratch`;
// after a barrier thread (s, e) adds the chunk partials of its entry in chunk order — fixed order, no
// atomics.  Blocks between free interior nodes cancel to round-off (rigid motions are in the kernel
// of their elements); they are summed like the others so no topology / BC case needs a special path.
// (First version: every thread scanned ALL block rows of the aggregate for its slot — the 36 threads
// of the self slot walked ~2,600 blocks one after the other; 20 ms of the 24 ms setup at 1M DOF.)
constexpr int kTlAsmChunk = 16;

__global__ void __launch_bounds__(kTlAsmGroups * 36)
tl_coarse_assemble_kernel(const int32_t* __restrict__ agg_ptr, const int32_t* __restrict__ agg_nodes,
                          const int32_t* __restrict__ nbr_ptr, const int32_t* __restrict__ nbr,
                          const int32_t* __restrict__ blk_slot, const int32_t* __restrict__ rowptr,
                          const int32_t* __restrict__ colidx, const double* __restrict__ Kvals,
                          const double* __restrict__ xyz, const double* __restrict__ centroid,
                          const uint8_t* __restrict__ free_mask, double* __restrict__ aug, int64_t ld,
                          double* __restrict__ scratch, int64_t scratch_per_agg) {
  const int I = blockIdx.x;
  const int e = threadIdx.x % 36, g = threadIdx.x / 36;
  const int r = e / 6, c = e % 6;
  const int first = agg_ptr[I], cnt = agg_ptr[I + 1] - first;
  const int s0 = nbr_ptr[I], ns = nbr_ptr[I + 1] - s0;
  const int nchunk = (cnt + kTlAsmChunk - 1) / kTlAsmChunk;
  const double cIx = centroid[3 * (size_t)I], cIy = centroid[3 * (size_t)I + 1], cIz = centroid[3 * (size_t)I + 2];
  double* part = scratch + (size_t)I * scratch_per_agg;       // [slot][chunk][36]
  for (int item = g; item < ns * nchunk; item += kTlAsmGroups) {
    const int ch = item % nchunk, s = item / nchunk;   // consecutive groups take consecutive chunks of one slot
    const int J = nbr[s0 + s];
    const double cJx = centroid[3 * (size_t)J], cJy = centroid[3 * (size_t)J + 1], cJz = centroid[3 * (size_t)J + 2];
    double acc = 0.0;
    const int k1 = min(cnt, (ch + 1) * kTlAsmChunk);
    for (int k = ch * kTlAsmChunk; k < k1; ++k) {
      const int i = agg_nodes[first + k];
      const int b0 = rowptr[i], b1 = rowptr[i + 1];
      bool have_ti = false;
      double ti[6];
      for (int b = b0; b < b1; ++b) {
        if (blk_slot[b] != s) continue;
        if (!have_ti) {
          const double* pi = xyz + 3 * (size_t)i;
          rbm_column(r, pi[0] - cIx, pi[1] - cIy, pi[2] - cIz, free_mask + 6 * (size_t)i, ti);
          have_ti = true;
        }
        const int j = colidx[b];
        const double* pj = xyz + 3 * (size_t)j;
        double tj[6];
        rbm_column(c, pj[0] - cJx, pj[1] - cJy, pj[2] - cJz, free_mask + 6 * (size_t)j, tj);
        const double* kb = Kvals + 36 * (size_t)b;
        double sum = 0.0;
#pragma unroll
        for (int a = 0; a < 6; ++a) {
          double row = 0.0;
#pragma unroll
          for (int d = 0; d < 6; ++d) row += kb[a * 6 + d] * tj[d];
          sum += ti[a] * row;
        }
        acc += sum;
      }
    }
    part[((size_t)s * nchunk + ch) * 36 + e] = acc;
  }
  __syncthreads();
  for (int s = g; s < ns; s += kTlAsmGroups) {
    const int J = nbr[s0 + s];
    double acc = 0.0;
    for (int ch = 0; ch < nchunk; ++ch) acc += part[((size_t)s * nchunk + ch) * 36 + e];
    if (J == I && r == c) acc = (acc > 0.0) ? acc * (1.0 + kTlRidge) : 1.0;   // fully fixed / empty aggregate: identity
    aug[(size_t)(6 * I + r) * ld + 6 * J + c] = acc;
  }
}

// identity on the padding of Kc and in the lower left block of the augmented matrix
__global__ void tl_aug_identity_kernel(double* __restrict__ aug, int64_t n, int64_t n_pad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad) return;
  const int64_t ld = 2 * n_pad;
  if (i >= n) aug[i * ld + i] = 1.0;
  aug[(n_pad + i) * ld + i] = 1.0;
}

__device__ __forceinline__ void ld6(const double* __restrict__ v, int node, double* u) {
  const double2* p = reinterpret_cast<const double2*>(v + (size_t)node * 6);
  const double2 a = p[0], b = p[1], c = p[2];
  u[0] = a.x; u[1] = a.y; u[2] = b.x; u[3] = b.y; u[4] = c.x; u[5] = c.y;
}
__device__ __forceinline__ void st6(double* __restrict__ v, int node, const double* u) {
  double2* p = reinterpret_cast<double2*>(v + (size_t)node * 6);
  p[0] = make_double2(u[0], u[1]); p[1] = make_double2(u[2], u[3]); p[2] = make_double2(u[4], u[5]);
}

// a thread of the per-aggregate kernels owns one 16-byte DOF pair (2 * part, 2 * part + 1) of a node:
// consecutive lanes touch consecutive 16-byte chunks, and a thread carries 2 instead of 6 values per
// vector (the thread-per-node form needed 96 registers and ran at 18 % warp occupancy, 24 us per launch)
__device__ __forceinline__ double2 ld2(const double* __restrict__ v, int node, int part) {
  return *(reinterpret_cast<const double2*>(v + (size_t)node * 6) + part);
}
__device__ __forceinline__ void st2(double* __restrict__ v, int node, int part, double2 a) {
  *(reinterpret_cast<double2*>(v + (size_t)node * 6) + part) = a;
}
// the pair as a 6-vector with zeros elsewhere (adding zeros is exact, so restrict_add stays the one formula)
__device__ __forceinline__ void expand_pair(int part, double2 a, double* u) {
#pragma unroll
  for (int c = 0; c < 3; ++c) { u[2 * c] = (c == part) ? a.x : 0.0; u[2 * c + 1] = (c == part) ? a.y : 0.0; }
}

// P^T r contribution of one node: translations take the force, rotations take rho x force + moment
__device__ __forceinline__ void restrict_add(const double* r, double rx, double ry, double rz, double* acc) {
  acc[0] += r[0]; acc[1] += r[1]; acc[2] += r[2];
  acc[3] += ry * r[2] - rz * r[1] + r[3];
  acc[4] += rz * r[0] - rx * r[2] + r[4];
  acc[5] += rx * r[1] - ry * r[0] + r[5];
}

// init: x = 0, r = b, p = q = 0 on the aggregate's nodes; rc = P^T b; publishes ||b||^2 into buffer 0
// (gamma_0 follows from tl_coarse_z_kernel with wr = 0)
__global__ void __launch_bounds__(kTlThreads, 3)
tl_init_kernel(const TlDev T, const double* __restrict__ b, double* __restrict__ x, double* __restrict__ r,
               double* __restrict__ p, double* __restrict__ q, const PcgLink L) {
  __shared__ double s_part[7 * kTlThreads / 32];
  const int I = blockIdx.x;
  const int first = T.agg_ptr[I], cnt = T.agg_ptr[I + 1] - first;
  const double cx = T.centroid[3 * (size_t)I], cy = T.centroid[3 * (size_t)I + 1], cz = T.centroid[3 * (size_t)I + 2];
  double v[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  for (int w = threadIdx.x; w < 3 * cnt; w += kTlThreads) {
    const int node = T.agg_nodes[first + w / 3], part = w % 3;
    const double2 bv = ld2(b, node, part);
    const double2 zz = make_double2(0.0, 0.0);
    st2(r, node, part, bv); st2(x, node, part, zz); st2(p, node, part, zz); st2(q, node, part, zz);
    const double* pp = T.xyz + 3 * (size_t)node;
    double b6[6];
    expand_pair(part, bv, b6);
    restrict_add(b6, pp[0] - cx, pp[1] - cy, pp[2] - cz, v);
    v[6] += bv.x * bv.x + bv.y * bv.y;
  }
  block_sum_all<kTlThreads, 7>(v, s_part);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int a = 0; a < 6; ++a) T.rc[6 * (size_t)I + a] = v[a];
  }
  if (threadIdx.x == 0) L.upd_partials[L.pstride + I] = v[6];
  if (I == 0 && threadIdx.x == 0) { L.flags[Flag::DONE] = 0; L.flags[Flag::ITERS] = 0; }
}

// update(it): consumes delta (operator) and gamma (coarse-z of the previous iteration / init) like
// pcg_update_linked_kernel, then p = z + beta p, q = s + beta q, x += alpha p, r -= alpha q on the
// aggregate's nodes, rc = P^T r, and publishes ||r||^2 into buffer (it + 1) & 1
__global__ void __launch_bounds__(kTlThreads, 3)
tl_update_kernel(const TlDev T, const double* __restrict__ z, const double* __restrict__ s, double* __restrict__ p,
                 double* __restrict__ q, double* __restrict__ x, double* __restrict__ r, const PcgLink L) {
  __shared__ double s_part[7 * kTlThreads / 32];
  if (L.flags[Flag::DONE]) return;
  const int rd = L.it & 1, wr = rd ^ 1;
  double tot[2] = {0.0, 0.0};
  {
    const double* pu = L.upd_partials + (size_t)rd * 2 * L.pstride;
    for (int i = threadIdx.x; i < L.n_op; i += kTlThreads) tot[0] += __ldcg(L.op_partials + i);
    for (int i = threadIdx.x; i < L.n_upd; i += kTlThreads) tot[1] += __ldcg(pu + i);
  }
  block_sum_all<kTlThreads, 2>(tot, s_part);
  const double delta = tot[0], gamma = tot[1];
  const bool first_it = (L.it == 0);
  const double beta = first_it ? 0.0 : gamma / L.scal[Scal::RZ0 + rd];
  const double den = first_it ? delta : delta - beta * gamma / L.scal[Scal::ALPHA + rd];
  const bool bad = !(den > 0.0);           // K_ff (or the preconditioner) not positive definite along p
  const double alpha = bad ? 0.0 : gamma / den;
  const int I = blockIdx.x;
  if (I == 0 && threadIdx.x == 0) {
    L.scal[Scal::RZ0 + wr] = gamma;
    L.scal[Scal::ALPHA + wr] = alpha;
    L.scal[Scal::PQ] = delta;
    if (bad) L.flags[Flag::DONE] = 2;
  }
  if (bad) return;
  const int first = T.agg_ptr[I], cnt = T.agg_ptr[I + 1] - first;
  const double cx = T.centroid[3 * (size_t)I], cy = T.centroid[3 * (size_t)I + 1], cz = T.centroid[3 * (size_t)I + 2];
  double v[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
  for (int w = threadIdx.x; w < 3 * cnt; w += kTlThreads) {
    const int node = T.agg_nodes[first + w / 3], part = w % 3;
    const double2 zv = ld2(z, node, part), sv = ld2(s, node, part);
    double2 pv = ld2(p, node, part), qv = ld2(q, node, part), xv = ld2(x, node, part), rv = ld2(r, node, part);
    const double* pp = T.xyz + 3 * (size_t)node;
    const double rx = pp[0] - cx, ry = pp[1] - cy, rz = pp[2] - cz;
    pv.x = zv.x + beta * pv.x; pv.y = zv.y + beta * pv.y;
    qv.x = sv.x + beta * qv.x; qv.y = sv.y + beta * qv.y;
    xv.x += alpha * pv.x; xv.y += alpha * pv.y;
    rv.x -= alpha * qv.x; rv.y -= alpha * qv.y;
    st2(p, node, part, pv); st2(q, node, part, qv); st2(x, node, part, xv); st2(r, node, part, rv);
    v[6] += rv.x * rv.x + rv.y * rv.y;
    double r6[6];
    expand_pair(part, rv, r6);
    restrict_add(r6, rx, ry, rz, v);
  }
  block_sum_all<kTlThreads, 7>(v, s_part);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int a = 0; a < 6; ++a) T.rc[6 * (size_t)I + a] = v[a];
  }
  if (threadIdx.x == 0) L.upd_partials[(size_t)wr * 2 * L.pstride + L.pstride + I] = v[6];
}

// y_I = (Kc^-1 rc)[6I .. 6I+6), then z = D^-1 r + P y on the aggregate's nodes; publishes the (r, z)
// partial into buffer wr.  The inverse is symmetric: rows are read, contiguously.
__global__ void __launch_bounds__(kTlThreads, 3)
tl_coarse_z_kernel(const TlDev T, const double* __restrict__ dinv, const double* __restrict__ r,
                   double* __restrict__ z, int wr, const PcgLink L) {
  __shared__ double s_part[6 * kTlThreads / 32];
  if (L.flags[Flag::DONE]) return;
  const int I = blockIdx.x;
  double y[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  {
    const double2* rc2 = reinterpret_cast<const double2*>(T.rc);
    const double* rows = T.inv + (size_t)6 * I * T.n_pad;
    const int n2 = T.n_pad >> 1;
#pragma unroll 2
    for (int j = threadIdx.x; j < n2; j += kTlThreads) {
      const double2 rv = __ldcg(rc2 + j);
#pragma unroll
      for (int m = 0; m < 6; ++m) {
        const double2 w = __ldg(reinterpret_cast<const double2*>(rows + (size_t)m * T.n_pad) + j);
        y[m] += w.x * rv.x + w.y * rv.y;
      }
    }
    block_sum_all<kTlThreads, 6>(y, s_part);
  }
  const int first = T.agg_ptr[I], cnt = T.agg_ptr[I + 1] - first;
  const double cx = T.centroid[3 * (size_t)I], cy = T.centroid[3 * (size_t)I + 1], cz = T.centroid[3 * (size_t)I + 2];
  double g[1] = {0.0};
  for (int w = threadIdx.x; w < 3 * cnt; w += kTlThreads) {
    const int node = T.agg_nodes[first + w / 3], part = w % 3;
    const double2 rv = ld2(r, node, part), dv = ld2(dinv, node, part);
    const double* pp = T.xyz + 3 * (size_t)node;
    const double rx = pp[0] - cx, ry = pp[1] - cy, rz = pp[2] - cz;
    const uint8_t* fm = T.free_mask + 6 * (size_t)node + 2 * part;
    double c[6];                                   // (P y)_node = (y_t + y_w x rho, y_w)
    c[0] = y[0] + y[4] * rz - y[5] * ry;
    c[1] = y[1] + y[5] * rx - y[3] * rz;
    c[2] = y[2] + y[3] * ry - y[4] * rx;
    c[3] = y[3]; c[4] = y[4]; c[5] = y[5];
    double ca = c[0], cb = c[1];
    if (part == 1) { ca = c[2]; cb = c[3]; }
    if (part == 2) { ca = c[4]; cb = c[5]; }
    double2 zv;
    zv.x = T.omega * dv.x * rv.x + (fm[0] ? ca : 0.0);
    zv.y = T.omega * dv.y * rv.y + (fm[1] ? cb : 0.0);
    g[0] += rv.x * zv.x + rv.y * zv.y;
    st2(z, node, part, zv);
  }
  block_sum_all<kTlThreads, 1>(g, s_part);
  if (threadIdx.x == 0) L.upd_partials[(size_t)wr * 2 * L.pstride + I] = g[0];
}

// one CTA: the operator-side decision alone (host poll), as pcg_decide_linked_kernel
__global__ void __launch_bounds__(128)
tl_decide_kernel(const PcgLink L) {
  __shared__ double s_part[2 * 128 / 32];
  if (L.flags[Flag::DONE]) return;
  pcg_link_decide<128>(L, s_part);
}

static int coarse_target_aggregates(const femb_handle* h) {
  int want = 3 * h->num_sms;                     // three CTAs per SM in the per-aggregate kernels
  if (const char* e = getenv("FEMB_COARSE_AGGS")) { const int v = atoi(e); if (v > 0) want = v; }
  const int64_t by_size = std::max<int64_t>(1, h->n_nodes / 24);   // at least ~24 nodes per aggregate
  want = (int)std::min<int64_t>(want, by_size);
  return std::max(1, std::min(want, kTlMaxAgg));
}

// aggregate tables for the current topology (host RCB on a snapshot of the device coordinates)
static int ensure_coarse_symbolic(femb_handle* h) {
  if (h->coarse_sym_ok) return FEMB_OK;
  std::vector<double> hx((size_t)h->n_nodes * 3);
  FEMB_CUDA(h, download(hx.data(), h->xyz.p, hx.size() * 8, h->stream));
  FEMB_CUDA(h, cudaStreamSynchronize(h->stream));
  const int n_agg = coarse_target_aggregates(h);
  std::vector<int32_t> agg;
  build_aggregates(h->n_nodes, hx.data(), n_agg, agg);
  CoarseSym C;
  build_coarse_symbolic(h->sym, agg, n_agg, C);
  FEMB_CUDA(h, upload(h->agg_ptr, C.agg_ptr, h->stream));
  FEMB_CUDA(h, upload(h->agg_nodes, C.agg_nodes, h->stream));
  FEMB_CUDA(h, upload(h->agg_nbr_ptr, C.nbr_ptr, h->stream));
  FEMB_CUDA(h, upload(h->agg_nbr, C.nbr, h->stream));
  FEMB_CUDA(h, upload(h->blk_slot, C.blk_slot, h->stream));
  FEMB_CUDA(h, cudaStreamSynchronize(h->stream));   // C is a local
  h->coarse_n_agg = n_agg;
  h->coarse_max_nbr = C.max_nbr;
  h->coarse_n = 6 * (int64_t)n_agg;
  h->coarse_n_pad = (h->coarse_n + 63) / 64 * 64;
  FEMB_CUDA(h, h->agg_centroid.alloc((size_t)n_agg * 3));
  FEMB_CUDA(h, h->coarse_inv.alloc((size_t)h->coarse_n_pad * h->coarse_n_pad));
  FEMB_CUDA(h, h->coarse_r.alloc((size_t)h->coarse_n_pad * 4));
  int max_cnt = 0;
  for (int a = 0; a < n_agg; ++a) max_cnt = std::max(max_cnt, C.agg_ptr[a + 1] - C.agg_ptr[a]);
  h->coarse_scratch_per_agg = (int64_t)std::max(1, C.max_nbr) * ((max_cnt + kTlAsmChunk - 1) / kTlAsmChunk) * 36;
  FEMB_CUDA(h, h->coarse_scratch.alloc((size_t)h->coarse_scratch_per_agg * n_agg));
  h->coarse_sym_ok = true;
  h->coarse_num_ok = false;
  return FEMB_OK;
}

// centroids, Kc = P^T A P, Kc^-1 for the current K and BC mask
static int ensure_coarse_numeric(femb_handle* h) {
  int rc = ensure_coarse_symbolic(h);
  if (rc) return rc;
  if (h->coarse_num_ok || h->coarse_failed) return FEMB_OK;
  const int n_agg = h->coarse_n_agg;
  const int64_t n = h->coarse_n, n_pad = h->coarse_n_pad, m = 2 * n_pad;
  FEMB_CUDA(h, h->coarse_aug.ensure((size_t)m * m));
  const bool trace = getenv("FEMB_TRACE") != nullptr;
  cudaEvent_t te[3] = {nullptr, nullptr, nullptr};
  if (trace) {
    for (auto& e : te) cudaEventCreate(&e);
    cudaEventRecord(te[0], h->stream);
  }
  FEMB_CUDA(h, cudaMemsetAsync(h->coarse_aug.p, 0, (size_t)m * m * sizeof(double), h->stream));
  FEMB_CUDA(h, cudaMemsetAsync(h->coarse_r.p, 0, h->coarse_r.bytes(), h->stream));
  tl_centroid_kernel<<<n_agg, kTlThreads, 0, h->stream>>>(h->agg_ptr.p, h->agg_nodes.p, h->xyz.p, h->agg_centroid.p, n_agg);
  tl_aug_identity_kernel<<<(unsigned)((n_pad + 255) / 256), 256, 0, h->stream>>>(h->coarse_aug.p, n, n_pad);
  tl_coarse_assemble_kernel<<<n_agg, kTlAsmGroups * 36, 0, h->stream>>>(
      h->agg_ptr.p, h->agg_nodes.p, h->agg_nbr_ptr.p, h->agg_nbr.p, h->blk_slot.p, h->rowptr.p, h->colidx.p,
      h->Kvals.p, h->xyz.p, h->agg_centroid.p, h->free_mask.p, h->coarse_aug.p, m, h->coarse_scratch.p,
      h->coarse_scratch_per_agg);
  h->launches += 3;
  FEMB_CUDA(h, cudaGetLastError());
  if (trace) cudaEventRecord(te[1], h->stream);
  bool ok = false;
  rc = coarse_invert(h, h->coarse_aug.p, n_pad, h->coarse_inv.p, &ok);
  if (rc) return rc;
  if (trace) {
    cudaEventRecord(te[2], h->stream);
    cudaEventSynchronize(te[2]);
    float a = 0.f, b = 0.f;
    cudaEventElapsedTime(&a, te[0], te[1]);
    cudaEventElapsedTime(&b, te[1], te[2]);
    for (auto& e : te) cudaEventDestroy(e);
    fprintf(stderr, "[femb trace] two-level numeric setup: coarse dim %lld, Galerkin assembly %.3f ms, inversion %.3f ms\n",
            (long long)n, a, b);
  }
  h->coarse_num_ok = ok;
  h->coarse_failed = !ok;
  return FEMB_OK;
}

constexpr int64_t kTlAutoNodes = 50000;   // FEMB_PRECOND_AUTO: below this the coarse setup costs more than it saves

bool twolevel_applicable(const femb_handle* h, const femb_solve_opts& o) {
  const bool want = o.precond == FEMB_PRECOND_TWO_LEVEL ||
                    (o.precond == FEMB_PRECOND_AUTO && h->n_nodes >= kTlAutoNodes);
  return want && h->bs == 6 && ebe_selected(h, o.op);
}

int pcg_twolevel(femb_handle* h, const femb_solve_opts& o, const double* d_b, femb_stats* st) {
  const int pstride = h->num_sms * 8;
  int rc = setup_precond_public(h, FEMB_PRECOND_JACOBI);
  if (rc) return rc;
  rc = ensure_coarse_numeric(h);
  if (rc) return rc;
  if (!h->coarse_num_ok) {
    // the Galerkin matrix could not be factored: plain Jacobi (stats.coarse_dim stays 0)
    femb_solve_opts oj = o;
    oj.precond = FEMB_PRECOND_JACOBI;
    return pcg_solve_rhs(h, oj, d_b, st);
  }
  FEMB_CUDA(h, h->fpartials.ensure((size_t)pstride * 6));
  FEMB_CUDA(h, cudaMemsetAsync(h->flags.p, 0, sizeof(int32_t) * Flag::COUNT, h->stream));
  FEMB_CUDA(h, cudaMemsetAsync(h->scal.p, 0, sizeof(double) * Scal::COUNT, h->stream));
  const int n_agg = h->coarse_n_agg;
  TlDev T;
  T.agg_ptr = h->agg_ptr.p; T.agg_nodes = h->agg_nodes.p; T.centroid = h->agg_centroid.p; T.xyz = h->xyz.p;
  T.free_mask = h->free_mask.p; T.inv = h->coarse_inv.p; T.rc = h->coarse_r.p; T.n_agg = n_agg; T.n_pad = (int)h->coarse_n_pad;
  T.omega = kTlOmega;
  if (const char* e = getenv("FEMB_TL_OMEGA")) { const double v = atof(e); if (v > 0.0) T.omega = v; }
  PcgLink L;
  L.upd_partials = h->fpartials.p; L.op_partials = h->fpartials.p + (size_t)4 * pstride;
  L.scal = h->scal.p; L.flags = h->flags.p;
  L.n_upd = n_agg; L.n_op = ebe_grid(h, 1, h->n_nodes); L.pstride = pstride;
  L.it = 0; L.max_iter = o.max_iter; L.rtol = o.rtol;
  tl_init_kernel<<<n_agg, kTlThreads, 0, h->stream>>>(T, d_b, h->x.p, h->r.p, h->p.p, h->q.p, L);
  auto coarse_z = [&](int wr_buf) {
    tl_coarse_z_kernel<<<n_agg, kTlThreads, 0, h->stream>>>(T, h->Dinv.p, h->r.p, h->z.p, wr_buf, L);
  };
  coarse_z(0);
  h->launches += 2;
  FEMB_CUDA(h, cudaGetLastError());
  struct Peek { int32_t flags[Flag::COUNT]; double scal[Scal::COUNT]; };
  Peek* peek = reinterpret_cast<Peek*>(h->pinned);
  const int check = o.check_every > 0 ? o.check_every : 50;
  const bool prof = o.profile != 0;
  std::vector<cudaEvent_t> evs;
  int spmv_launches = 0, it = 0, done = 0;
  while (!done && it < o.max_iter) {
    const int batch = std::min(check, o.max_iter - it);
    for (int k = 0; k < batch; ++k, ++it) {
      const bool timed = prof && (it % o.profile) == 0;
      cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
      if (timed) {
        if (h->ev_pool.size() < evs.size() + 3) {
          const size_t old = h->ev_pool.size();
          h->ev_pool.resize(old + 768);
          for (size_t e = old; e < h->ev_pool.size(); ++e) cudaEventCreate(&h->ev_pool[e]);
        }
        e0 = h->ev_pool[evs.size()]; e1 = h->ev_pool[evs.size() + 1]; e2 = h->ev_pool[evs.size() + 2];
        cudaEventRecord(e0, h->stream);
      }
      L.it = it;
      rc = launch_ebe(h, h->z.p, h->s.p, 1, true, nullptr, nullptr, nullptr, nullptr, &L);
      if (timed) cudaEventRecord(e1, h->stream);
      if (rc) return rc;
      ++spmv_launches;
      tl_update_kernel<<<n_agg, kTlThreads, 0, h->stream>>>(T, h->z.p, h->s.p, h->p.p, h->q.p, h->x.p, h->r.p, L);
      coarse_z((it & 1) ^ 1);
      if (timed) { cudaEventRecord(e2, h->stream); evs.push_back(e0); evs.push_back(e1); evs.push_back(e2); }
      h->launches += 2;
    }
    L.it = it;
    tl_decide_kernel<<<1, 128, 0, h->stream>>>(L);
    h->launches++;
    FEMB_CUDA(h, cudaGetLastError());
    FEMB_CUDA(h, cudaMemcpyAsync(peek->flags, h->flags.p, sizeof(peek->flags), cudaMemcpyDeviceToHost, h->stream));
    FEMB_CUDA(h, cudaMemcpyAsync(peek->scal, h->scal.p, sizeof(peek->scal), cudaMemcpyDeviceToHost, h->stream));
    FEMB_CUDA(h, cudaStreamSynchronize(h->stream));
    done = peek->flags[Flag::DONE];
  }
  if (st) {
    st->method_used = FEMB_SOLVER_PCG;
    st->op_used = FEMB_OP_EBE;
    st->coarse_dim = (int32_t)h->coarse_n;
    st->precond_used = FEMB_PRECOND_TWO_LEVEL;
    st->iterations = peek->flags[Flag::ITERS];
    st->converged = (done == 1);
    st->spmv_launches = spmv_launches;
    const double bb = peek->scal[Scal::BB];
    st->rel_residual = bb > 0.0 ? sqrt(peek->scal[Scal::RR] / bb) : 0.0;
    st->spmv_ms = 0.0;
    st->update_ms = 0.0;
    for (size_t i = 0; i + 2 < evs.size(); i += 3) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, evs[i], evs[i + 1]);
      st->spmv_ms += ms;
      cudaEventElapsedTime(&ms, evs[i + 1], evs[i + 2]);
      st->update_ms += ms;
    }
    st->spmv_timed = (int32_t)(evs.size() / 3);
  }
  if (done == 2) return fail(h, FEMB_ERR_SINGULAR, "PCG breakdown: p^T K p <= 0 (K_ff is not positive definite — unconstrained rigid-body motion or zero section properties?)");
  if (done != 1) return fail(h, FEMB_ERR_NOT_CONVERGED, "PCG did not reach rtol within max_iter");
  return FEMB_OK;
}

}  // namespace femb
