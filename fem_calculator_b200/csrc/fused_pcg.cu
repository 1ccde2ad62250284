// One kernel per CG iteration for frames solved with the matrix-free operator.
//
// The two-kernel Chronopoulos-Gear iteration (solver.cu) is
//     (1) s = A z, delta = (z, s)                 operator: FP64 / gather-latency bound
//     (2) p, q, x, r, z, gamma, ||r||^2           vector update: streaming, memory bound
// and each kernel pays a launch gap plus a grid-reduction tail (~7 us of a 46 us iteration at 1M
// DOF).  With the matrix-free operator the product at a node only needs z at the node and at its
// element neighbours — and z_j' = D_j (r_j - alpha (s_j + beta q_j)) can be RECOMPUTED from the old
// r, q, s of the neighbour, because alpha and beta are known when the kernel starts.  So update(it)
// and operator(it+1) become one kernel:
//     own node i : p_i = D_i r_i + beta p_i, q_i' = s_i + beta q_i, x_i += alpha p_i,
//                  r_i' = r_i - alpha q_i', z_i' = D_i r_i'             (p, x in place)
//     neighbour j: q_j', r_j', z_j' by the same instruction sequence (bit-identical to what j's own
//                  lanes compute), from the OLD r, q, s -> these three vectors are ping-ponged
//     s_i' = sum over the node's element ends of K_e[a][a] z_i' + K_e[a][1-a] z_j'   (ebe.cuh)
//     partial sums of delta' = (z', s'), gamma' = (r', z'), ||r'||^2
// z is never stored.  The scalars are not finished by a "last CTA": every CTA publishes its three
// partial sums, and every CTA of the NEXT kernel adds all partials in the same fixed order
// (bit-identical in all CTAs, no atomics, no ticket, no serial tail).  Memory and FP64 work overlap
// inside one kernel instead of alternating between two.
//
// Vectors per iteration: 6 read + 5 written (own node) — the gathers hit L1/L2.
// Reproducible run to run; preconditioner: scalar Jacobi (or none).
//
// MEASURED (1M-DOF frame, one B200, profiles/r01_pcg_iteration_experiments.log): 52.0 us per iteration against 45.1 us
// for operator + update as two kernels — same iteration count, same answer.  The four-fold gather
// volume per neighbour (D, r, q, s instead of z) costs more than the saved launch gap and reduction
// tail, so this path is opt-in (FEMB_OP_EBE_FUSED / FEMB_FUSED_PCG=1) and the two-kernel iteration
// stays the default.
#include <algorithm>
#include <cmath>

#include "common.cuh"
#include "ebe.cuh"
#include "pcg_common.cuh"

namespace femb {

struct FusedVecs {
  const double* b;        // masked right-hand side (read by the INIT launch only)
  const double* dinv;     // scalar Jacobi (ndof)
  double* p;
  double* x;
  double* r[2];
  double* q[2];
  double* s[2];
};

// q' = s + beta q, r' = r - alpha q', z' = d r'   — one fixed instruction sequence for own node and
// neighbours alike, so both sides of an element see the same z'
__device__ __forceinline__ void cg_advance(double d, double r, double q, double s, double alpha, double beta,
                                           double& qn, double& rn, double& zn) {
  qn = __fma_rn(beta, q, s);
  rn = __fma_rn(-alpha, qn, r);
  zn = __dmul_rn(d, rn);
}

__device__ __forceinline__ void load6s(const double* __restrict__ v, int node, double* u) {   // plain (coherent) loads
  const double2* p = reinterpret_cast<const double2*>(v + (size_t)node * 6);
  const double2 a = p[0], b = p[1], c = p[2];
  u[0] = a.x; u[1] = a.y; u[2] = b.x; u[3] = b.y; u[4] = c.x; u[5] = c.y;
}
__device__ __forceinline__ void store6(double* __restrict__ v, int node, const double* u) {
  double2* p = reinterpret_cast<double2*>(v + (size_t)node * 6);
  p[0] = make_double2(u[0], u[1]); p[1] = make_double2(u[2], u[3]); p[2] = make_double2(u[4], u[5]);
}

// it = -1: INIT (x = 0, r = b, p = q = 0, z = D b, s = A z); it >= 0: update(it) + operator(it+1).
// T lanes share a node (pairs part, part+T, ...).  partials: [2 buffers][3 values][pstride].
template <int T, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
frame_fused_pcg_kernel(const FrameParams P, const int4* __restrict__ pair_rec, const int4* __restrict__ node_rec,
                       int n_nodes, const uint8_t* __restrict__ free_mask, const FusedVecs V, int it, int max_iter,
                       double rtol, double* partials, int pstride, double* scal, int* flags) {
  constexpr int NPC = THREADS / T;
  static_assert(32 % T == 0 && THREADS % 32 == 0, "a node's lanes share a warp");
  __shared__ double s_part[3 * THREADS / 32];
  if (flags[Flag::DONE]) return;
  const int part = threadIdx.x % T;
  double alpha = 0.0, beta = 0.0;
  if (it >= 0) {
    // ---- finish the previous kernel's reductions: every CTA adds all partials in the same order
    const double* pb = partials + (size_t)(it & 1) * 3 * pstride;
    double tot[3] = {0.0, 0.0, 0.0};
    for (int i = threadIdx.x; i < (int)gridDim.x; i += THREADS) {
      tot[0] += __ldcg(pb + i); tot[1] += __ldcg(pb + pstride + i); tot[2] += __ldcg(pb + 2 * pstride + i);
    }
    block_sum_all<THREADS, 3>(tot, s_part);
    const double delta = tot[0], gamma = tot[1], rr = tot[2];
    int done = 0;
    double tol2;
    if (it == 0) {
      tol2 = rtol * rtol * rr;                 // rr of the INIT launch is ||b||^2
      if (rr == 0.0) done = 1;                 // zero load: u = 0 is the answer
    } else {
      tol2 = scal[Scal::TOL2];
      if (rr <= tol2) done = 1;
      else if (it >= max_iter) done = 3;
    }
    const int rd = it & 1, wr = rd ^ 1;
    if (!done) {
      beta = (it == 0) ? 0.0 : gamma / scal[Scal::RZ0 + rd];
      const double den = (it == 0) ? delta : delta - beta * gamma / scal[Scal::ALPHA + rd];
      if (!(den > 0.0)) done = 2;              // K_ff not positive definite along p
      else alpha = gamma / den;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      scal[Scal::RZ0 + wr] = gamma;
      scal[Scal::ALPHA + wr] = alpha;
      scal[Scal::RR] = rr;
      if (it == 0) { scal[Scal::BB] = rr; scal[Scal::TOL2] = tol2; }
      flags[Flag::ITERS] = it;                 // updates completed so far
      if (done) flags[Flag::DONE] = done;
    }
    if (done) return;
  }
  const int rd = (it < 0) ? 0 : (it & 1), wr = (it < 0) ? 0 : (rd ^ 1);
  const double* __restrict__ r_old = V.r[rd];
  const double* __restrict__ q_old = V.q[rd];
  const double* __restrict__ s_old = V.s[rd];
  double* __restrict__ r_new = V.r[wr];
  double* __restrict__ q_new = V.q[wr];
  double* __restrict__ s_new = V.s[wr];
  double sums[3] = {0.0, 0.0, 0.0};            // delta', gamma', rr'
  const int per = (n_nodes + (int)gridDim.x - 1) / (int)gridDim.x;   // contiguous, equally sized node ranges
  const int lo = blockIdx.x * per;
  const int hi = min(n_nodes, lo + per);
  for (int base = lo; base < hi; base += NPC) {
    const int node = base + threadIdx.x / T;
    const bool active = node < hi;
    int first = 0, count = 0;
    double px = 0.0, py = 0.0, pz = 0.0;
    double zi[6], ri[6], acc[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) { zi[c] = 0.0; ri[c] = 0.0; acc[c] = 0.0; }
    if (active) {
      const int4 nr = __ldg(node_rec + node);
      first = nr.x; count = nr.y;
      const double* pp = P.xyz + 3 * (size_t)node;
      px = __ldg(pp); py = __ldg(pp + 1); pz = __ldg(pp + 2);
      double d[6];
      load6(V.dinv, node, d);
      if (it < 0) {
        double bi[6];
        load6(V.b, node, bi);
#pragma unroll
        for (int c = 0; c < 6; ++c) { ri[c] = bi[c]; zi[c] = __dmul_rn(d[c], bi[c]); }
        if (part == 0) {
          const double zero6[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
          store6(V.x, node, zero6); store6(V.p, node, zero6); store6(q_new, node, zero6); store6(r_new, node, ri);
        }
      } else {
        double r6[6], q6[6], s6[6], p6[6], x6[6], qn[6];
        load6s(r_old, node, r6); load6s(q_old, node, q6); load6s(s_old, node, s6);
        load6s(V.p, node, p6); load6s(V.x, node, x6);
#pragma unroll
        for (int c = 0; c < 6; ++c) {
          const double zold = __dmul_rn(d[c], r6[c]);
          p6[c] = __fma_rn(beta, p6[c], zold);
          x6[c] = __fma_rn(alpha, p6[c], x6[c]);
          cg_advance(d[c], r6[c], q6[c], s6[c], alpha, beta, qn[c], ri[c], zi[c]);
        }
        if (part == 0) { store6(V.p, node, p6); store6(V.x, node, x6); store6(q_new, node, qn); store6(r_new, node, ri); }
      }
    }
#pragma unroll 1
    for (int j = part; j < count; j += T) {
      const int4 rec = __ldg(pair_rec + first + j);
      const int a = (rec.w >> 24) & 1;
      const int other = rec.y;
      const double* po = P.xyz + 3 * (size_t)other;
      const double ox = __ldg(po), oy = __ldg(po + 1), oz = __ldg(po + 2);
      double zo[6];
      {
        double d[6];
        load6(V.dinv, other, d);
        if (it < 0) {
          double bo[6];
          load6(V.b, other, bo);
#pragma unroll
          for (int c = 0; c < 6; ++c) zo[c] = __dmul_rn(d[c], bo[c]);
        } else {
          double r6[6], q6[6], s6[6];
          load6s(r_old, other, r6); load6s(q_old, other, q6); load6s(s_old, other, s6);
#pragma unroll
          for (int c = 0; c < 6; ++c) {
            double qn, rn;
            cg_advance(d[c], r6[c], q6[c], s6[c], alpha, beta, qn, rn, zo[c]);
          }
        }
      }
      const double* sp = P.sec_props + 8 * (size_t)(rec.w & 0xFFFFFF);
      FrameIn in;
      in.dx = a ? px - ox : ox - px; in.dy = a ? py - oy : oy - py; in.dz = a ? pz - oz : oz - pz;
      in.A = __ldg(sp); in.Ix = __ldg(sp + 1); in.Iy = __ldg(sp + 2); in.J = __ldg(sp + 3);
      in.ky = __ldg(sp + 4); in.kz = __ldg(sp + 5);
      KRec k;
      krec_from(P, in, a, k);
      double o6[6];
      ebe_apply(k, zi, zo, o6);
#pragma unroll
      for (int c = 0; c < 6; ++c) acc[c] += o6[c];
    }
    if (T > 1) {
#pragma unroll
      for (int off = 1; off < T; off <<= 1)
#pragma unroll
        for (int c = 0; c < 6; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], off);
    }
    if (active && part == 0) {
      const uint8_t* fm = free_mask + (size_t)node * 6;
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        if (!fm[c]) acc[c] = zi[c];            // identity rows on the fixed DOFs
        sums[0] += zi[c] * acc[c];
        sums[1] += ri[c] * zi[c];
        sums[2] += ri[c] * ri[c];
      }
      store6(s_new, node, acc);
    }
  }
  // ---- publish this CTA's partial sums for the next kernel (buffer (it+1) & 1)
  block_sum_all<THREADS, 3>(sums, s_part);
  if (threadIdx.x == 0) {
    double* pb = partials + (size_t)((it + 1) & 1) * 3 * pstride;
    pb[blockIdx.x] = sums[0]; pb[pstride + blockIdx.x] = sums[1]; pb[2 * pstride + blockIdx.x] = sums[2];
  }
}

bool fused_pcg_applicable(const femb_handle* h, const femb_solve_opts& o) {
  static int on = -1;   // FEMB_FUSED_PCG=1: use it wherever the matrix-free operator is selected
  if (on < 0) { const char* e = getenv("FEMB_FUSED_PCG"); on = (e && e[0] == '1') ? 1 : 0; }
  if (!on && o.op != FEMB_OP_EBE_FUSED) return false;
  return ebe_selected(h, o.op) && o.precond != FEMB_PRECOND_BLOCK_JACOBI && h->bs == 6;
}

// K_ff x = b (masked right-hand side d_b on the device); solution in h->x.  Same contract as pcg_core.
int pcg_fused(femb_handle* h, const femb_solve_opts& o, const double* d_b, femb_stats* st) {
  const int64_t n = h->ndof;
  int rc = setup_precond_public(h, o.precond);
  if (rc) return rc;
  FEMB_CUDA(h, h->r2.ensure((size_t)n));
  FEMB_CUDA(h, h->q2.ensure((size_t)n));
  const int pstride = h->num_sms * 8;
  FEMB_CUDA(h, h->fpartials.ensure((size_t)pstride * 6));
  FEMB_CUDA(h, cudaMemsetAsync(h->flags.p, 0, sizeof(int32_t) * Flag::COUNT, h->stream));
  FEMB_CUDA(h, cudaMemsetAsync(h->scal.p, 0, sizeof(double) * Scal::COUNT, h->stream));
  FusedVecs V;
  V.b = d_b; V.dinv = h->Dinv.p; V.p = h->p.p; V.x = h->x.p;
  V.r[0] = h->r.p; V.r[1] = h->r2.p; V.q[0] = h->q.p; V.q[1] = h->q2.p; V.s[0] = h->s.p; V.s[1] = h->z.p;
  FrameParams P;
  P.xyz = h->xyz.p; P.conn = h->conn.p; P.elem_sec = h->elem_sec.p; P.sec_props = h->sec_props.p;
  P.E = h->E; P.G = h->G; P.rho = h->rho;
  constexpr int T = 2, THREADS = 128, MINB = 4;
  const int n_nodes = (int)h->n_nodes;
  const int need = (n_nodes + THREADS / T - 1) / (THREADS / T);
  const int grid = std::max(1, std::min(need, h->num_sms * MINB));
  const int4* pr = reinterpret_cast<const int4*>(h->pair_rec.p);
  const int4* nr = reinterpret_cast<const int4*>(h->pair_node_rec.p);
  auto launch = [&](int it) {
    frame_fused_pcg_kernel<T, THREADS, MINB><<<grid, THREADS, 0, h->stream>>>(
        P, pr, nr, n_nodes, h->free_mask.p, V, it, o.max_iter, o.rtol, h->fpartials.p, pstride, h->scal.p, h->flags.p);
    h->launches++;
  };
  struct Peek { int32_t flags[Flag::COUNT]; double scal[Scal::COUNT]; };
  Peek* peek = reinterpret_cast<Peek*>(h->pinned);
  const int check = o.check_every > 0 ? o.check_every : 50;
  const bool prof = o.profile != 0;
  std::vector<cudaEvent_t> evs;
  launch(-1);
  int it = 0, done = 0, launches = 1;
  while (!done && it <= o.max_iter) {
    const int batch = std::min(check, o.max_iter + 1 - it);
    for (int k = 0; k < batch; ++k, ++it) {
      const bool timed = prof && (it % o.profile) == 0;
      if (timed) {
        if (h->ev_pool.size() < evs.size() + 2) {
          const size_t old = h->ev_pool.size();
          h->ev_pool.resize(old + 768);
          for (size_t e = old; e < h->ev_pool.size(); ++e) cudaEventCreate(&h->ev_pool[e]);
        }
        cudaEventRecord(h->ev_pool[evs.size()], h->stream);
      }
      launch(it);
      ++launches;
      if (timed) {
        cudaEventRecord(h->ev_pool[evs.size() + 1], h->stream);
        evs.push_back(h->ev_pool[evs.size()]);
        evs.push_back(h->ev_pool[evs.size()]);
      }
    }
    FEMB_CUDA(h, cudaGetLastError());
    FEMB_CUDA(h, cudaMemcpyAsync(peek->flags, h->flags.p, sizeof(peek->flags), cudaMemcpyDeviceToHost, h->stream));
    FEMB_CUDA(h, cudaMemcpyAsync(peek->scal, h->scal.p, sizeof(peek->scal), cudaMemcpyDeviceToHost, h->stream));
    FEMB_CUDA(h, cudaStreamSynchronize(h->stream));
    done = peek->flags[Flag::DONE];
  }
  if (st) {
    st->method_used = FEMB_SOLVER_PCG;
    st->op_used = FEMB_OP_EBE;
    st->iterations = peek->flags[Flag::ITERS];
    st->converged = (done == 1);
    st->spmv_launches = launches;
    const double bb = peek->scal[Scal::BB];
    st->rel_residual = bb > 0.0 ? std::sqrt(peek->scal[Scal::RR] / bb) : 0.0;
    st->spmv_ms = 0.0;
    st->update_ms = 0.0;
    for (size_t i = 0; i + 1 < evs.size(); i += 2) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, evs[i], evs[i + 1]);
      st->spmv_ms += ms;              // whole fused iteration kernel
    }
    st->spmv_timed = (int32_t)(evs.size() / 2);
  }
  if (done == 2) return fail(h, FEMB_ERR_SINGULAR, "PCG breakdown: p^T K p <= 0 (K_ff is not positive definite — unconstrained rigid-body motion or zero section properties?)");
  if (done != 1) return fail(h, FEMB_ERR_NOT_CONVERGED, "PCG did not reach rtol within max_iter");
  return FEMB_OK;
}

}  // namespace femb
