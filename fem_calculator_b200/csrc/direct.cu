// Direct solvers for the cases CG cannot serve:
//   * chains (path graphs): block-tridiagonal factorisation with 6x6 blocks — beam chains
//     have cond(K) ~ (L/h)^4, hopeless for Jacobi-PCG (SURVEY §7 "hard parts");
//   * batched independent chain models (BASELINE config 4): the same recurrence with the
//     element blocks generated on the fly from the section record (fused element + solve),
//     one thread per model, model-interleaved scratch so every access is coalesced;
//   * small systems: dense FP64 Cholesky of the masked operator.
// All replace np.linalg.solve(k_ff, f_f) (BeamSolver.py:417) / spsolve (ReactionSolver.py:201).
#include <algorithm>

#include <cstdlib>

#include "common.cuh"
#include "elements.cuh"

namespace femb {

// ---- 6x6 helpers (row-major, fully unrolled so everything stays in registers) ----------
// Gauss-Jordan inverse of an SPD 6x6 (no pivoting).  Returns false on a non-positive pivot.
__device__ __forceinline__ bool inv6_spd(const double* S, double* G) {
  double A[36];
#pragma unroll
  for (int i = 0; i < 36; ++i) { A[i] = S[i]; G[i] = (i % 7 == 0) ? 1.0 : 0.0; }
  bool ok = true;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const double piv = A[k * 6 + k];
    if (!(piv > 0.0)) ok = false;
    const double ip = 1.0 / piv;
#pragma unroll
    for (int c = 0; c < 6; ++c) { A[k * 6 + c] *= ip; G[k * 6 + c] *= ip; }
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      if (r == k) continue;
      const double m = A[r * 6 + k];
#pragma unroll
      for (int c = 0; c < 6; ++c) { A[r * 6 + c] -= m * A[k * 6 + c]; G[r * 6 + c] -= m * G[k * 6 + c]; }
    }
  }
  // symmetrise (the exact inverse is symmetric)
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int c = r + 1; c < 6; ++c) {
      const double v = 0.5 * (G[r * 6 + c] + G[c * 6 + r]);
      G[r * 6 + c] = v; G[c * 6 + r] = v;
    }
  return ok;
}

// apply the DOF mask of one node pair to a block: rows masked by mr, cols by mc (bit c set = free)
__device__ __forceinline__ void mask_block(double* B, unsigned mr, unsigned mc, bool diag) {
#pragma unroll
  for (int r = 0; r < 6; ++r)
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      const bool fr = ((mr >> r) & 1u) && ((mc >> c) & 1u);
      if (!fr) B[r * 6 + c] = (diag && r == c) ? 1.0 : 0.0;
    }
}

__device__ __forceinline__ unsigned node_mask(const uint8_t* free_mask, int64_t node) {
  unsigned m = 0;
#pragma unroll
  for (int c = 0; c < 6; ++c) m |= (free_mask[node * 6 + c] ? 1u : 0u) << c;
  return m;
}

// ---- single-model chain solver on the assembled BSR matrix -------------------------------
// factor: G_k = S_k^-1 and W_k = O_k G_k for every chain position (sequential recurrence, one
// thread); apply: one thread per right-hand side walks forward (y) and backward (x), reading
// the shared factors as warp-broadcast loads.  The factors persist in the handle so the modal
// solver's shift-invert steps only pay for the sweeps.
__global__ void chain_factor_kernel(const int32_t* __restrict__ order, const int32_t* __restrict__ rowptr,
                                    const int32_t* __restrict__ colidx, const int32_t* __restrict__ diag_blk,
                                    const double* __restrict__ vals, const uint8_t* __restrict__ free_mask,
                                    double* __restrict__ Gs, double* __restrict__ W, int64_t n, int* status) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double S[36], G[36], O[36], Wk[36];
  bool ok = true;
  {
    const int64_t i0 = order[0];
    const unsigned m0 = node_mask(free_mask, i0);
    const double* d = vals + (size_t)diag_blk[i0] * 36;
    for (int q = 0; q < 36; ++q) S[q] = d[q];
    mask_block(S, m0, m0, true);
  }
  for (int64_t k = 0; k < n; ++k) {
    const int64_t i = order[k];
    ok = inv6_spd(S, G) && ok;
    for (int q = 0; q < 36; ++q) Gs[k * 36 + q] = G[q];
    if (k + 1 == n) break;
    const int64_t j = order[k + 1];
    int blk = -1;  // O = K[j][i] (zero if j is not a neighbour: start of another path)
    for (int bb = rowptr[j]; bb < rowptr[j + 1]; ++bb)
      if (colidx[bb] == (int32_t)i) blk = bb;
    const unsigned mi = node_mask(free_mask, i), mj = node_mask(free_mask, j);
    if (blk >= 0) {
      const double* o = vals + (size_t)blk * 36;
      for (int q = 0; q < 36; ++q) O[q] = o[q];
      mask_block(O, mj, mi, false);
    } else {
      for (int q = 0; q < 36; ++q) O[q] = 0.0;
    }
    for (int r = 0; r < 6; ++r)
      for (int c = 0; c < 6; ++c) {
        double sacc = 0.0;
        for (int t = 0; t < 6; ++t) sacc += O[r * 6 + t] * G[t * 6 + c];
        Wk[r * 6 + c] = sacc;
      }
    for (int q = 0; q < 36; ++q) W[k * 36 + q] = Wk[q];
    const double* d = vals + (size_t)diag_blk[j] * 36;
    for (int q = 0; q < 36; ++q) S[q] = d[q];
    mask_block(S, mj, mj, true);
    for (int r = 0; r < 6; ++r)
      for (int c = 0; c < 6; ++c) {
        double t = 0.0;
        for (int q = 0; q < 6; ++q) t += Wk[r * 6 + q] * O[c * 6 + q];
        S[r * 6 + c] -= t;
      }
  }
  *status = ok ? 0 : 1;
}

// b, x: nrhs vectors with leading dimension ld (x may alias nothing; z is staged in x)
__global__ void chain_apply_kernel(const int32_t* __restrict__ order, const double* __restrict__ Gs,
                                   const double* __restrict__ W, const double* __restrict__ b,
                                   double* __restrict__ x, int64_t n, int nrhs, int64_t ld) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nrhs) return;
  const double* bq = b + (size_t)q * ld;
  double* xq = x + (size_t)q * ld;
  double y[6], z[6];
  {
    const int64_t i0 = order[0];
    for (int c = 0; c < 6; ++c) y[c] = bq[i0 * 6 + c];
  }
  for (int64_t k = 0; k < n; ++k) {
    const int64_t i = order[k];
    const double* G = Gs + k * 36;
    for (int r = 0; r < 6; ++r) {
      double sacc = 0.0;
      for (int c = 0; c < 6; ++c) sacc += G[r * 6 + c] * y[c];
      z[r] = sacc;
    }
    for (int c = 0; c < 6; ++c) xq[i * 6 + c] = z[c];   // z_k staged in x
    if (k + 1 == n) break;
    const int64_t j = order[k + 1];
    const double* Wk = W + k * 36;
    double yn[6];
    for (int r = 0; r < 6; ++r) {
      double sacc = bq[j * 6 + r];
      for (int c = 0; c < 6; ++c) sacc -= Wk[r * 6 + c] * y[c];
      yn[r] = sacc;
    }
    for (int c = 0; c < 6; ++c) y[c] = yn[c];
  }
  double xn[6];
  for (int c = 0; c < 6; ++c) xn[c] = z[c];              // x_{n-1} = z_{n-1}
  for (int64_t k = n - 2; k >= 0; --k) {
    const int64_t i = order[k];
    const double* Wk = W + k * 36;
    double xk[6];
    for (int r = 0; r < 6; ++r) {
      double sacc = xq[i * 6 + r];
      for (int c = 0; c < 6; ++c) sacc -= Wk[c * 6 + r] * xn[c];
      xk[r] = sacc;
    }
    for (int c = 0; c < 6; ++c) { xn[c] = xk[c]; xq[i * 6 + c] = xk[c]; }
  }
}

// defined in solver.cu
int setup_rhs_for_direct(femb_handle* h);

int chain_factor(femb_handle* h) {
  if (h->kind != Kind::Frame || !h->sym.is_chain) return fail(h, FEMB_ERR_ARG, "mesh is not a chain (path graph)");
  if (h->chain_factored) return FEMB_OK;
  DevBuf<int> status;
  FEMB_CUDA(h, upload(h->chain_order, h->sym.chain_order, h->stream));
  FEMB_CUDA(h, h->chainW.alloc((size_t)h->n_nodes * 36));
  FEMB_CUDA(h, h->chainG.alloc((size_t)h->n_nodes * 36));
  FEMB_CUDA(h, status.alloc(1));
  chain_factor_kernel<<<1, 32, 0, h->stream>>>(h->chain_order.p, h->rowptr.p, h->colidx.p, h->diag_blk.p, h->Kvals.p,
                                               h->free_mask.p, h->chainG.p, h->chainW.p, h->n_nodes, status.p);
  h->launches++;
  FEMB_CUDA(h, cudaGetLastError());
  int* hs = reinterpret_cast<int*>(h->pinned);
  FEMB_CUDA(h, cudaMemcpyAsync(hs, status.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  FEMB_CUDA(h, cudaStreamSynchronize(h->stream));
  if (*hs != 0) return fail(h, FEMB_ERR_SINGULAR, "chain factorisation hit a non-positive pivot (K_ff not positive definite)");
  h->chain_factored = true;
  return FEMB_OK;
}

int chain_apply(femb_handle* h, const double* d_b, double* d_x, int nrhs, int64_t ld) {
  chain_apply_kernel<<<(nrhs + 31) / 32, 32, 0, h->stream>>>(h->chain_order.p, h->chainG.p, h->chainW.p, d_b, d_x,
                                                             h->n_nodes, nrhs, ld);
  h->launches++;
  FEMB_CUDA(h, cudaGetLastError());
  return FEMB_OK;
}

int run_chain_solve(femb_handle* h, femb_stats* st) {
  int rc = setup_rhs_for_direct(h);
  if (rc) return rc;
  rc = chain_factor(h);
  if (rc) return rc;
  rc = chain_apply(h, h->b.p, h->x.p, 1, h->ndof);
  if (rc) return rc;
  if (st) { st->method_used = FEMB_SOLVER_CHAIN; st->iterations = 1; st->converged = 1; st->rel_residual = 0.0; }
  return FEMB_OK;
}

// ---- batched chain models: fused element generation + block-tridiagonal solve ------------
// BASELINE config 4: thousands of independent chain models.  SIX lanes share a model — lane r owns ROW r of every
// 6x6 block of the recurrence (S_k, G_k = S_k^-1, O_k, W_k = O_k G_k) and entry r of the 6-vectors — and five models
// share a warp.  The first version gave each model ONE thread: 8,192 threads on 148 SMs walking a 2,000-step
// dependent recurrence of ~700 FP64 operations per step, 10.2 ms whether 1,024 or 8,192 models were solved.  With
// the rows spread over lanes a step is ~130 FP64 operations per lane plus ~120 shuffles, six times the threads hide
// the latency, and the time follows the model count.  Scratch (W_k, z_k for the back substitution) is model-major:
// a lane writes / reads its 48-byte row, a model's 336 bytes per step are contiguous.
struct BatchParams {
  const double* xyz;        // (n_nodes,3) shared
  const double* sec_props;  // (n_models,8)
  const uint8_t* fixed;     // (ndof) 1 = fixed
  const double* f;          // (n_models, ndof)
  double* u;                // (n_models, ndof)
  double* W;                // scratch (n_models, n_nodes, 36)
  double* z;                // scratch (n_models, n_nodes, 6)
  int* status;              // count of models with a non-positive pivot
  int64_t n_models, n_elem;
  double E, G;
};

// row `row` of block [a][b] of R^T k R (the expressions of frame_kblock, one row, no dynamic indexing)
__device__ __forceinline__ void frame_krow(const FrameRec& R, int a, int b, int row, double* out) {
  const bool same = (a == b);
  const double suu = same ? 1.0 : -1.0;
  const double sa = (a == 0) ? 1.0 : -1.0, sb = (b == 0) ? 1.0 : -1.0;
  const int rr = row < 3 ? row : row - 3;
  const double tr = rr == 0 ? R.t[0] : (rr == 1 ? R.t[1] : R.t[2]);
  const double n1r = rr == 0 ? R.n1[0] : (rr == 1 ? R.n1[1] : R.n1[2]);
  const double n2r = rr == 0 ? R.n2[0] : (rr == 1 ? R.n2[1] : R.n2[2]);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double tt = tr * R.t[c], n11 = n1r * R.n1[c], n22 = n2r * R.n2[c], n12 = n1r * R.n2[c], n21 = n2r * R.n1[c];
    if (row < 3) {
      out[c] = suu * (R.ax * tt + R.k11z * n11 + R.k11y * n22);
      out[3 + c] = sa * (R.k12z * n12 - R.k12y * n21);
    } else {
      out[c] = sb * (R.k12z * n21 - R.k12y * n12);
      out[3 + c] = (same ? R.tor : -R.tor) * tt + (same ? R.k22y : R.k23y) * n11 + (same ? R.k22z : R.k23z) * n22;
    }
  }
}

constexpr int kBcThreads = 192;       // 6 warps x 5 models

__global__ void __launch_bounds__(kBcThreads)
batch_chain_rows_kernel(BatchParams B) {
  constexpr unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grp = lane / 6, r = lane - grp * 6;
  const int64_t m_raw = ((int64_t)blockIdx.x * (kBcThreads / 32) + warp) * 5 + grp;
  const bool active = grp < 5 && m_raw < B.n_models;
  const int64_t m = active ? m_raw : 0;               // idle lanes shadow model 0 (they store nothing)
  const int base = (grp < 5 ? grp : 4) * 6;           // first lane of the group the shuffles read from
  const int64_t nn = B.n_elem + 1, ndof = nn * 6;
  int32_t conn2[2] = {0, 1};
  int32_t sec0 = 0;
  FrameParams P;
  P.conn = conn2; P.elem_sec = &sec0; P.sec_props = B.sec_props + 8 * m;
  P.E = B.E; P.G = B.G; P.rho = 0.0;
  const double* f = B.f + m * ndof;
  double* Wm = B.W + (size_t)m * nn * 36;
  double* zm = B.z + (size_t)m * nn * 6;
  auto freemask = [&](int64_t node) {
    unsigned mk = 0;
#pragma unroll
    for (int c = 0; c < 6; ++c) mk |= (B.fixed[node * 6 + c] ? 0u : 1u) << c;
    return mk;
  };
  bool ok = true;
  FrameRec R;
  P.xyz = B.xyz;
  frame_record(P, 0, R);
  double a[6], y;
  unsigned mi = freemask(0);
  frame_krow(R, 0, 0, r, a);                           // D_0 = K_e0[0][0]
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    const bool fr = ((mi >> r) & 1u) && ((mi >> c) & 1u);
    if (!fr) a[c] = (c == r) ? 1.0 : 0.0;
  }
  y = ((mi >> r) & 1u) ? f[r] : 0.0;
  for (int64_t k = 0; k < nn; ++k) {
    // G = S^-1: Gauss-Jordan, rows spread over the six lanes
    double g[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) g[c] = (c == r) ? 1.0 : 0.0;
#pragma unroll
    for (int p = 0; p < 6; ++p) {
      double ap[6], gp[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        ap[c] = (c >= p) ? __shfl_sync(FULL, a[c], base + p) : 0.0;
        gp[c] = (c <= p) ? __shfl_sync(FULL, g[c], base + p) : 0.0;
      }
      const double piv = ap[p];
      if (!(piv > 0.0)) ok = false;
      const double ip = 1.0 / piv;
      const double fct = (r == p) ? 0.0 : a[p] * ip;
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        if (c >= p) a[c] = (r == p) ? ap[c] * ip : a[c] - fct * ap[c];
        if (c <= p) g[c] = (r == p) ? gp[c] * ip : g[c] - fct * gp[c];
      }
    }
    double yc[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) yc[c] = __shfl_sync(FULL, y, base + c);
    double zk = 0.0;
#pragma unroll
    for (int c = 0; c < 6; ++c) zk = fma(g[c], yc[c], zk);
    if (active) zm[k * 6 + r] = zk;
    if (k + 1 == nn) break;
    // element k joins nodes k, k + 1 (R holds it): O = K[k+1][k], W = O G
    const unsigned mj = freemask(k + 1);
    double o[6];
    frame_krow(R, 1, 0, r, o);
#pragma unroll
    for (int c = 0; c < 6; ++c)
      if (!(((mj >> r) & 1u) && ((mi >> c) & 1u))) o[c] = 0.0;
    double w[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int t = 0; t < 6; ++t) {
#pragma unroll
      for (int c = 0; c < 6; ++c) w[c] = fma(o[t], __shfl_sync(FULL, g[c], base + t), w[c]);
    }
    if (active) {
      double2* wp = reinterpret_cast<double2*>(Wm + k * 36 + r * 6);
      wp[0] = make_double2(w[0], w[1]); wp[1] = make_double2(w[2], w[3]); wp[2] = make_double2(w[4], w[5]);
    }
    // S_{k+1} = D_{k+1} - W O^T,  D_{k+1} = K_e(k)[1][1] + K_e(k+1)[0][0]
    frame_krow(R, 1, 1, r, a);
    double s_sub[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int c = 0; c < 6; ++c) {
#pragma unroll
      for (int q = 0; q < 6; ++q) s_sub[c] = fma(w[q], __shfl_sync(FULL, o[q], base + c), s_sub[c]);
    }
    if (k + 1 < B.n_elem) {
      P.xyz = B.xyz + 3 * (k + 1);
      frame_record(P, 0, R);
      double d2[6];
      frame_krow(R, 0, 0, r, d2);
#pragma unroll
      for (int c = 0; c < 6; ++c) a[c] += d2[c];
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      const bool fr = ((mj >> r) & 1u) && ((mj >> c) & 1u);
      a[c] = fr ? a[c] - s_sub[c] : ((c == r) ? 1.0 : 0.0);
    }
    double yn = ((mj >> r) & 1u) ? f[(k + 1) * 6 + r] : 0.0;
#pragma unroll
    for (int c = 0; c < 6; ++c) yn = fma(-w[c], yc[c], yn);
    y = yn;
    mi = mj;
  }
  // back substitution: x_k = z_k - W_k^T x_{k+1}; lane r reads column r of W_k
  double* u = B.u + m * ndof;
  double xn = active ? zm[(nn - 1) * 6 + r] : 0.0;
  if (active) u[(nn - 1) * 6 + r] = xn;
#pragma unroll 4
  for (int64_t k = nn - 2; k >= 0; --k) {
    double wc[6], xk = 0.0;
    if (active) {
      xk = zm[k * 6 + r];
#pragma unroll
      for (int c = 0; c < 6; ++c) wc[c] = Wm[k * 36 + c * 6 + r];
    } else {
#pragma unroll
      for (int c = 0; c < 6; ++c) wc[c] = 0.0;
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) xk = fma(-wc[c], __shfl_sync(FULL, xn, base + c), xk);
    xn = xk;
    if (active) u[k * 6 + r] = xk;
  }
  if (!ok && active && r == 0) atomicAdd(B.status, 1);
}

int run_batch_chain(femb_handle* h, int64_t n_models, int64_t n_elem, const double* xyz,
                    const double* sec_props, double E, double G, const uint8_t* fixed_mask,
                    const double* f, double* u, femb_stats* st) {
  const int64_t nn = n_elem + 1, ndof = nn * 6;
  // (f and u are hundreds of MB: a caller that page-locked them with femb_host_register gets PCIe-speed async copies,
  // any other buffer goes through the driver's pageable staging — same result)
  DevBuf<double> dxyz, dsec;
  DevBuf<uint8_t> dfix;
  DevBuf<int> status;
  FEMB_CUDA(h, upload(dxyz, xyz, (size_t)nn * 3, h->stream));
  FEMB_CUDA(h, upload(dsec, sec_props, (size_t)n_models * 8, h->stream));
  FEMB_CUDA(h, upload(dfix, fixed_mask, (size_t)ndof, h->stream));
  // the big device buffers persist in the handle (loads, solutions, 336 B of scratch per node and model)
  FEMB_CUDA(h, h->batch_f.ensure((size_t)n_models * ndof));
  FEMB_CUDA(h, h->batch_u.ensure((size_t)n_models * ndof));
  FEMB_CUDA(h, h->batch_W.ensure((size_t)n_models * nn * 36));
  FEMB_CUDA(h, h->batch_z.ensure((size_t)n_models * nn * 6));
  g_h2d_bytes += (long long)n_models * ndof * 8;
  FEMB_CUDA(h, cudaMemcpyAsync(h->batch_f.p, f, (size_t)n_models * ndof * 8, cudaMemcpyHostToDevice, h->stream));
  FEMB_CUDA(h, status.alloc(1));
  FEMB_CUDA(h, cudaMemsetAsync(status.p, 0, sizeof(int), h->stream));
  BatchParams B{dxyz.p, dsec.p, dfix.p, h->batch_f.p, h->batch_u.p, h->batch_W.p, h->batch_z.p, status.p, n_models, n_elem, E, G};
  FEMB_CUDA(h, cudaEventRecord(h->ev0, h->stream));
  const int64_t per_cta = (kBcThreads / 32) * 5;
  batch_chain_rows_kernel<<<(unsigned)((n_models + per_cta - 1) / per_cta), kBcThreads, 0, h->stream>>>(B);
  h->launches++;
  FEMB_CUDA(h, cudaGetLastError());
  FEMB_CUDA(h, cudaEventRecord(h->ev1, h->stream));
  int* hs = reinterpret_cast<int*>(h->pinned);
  FEMB_CUDA(h, cudaMemcpyAsync(hs, status.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  if (u) FEMB_CUDA(h, download(u, h->batch_u.p, (size_t)n_models * ndof * 8, h->stream));
  FEMB_CUDA(h, cudaStreamSynchronize(h->stream));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, h->ev0, h->ev1);
  if (st) { st->method_used = FEMB_SOLVER_CHAIN; st->iterations = 1; st->converged = (*hs == 0); st->device_ms = ms; }
  if (*hs != 0) return fail(h, FEMB_ERR_SINGULAR, "batched chain solve: non-positive pivot in at least one model");
  return FEMB_OK;
}

// ---- dense FP64 Cholesky of the masked operator: blocked, trailing update on DMMA -----------------
// For reduced systems small enough to be a real dense contraction (north_star (4); robust where
// cond(K_ff) defeats CG).  A = P K P + (I - P) is expanded to a dense row-major matrix padded to a
// multiple of 64 (identity on the padding) and factored right-looking in 64-column panels:
//   1. chol_potrf_block_lean_kernel  64x64 diagonal block, one CTA, one row per thread in registers
//   2. chol_trsm_kernel          panel rows x L_kk^-T, one thread per row (row in registers)
//   3. chol_syrk_dmma_kernel     trailing A_ij -= P_i P_j^T on the FP64 tensor cores:
//                                mma.sync.m8n8k4.f64 (SASS DMMA), 64x64 tile per CTA, both panel
//                                tiles staged once in shared memory (K = 64 fits), n^3/3 flops
// After the factorisation L^T is mirrored into the upper triangle so both substitution sweeps read
// contiguous rows.  Non-positive pivots are reported, never patched silently.
constexpr int kCB = 64;           // panel width / tile edge
constexpr int kCBLd = kCB + 4;    // shared row stride (doubles): conflict-free DMMA fragment loads
constexpr int64_t kDenseMax = 16384;

__global__ void dense_fill_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                  const double* __restrict__ vals, const uint8_t* __restrict__ mask,
                                  double* __restrict__ A, int64_t n, int64_t ld, int bs) {
  // one CTA per row of the padded matrix: zero the row, then scatter the row of P K P + (I - P)
  const int64_t g = blockIdx.x;
  double* row = A + (size_t)g * ld;
  for (int64_t c = threadIdx.x; c < ld; c += blockDim.x) row[c] = 0.0;
  __syncthreads();
  if (g >= n || !mask[g]) { if (threadIdx.x == 0) row[g] = 1.0; return; }
  const int64_t node = g / bs;
  const int r = (int)(g - node * bs);
  const int b0 = rowptr[node], b1 = rowptr[node + 1];
  for (int t = threadIdx.x; t < (b1 - b0) * bs; t += blockDim.x) {
    const int b = b0 + t / bs, c = t % bs;
    const int64_t col = (int64_t)colidx[b] * bs + c;
    if (mask[col]) row[col] = vals[(size_t)b * bs * bs + r * bs + c];
  }
}

// 64x64 diagonal block, one thread per ROW with the row in registers (fully unrolled, static indexing),
// one barrier per column: at step j every thread i >= j publishes its a_ij, all read the pivot d = a_jj
// and the column, and update a_ik -= (a_ij / d) a_kj.  The update of row i runs unpredicated over
// k = j+1..63 (entries right of the diagonal are never published nor stored), the column is read with
// 16-byte shared loads, and 1/d, 1/sqrt(d), sqrt(d) all come from one rsqrt.  History (ncu launch list of
// the coarse inversion, profiles/r01_ncu_full_two_level_setup.txt): a shared-memory version with eight
// warps and three barriers per column 64 us per panel; a predicated register version 9.5k SASS
// instructions, instruction-fetch bound (stall no_instruction 5.6 per issue); this one 5.8k, no spills.
__global__ void __launch_bounds__(kCB)
chol_potrf_block_lean_kernel(double* __restrict__ A, int64_t ld, int k0, int* status) {
  __shared__ __align__(16) double col[2][kCB];
  const int i = threadIdx.x;
  double a[kCB];
  {
    const double* row = A + (size_t)(k0 + i) * ld + k0;
#pragma unroll
    for (int c = 0; c < kCB; c += 2) { const double2 v = *reinterpret_cast<const double2*>(row + c); a[c] = v.x; a[c + 1] = v.y; }
  }
#pragma unroll
  for (int j = 0; j < kCB; ++j) {
    double* cj = col[j & 1];
    if (i >= j) cj[i] = a[j];
    __syncthreads();
    double d = cj[j];
    if (!(d > 0.0)) { if (i == j) *status = 1; d = 1.0; }
    const double rs = rsqrt(d);               // 1/sqrt(d); d * rs = sqrt(d); rs * rs = 1/d
    const double t = a[j] * (rs * rs);
    if (((j + 1) & 1) && j + 1 < kCB) a[j + 1] -= t * cj[j + 1];          // odd first column: scalar, then pairs
#pragma unroll
    for (int k = (j + 2) & ~1; k < kCB; k += 2) {
      const double2 c2 = *reinterpret_cast<const double2*>(cj + k);
      a[k] -= t * c2.x;
      a[k + 1] -= t * c2.y;
    }
    a[j] = (i == j) ? d * rs : a[j] * rs;
  }
  double* row = A + (size_t)(k0 + i) * ld + k0;
#pragma unroll
  for (int c = 0; c < kCB; ++c)
    if (c <= i) row[c] = a[c];
}

static void launch_potrf_block(femb_handle* h, double* A, int64_t ld, int k0, int* status) {
  chol_potrf_block_lean_kernel<<<1, kCB, 0, h->stream>>>(A, ld, k0, status);
}

// rows below the diagonal block: X L_kk^T = A_panel, one thread per row
__global__ void __launch_bounds__(128)
chol_trsm_kernel(double* __restrict__ A, int64_t ld, int k0, int n_pad) {
  __shared__ double L[kCB][kCB + 1];
  __shared__ double invd[kCB];
  for (int t = threadIdx.x; t < kCB * kCB; t += 128) L[t / kCB][t % kCB] = A[(size_t)(k0 + t / kCB) * ld + k0 + t % kCB];
  __syncthreads();
  if (threadIdx.x < kCB) invd[threadIdx.x] = 1.0 / L[threadIdx.x][threadIdx.x];
  __syncthreads();
  const int row = k0 + kCB + blockIdx.x * 128 + threadIdx.x;
  if (row >= n_pad) return;
  double* a = A + (size_t)row * ld + k0;
  double x[kCB];
#pragma unroll
  for (int c = 0; c < kCB; c += 2) { const double2 v = *reinterpret_cast<const double2*>(a + c); x[c] = v.x; x[c + 1] = v.y; }
#pragma unroll
  for (int c = 0; c < kCB; ++c) {
    double v = x[c];
#pragma unroll
    for (int j = 0; j < c; ++j) v -= x[j] * L[c][j];
    x[c] = v * invd[c];
  }
#pragma unroll
  for (int c = 0; c < kCB; c += 2) *reinterpret_cast<double2*>(a + c) = make_double2(x[c], x[c + 1]);
}

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// trailing update of the lower tiles (ti >= tj): A[ti][tj] -= P_ti P_tj^T, P = the panel just solved
__global__ void __launch_bounds__(128)
chol_syrk_dmma_kernel(double* __restrict__ A, int64_t ld, int k0) {
  const int ti = blockIdx.y, tj = blockIdx.x;
  if (tj > ti) return;
  extern __shared__ __align__(16) double smem[];
  double* Pa = smem;                       // [64][kCBLd]
  double* Pb = smem + kCB * kCBLd;
  const int r0 = k0 + kCB + ti * kCB, c0 = k0 + kCB + tj * kCB;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wr = (w >> 1) * 32, wc = (w & 1) * 32;      // this warp's 32x32 quadrant
  const int fr = lane >> 2, fk = lane & 3;               // fragment row / k index of this lane
  // D = (-P_i) P_j^T + C: the accumulators start from the C tile, whose loads are issued before the
  // panel tiles are staged, so their latency hides behind the staging and no read-modify-write waits
  // at the end of the kernel
  double acc[4][4][2];
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      const double2 v = *reinterpret_cast<const double2*>(A + (size_t)(r0 + wr + mi * 8 + fr) * ld + c0 + wc + ni * 8 + 2 * fk);
      acc[mi][ni][0] = v.x; acc[mi][ni][1] = v.y;
    }
  for (int t = threadIdx.x; t < kCB * kCB / 2; t += 128) {
    const int i = t / (kCB / 2), c = (t % (kCB / 2)) * 2;
    const double2 va = *reinterpret_cast<const double2*>(A + (size_t)(r0 + i) * ld + k0 + c);
    const double2 vb = *reinterpret_cast<const double2*>(A + (size_t)(c0 + i) * ld + k0 + c);
    Pa[i * kCBLd + c] = -va.x; Pa[i * kCBLd + c + 1] = -va.y;
    Pb[i * kCBLd + c] = vb.x; Pb[i * kCBLd + c + 1] = vb.y;
  }
  __syncthreads();
#pragma unroll 4
  for (int kk = 0; kk < kCB; kk += 4) {
    double a[4], b[4];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) a[mi] = Pa[(wr + mi * 8 + fr) * kCBLd + kk + fk];
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) b[ni] = Pb[(wc + ni * 8 + fr) * kCBLd + kk + fk];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) dmma_m8n8k4(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
  }
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni)
      *reinterpret_cast<double2*>(A + (size_t)(r0 + wr + mi * 8 + fr) * ld + c0 + wc + ni * 8 + 2 * fk) =
          make_double2(acc[mi][ni][0], acc[mi][ni][1]);
}

// upper triangle <- transpose of the lower one (so that column sweeps read contiguous rows)
__global__ void chol_mirror_kernel(double* __restrict__ A, int64_t ld, int n_pad) {
  __shared__ double t[32][33];
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bj > bi) return;
  const int i = bi * 32 + threadIdx.y, j = bj * 32 + threadIdx.x;
  t[threadIdx.y][threadIdx.x] = A[(size_t)i * ld + j];
  __syncthreads();
  const int ui = bj * 32 + threadIdx.y, uj = bi * 32 + threadIdx.x;   // element (ui, uj) = L[uj][ui]
  if (uj > ui) A[(size_t)ui * ld + uj] = t[threadIdx.x][threadIdx.y];
}

// L y = b, L^T x = y with one CTA per right-hand side; the working vector lives in shared memory and
// both sweeps stream contiguous rows (row j of the mirrored upper triangle = column j of L).
__global__ void __launch_bounds__(1024)
dense_apply_kernel(const double* __restrict__ A, const double* __restrict__ b, double* __restrict__ x,
                   int n, int64_t lda, int64_t ld) {
  extern __shared__ double s_col[];
  const int tid = threadIdx.x, nt = blockDim.x;
  const double* bq = b + (size_t)blockIdx.x * ld;
  double* xq = x + (size_t)blockIdx.x * ld;
  for (int i = tid; i < n; i += nt) s_col[i] = bq[i];
  __syncthreads();
  for (int j = 0; j < n; ++j) {            // forward: y_j, then s[i] -= L[i][j] y_j (i > j)
    const double* row = A + (size_t)j * lda;
    const double yj = s_col[j] / row[j];
    __syncthreads();
    if (tid == 0) s_col[j] = yj;
    for (int i = j + 1 + tid; i < n; i += nt) s_col[i] -= row[i] * yj;
    __syncthreads();
  }
  for (int j = n - 1; j >= 0; --j) {       // backward: x_j, then s[i] -= L[j][i] x_j (i < j)
    const double* row = A + (size_t)j * lda;
    const double xj = s_col[j] / row[j];
    __syncthreads();
    if (tid == 0) s_col[j] = xj;
    for (int i = tid; i < j; i += nt) s_col[i] -= row[i] * xj;
    __syncthreads();
  }
  for (int i = tid; i < n; i += nt) xq[i] = s_col[i];
}

int64_t dense_max_dof() { return kDenseMax; }

int dense_factor(femb_handle* h) {
  const int64_t n = h->ndof;
  if (n > kDenseMax) return fail(h, FEMB_ERR_ARG, "dense solver is limited to 16384 DOFs");
  if (h->dense_factored) return FEMB_OK;
  const int64_t n_pad = (n + kCB - 1) / kCB * kCB;
  DevBuf<int> status;
  FEMB_CUDA(h, h->denseL.alloc((size_t)n_pad * n_pad));
  FEMB_CUDA(h, status.alloc(1));
  FEMB_CUDA(h, cudaMemsetAsync(status.p, 0, sizeof(int), h->stream));
  dense_fill_kernel<<<(unsigned)n_pad, 256, 0, h->stream>>>(h->rowptr.p, h->colidx.p, h->Kvals.p, h->free_mask.p, h->denseL.p, n, n_pad, h->bs);
  h->launches++;
  const size_t smem = (size_t)2 * kCB * kCBLd * sizeof(double);
  FEMB_CUDA(h, cudaFuncSetAttribute(chol_syrk_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int64_t k0 = 0; k0 < n_pad; k0 += kCB) {
    launch_potrf_block(h, h->denseL.p, n_pad, (int)k0, status.p);
    h->launches++;
    const int64_t rem = n_pad - k0 - kCB;
    if (rem > 0) {
      chol_trsm_kernel<<<(unsigned)((rem + 127) / 128), 128, 0, h->stream>>>(h->denseL.p, n_pad, (int)k0, (int)n_pad);
      const unsigned nt = (unsigned)(rem / kCB);
      chol_syrk_dmma_kernel<<<dim3(nt, nt), 128, smem, h->stream>>>(h->denseL.p, n_pad, (int)k0);
      h->launches += 2;
    }
  }
  chol_mirror_kernel<<<dim3((unsigned)(n_pad / 32), (unsigned)(n_pad / 32)), dim3(32, 32), 0, h->stream>>>(h->denseL.p, n_pad, (int)n_pad);
  h->launches++;
  FEMB_CUDA(h, cudaGetLastError());
  int* hs = reinterpret_cast<int*>(h->pinned);
  FEMB_CUDA(h, cudaMemcpyAsync(hs, status.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  FEMB_CUDA(h, cudaStreamSynchronize(h->stream));
  if (*hs != 0) return fail(h, FEMB_ERR_SINGULAR, "dense Cholesky hit a non-positive pivot (K_ff not positive definite)");
  h->dense_factored = true;
  return FEMB_OK;
}

int dense_apply(femb_handle* h, const double* d_b, double* d_x, int nrhs, int64_t ld) {
  const int64_t n = h->ndof;
  const int64_t n_pad = (n + kCB - 1) / kCB * kCB;
  const size_t smem = (size_t)n * sizeof(double);
  FEMB_CUDA(h, cudaFuncSetAttribute(dense_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024)));
  dense_apply_kernel<<<nrhs, 1024, smem, h->stream>>>(h->denseL.p, d_b, d_x, (int)n, n_pad, ld);
  h->launches++;
  FEMB_CUDA(h, cudaGetLastError());
  return FEMB_OK;
}

int run_dense_solve(femb_handle* h, femb_stats* st) {
  int rc = setup_rhs_for_direct(h);
  if (rc) return rc;
  rc = dense_factor(h);
  if (rc) return rc;
  rc = dense_apply(h, h->b.p, h->x.p, 1, h->ndof);
  if (rc) return rc;
  if (st) { st->method_used = FEMB_SOLVER_DENSE; st->iterations = 1; st->converged = 1; st->rel_residual = 0.0; }
  return FEMB_OK;
}

// ---- explicit inverse of the coarse Galerkin matrix (two-level preconditioner, twolevel.cu) --------
// The coarse solve runs once per CG iteration between two 20-us kernels; two triangular sweeps
// (n sequential steps) would cost more than the whole iteration, a product with the explicit inverse
// is one bandwidth-bound pass over n^2 numbers that stay in L2.  The inverse comes out of the SAME
// three kernels as the factorisation: eliminate the first n columns of the augmented matrix
//     [ Kc  . ]        after panel k0 the identity rows n+i with i >= k0+64 are still zero in
//     [ I   0 ]        every eliminated column, so each step works on the n rows [k0+64, n+k0+64)
// and the Schur complement left in the lower right block is 0 - I Kc^-1 I = -Kc^-1
// (chol_trsm turns the identity rows into L^-T, chol_syrk_dmma subtracts their Gram matrix).
// n^3 DMMA flops in n/64 steps of equal size.  `aug` is (2 n_pad)^2, zero except Kc and the identity.
__global__ void coarse_extract_kernel(const double* __restrict__ aug, int64_t ld, int64_t n_pad,
                                      double* __restrict__ inv) {
  const int64_t i = blockIdx.y, j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_pad) return;
  const int64_t r = i > j ? i : j, c = i > j ? j : i;      // the lower tiles hold the result
  inv[i * n_pad + j] = -aug[(n_pad + r) * ld + n_pad + c];
}

// the same elimination enqueued on `stream` without a host round trip: *status_dev (zeroed by the caller) counts
// non-positive pivots; used by the line preconditioner's setup to run its three inversions on three streams
int coarse_invert_async(femb_handle* h, cudaStream_t stream, double* aug, int64_t n_pad, double* inv, int* status_dev) {
  const int64_t m = 2 * n_pad;
  const size_t smem = (size_t)2 * kCB * kCBLd * sizeof(double);
  FEMB_CUDA(h, cudaFuncSetAttribute(chol_syrk_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const unsigned nt = (unsigned)(n_pad / kCB);
  for (int64_t k0 = 0; k0 < n_pad; k0 += kCB) {
    chol_potrf_block_lean_kernel<<<1, kCB, 0, stream>>>(aug, m, (int)k0, status_dev);
    const int64_t row_end = n_pad + k0 + kCB;
    chol_trsm_kernel<<<(unsigned)((n_pad + 127) / 128), 128, 0, stream>>>(aug, m, (int)k0, (int)row_end);
    chol_syrk_dmma_kernel<<<dim3(nt, nt), 128, smem, stream>>>(aug, m, (int)k0);
    h->launches += 3;
  }
  coarse_extract_kernel<<<dim3((unsigned)((n_pad + 255) / 256), (unsigned)n_pad), 256, 0, stream>>>(aug, m, n_pad, inv);
  h->launches++;
  FEMB_CUDA(h, cudaGetLastError());
  return FEMB_OK;
}

int coarse_invert(femb_handle* h, double* aug, int64_t n_pad, double* inv, bool* ok) {
  const int64_t m = 2 * n_pad;
  DevBuf<int> status;
  FEMB_CUDA(h, status.alloc(1));
  FEMB_CUDA(h, cudaMemsetAsync(status.p, 0, sizeof(int), h->stream));
  const size_t smem = (size_t)2 * kCB * kCBLd * sizeof(double);
  FEMB_CUDA(h, cudaFuncSetAttribute(chol_syrk_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const unsigned nt = (unsigned)(n_pad / kCB);
  for (int64_t k0 = 0; k0 < n_pad; k0 += kCB) {
    launch_potrf_block(h, aug, m, (int)k0, status.p);
    const int64_t row_end = n_pad + k0 + kCB;                 // <= m; rows beyond are still zero in this panel
    chol_trsm_kernel<<<(unsigned)((n_pad + 127) / 128), 128, 0, h->stream>>>(aug, m, (int)k0, (int)row_end);
    chol_syrk_dmma_kernel<<<dim3(nt, nt), 128, smem, h->stream>>>(aug, m, (int)k0);
    h->launches += 3;
  }
  coarse_extract_kernel<<<dim3((unsigned)((n_pad + 255) / 256), (unsigned)n_pad), 256, 0, h->stream>>>(aug, m, n_pad, inv);
  h->launches++;
  FEMB_CUDA(h, cudaGetLastError());
  int* hs = reinterpret_cast<int*>(h->pinned);
  FEMB_CUDA(h, cudaMemcpyAsync(hs, status.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  FEMB_CUDA(h, cudaStreamSynchronize(h->stream));
  *ok = (*hs == 0);
  return FEMB_OK;
}

}  // namespace femb
