// Direct solvers for the cases CG cannot serve:
//   * chains (path graphs): block-tridiagonal factorisation with 6x6 blocks — beam chains
//     have cond(K) ~ (L/h)^4, hopeless for Jacobi-PCG (SURVEY §7 "hard parts");
//   * batched independent chain models (BASELINE config 4): the same recurrence with the
//     element blocks generated on the fly from the section record (fused element + solve),
//     one thread per model, model-interleaved scratch so every access is coalesced;
//   * small systems: dense FP64 Cholesky of the masked operator.
// All replace np.linalg.solve(k_ff, f_f) (BeamSolver.py:417) / spsolve (ReactionSolver.py:201).
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
// Scratch layout is model-interleaved: W[(k*36+q)*nm + m], z[(k*6+c)*nm + m], so the 32
// threads of a warp (32 consecutive models) touch 32 consecutive doubles.
struct BatchParams {
  const double* xyz;        // (n_nodes,3) shared
  const double* sec_props;  // (n_models,8)
  const uint8_t* fixed;     // (ndof) 1 = fixed
  const double* f;          // (n_models, ndof)
  double* u;                // (n_models, ndof)
  double* W;                // scratch
  double* z;                // scratch
  int* status;              // count of models with a non-positive pivot
  int64_t n_models, n_elem;
  double E, G;
};

__global__ void __launch_bounds__(64)
batch_chain_kernel(BatchParams B) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= B.n_models) return;
  const int64_t nm = B.n_models, nn = B.n_elem + 1, ndof = nn * 6;
  // per-model element generator: same record code as the assembly path
  int32_t conn2[2] = {0, 1};
  int32_t sec0 = 0;
  FrameParams P;
  P.conn = conn2; P.elem_sec = &sec0; P.sec_props = B.sec_props + 8 * m;
  P.E = B.E; P.G = B.G; P.rho = 0.0;
  const double* f = B.f + m * ndof;
  double S[36], G[36], O[36], Wk[36], y[6];
  FrameRec R;
  auto fixmask = [&](int64_t node) {
    unsigned mk = 0;
#pragma unroll
    for (int c = 0; c < 6; ++c) mk |= (B.fixed[node * 6 + c] ? 0u : 1u) << c;
    return mk;
  };
  bool ok = true;
  // node 0: D_0 = K_e0[0][0]
  P.xyz = B.xyz;
  frame_record(P, 0, R);
  frame_kblock<true>(R, 0, 0, S);
  unsigned mi = fixmask(0);
  mask_block(S, mi, mi, true);
#pragma unroll
  for (int c = 0; c < 6; ++c) y[c] = ((mi >> c) & 1u) ? f[c] : 0.0;
  for (int64_t k = 0; k < nn; ++k) {
    ok = inv6_spd(S, G) && ok;
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      double s = 0.0;
#pragma unroll
      for (int c = 0; c < 6; ++c) s += G[r * 6 + c] * y[c];
      B.z[(k * 6 + r) * nm + m] = s;
    }
    if (k + 1 == nn) break;
    // element k joins nodes k, k+1: record R currently holds element k
    const unsigned mj = fixmask(k + 1);
    frame_kblock<true>(R, 1, 0, O);          // K[k+1][k]
    mask_block(O, mj, mi, false);
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        double s = 0.0;
#pragma unroll
        for (int t = 0; t < 6; ++t) s += O[r * 6 + t] * G[t * 6 + c];
        Wk[r * 6 + c] = s;
      }
#pragma unroll
    for (int q = 0; q < 36; ++q) B.W[(k * 36 + q) * nm + m] = Wk[q];
    frame_kblock<true>(R, 1, 1, S);          // element k's share of D_{k+1}
    if (k + 1 < B.n_elem) {                  // plus element k+1's [0][0]
      P.xyz = B.xyz + 3 * (k + 1);
      frame_record(P, 0, R);
      frame_kblock<false>(R, 0, 0, S);
    }
    mask_block(S, mj, mj, true);
    double yn[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      double s = ((mj >> r) & 1u) ? f[(k + 1) * 6 + r] : 0.0;
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        s -= Wk[r * 6 + c] * y[c];
        double t = 0.0;
#pragma unroll
        for (int q = 0; q < 6; ++q) t += Wk[r * 6 + q] * O[c * 6 + q];
        S[r * 6 + c] -= t;
      }
      yn[r] = s;
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) y[c] = yn[c];
    mi = mj;
  }
  double xn[6];
  double* u = B.u + m * ndof;
#pragma unroll
  for (int c = 0; c < 6; ++c) { xn[c] = B.z[((nn - 1) * 6 + c) * nm + m]; u[(nn - 1) * 6 + c] = xn[c]; }
  for (int64_t k = nn - 2; k >= 0; --k) {
    double xk[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) xk[r] = B.z[(k * 6 + r) * nm + m];
#pragma unroll
    for (int c = 0; c < 6; ++c) {
#pragma unroll
      for (int r = 0; r < 6; ++r) xk[r] -= B.W[(k * 36 + c * 6 + r) * nm + m] * xn[c];
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) { xn[c] = xk[c]; u[k * 6 + c] = xk[c]; }
  }
  if (!ok) atomicAdd(B.status, 1);
}

int run_batch_chain(femb_handle* h, int64_t n_models, int64_t n_elem, const double* xyz,
                    const double* sec_props, double E, double G, const uint8_t* fixed_mask,
                    const double* f, double* u, femb_stats* st) {
  const int64_t nn = n_elem + 1, ndof = nn * 6;
  DevBuf<double> dxyz, dsec, df, du, W, z;
  DevBuf<uint8_t> dfix;
  DevBuf<int> status;
  FEMB_CUDA(h, upload(dxyz, xyz, (size_t)nn * 3, h->stream));
  FEMB_CUDA(h, upload(dsec, sec_props, (size_t)n_models * 8, h->stream));
  FEMB_CUDA(h, upload(dfix, fixed_mask, (size_t)ndof, h->stream));
  FEMB_CUDA(h, upload(df, f, (size_t)n_models * ndof, h->stream));
  FEMB_CUDA(h, du.alloc((size_t)n_models * ndof));
  FEMB_CUDA(h, W.alloc((size_t)n_models * nn * 36));
  FEMB_CUDA(h, z.alloc((size_t)n_models * nn * 6));
  FEMB_CUDA(h, status.alloc(1));
  FEMB_CUDA(h, cudaMemsetAsync(status.p, 0, sizeof(int), h->stream));
  BatchParams B{dxyz.p, dsec.p, dfix.p, df.p, du.p, W.p, z.p, status.p, n_models, n_elem, E, G};
  FEMB_CUDA(h, cudaEventRecord(h->ev0, h->stream));
  batch_chain_kernel<<<(unsigned)((n_models + 63) / 64), 64, 0, h->stream>>>(B);
  h->launches++;
  FEMB_CUDA(h, cudaGetLastError());
  FEMB_CUDA(h, cudaEventRecord(h->ev1, h->stream));
  int* hs = reinterpret_cast<int*>(h->pinned);
  FEMB_CUDA(h, cudaMemcpyAsync(hs, status.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  if (u) FEMB_CUDA(h, download(u, du.p, (size_t)n_models * ndof * 8, h->stream));
  FEMB_CUDA(h, cudaStreamSynchronize(h->stream));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, h->ev0, h->ev1);
  if (st) { st->method_used = FEMB_SOLVER_CHAIN; st->iterations = 1; st->converged = (*hs == 0); st->device_ms = ms; }
  if (*hs != 0) return fail(h, FEMB_ERR_SINGULAR, "batched chain solve: non-positive pivot in at least one model");
  return FEMB_OK;
}

// ---- small dense Cholesky of the masked operator ----------------------------------------
__global__ void dense_fill_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                  const double* __restrict__ vals, const uint8_t* __restrict__ mask,
                                  double* __restrict__ A, int64_t n, int bs) {
  // one thread per scalar row: writes the row of A = P K P + (I - P), column-major is the
  // same as row-major here (symmetric); we store row-major A[r*n + c].
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  const int64_t node = g / bs;
  const int r = (int)(g - node * bs);
  double* row = A + (size_t)g * n;
  for (int64_t c = 0; c < n; ++c) row[c] = 0.0;
  if (!mask[g]) { row[g] = 1.0; return; }
  for (int b = rowptr[node]; b < rowptr[node + 1]; ++b) {
    const int64_t cn = colidx[b];
    for (int c = 0; c < bs; ++c)
      if (mask[cn * bs + c]) row[cn * bs + c] = vals[(size_t)b * bs * bs + r * bs + c];
  }
}

// right-looking Cholesky, one CTA, A row-major lower triangle overwritten by L.
__global__ void __launch_bounds__(1024)
dense_chol_kernel(double* __restrict__ A, int n, int* status) {
  __shared__ double s_col[2048];
  __shared__ int s_bad;
  const int tid = threadIdx.x, nt = blockDim.x;
  if (tid == 0) s_bad = 0;
  __syncthreads();
  for (int j = 0; j < n; ++j) {
    const double ajj = A[(size_t)j * n + j];
    if (!(ajj > 0.0)) { if (tid == 0) s_bad = 1; }
    const double d = sqrt(ajj > 0.0 ? ajj : 1.0);
    const double id = 1.0 / d;
    for (int i = j + tid; i < n; i += nt) {
      const double v = (i == j) ? d : A[(size_t)i * n + j] * id;
      A[(size_t)i * n + j] = v;
      s_col[i] = v;
    }
    __syncthreads();
    // trailing update: A[i][c] -= L[i][j] L[c][j], j < c <= i ; rows distributed by warps
    const int w = tid >> 5, l = tid & 31, nw = nt >> 5;
    for (int i = j + 1 + w; i < n; i += nw) {
      const double lij = s_col[i];
      for (int c = j + 1 + l; c <= i; c += 32) A[(size_t)i * n + c] -= lij * s_col[c];
    }
    __syncthreads();
  }
  if (tid == 0) *status = s_bad;
}

// forward / backward substitution with the Cholesky factor; one CTA per right-hand side.
__global__ void __launch_bounds__(1024)
dense_apply_kernel(const double* __restrict__ A, const double* __restrict__ b, double* __restrict__ x,
                   int n, int64_t ld) {
  __shared__ double s_col[2048];
  const int tid = threadIdx.x, nt = blockDim.x;
  const double* bq = b + (size_t)blockIdx.x * ld;
  double* xq = x + (size_t)blockIdx.x * ld;
  for (int i = tid; i < n; i += nt) s_col[i] = bq[i];
  __syncthreads();
  for (int j = 0; j < n; ++j) {            // L y = b (column sweep)
    if (tid == 0) s_col[j] = s_col[j] / A[(size_t)j * n + j];
    __syncthreads();
    const double yj = s_col[j];
    for (int i = j + 1 + tid; i < n; i += nt) s_col[i] -= A[(size_t)i * n + j] * yj;
    __syncthreads();
  }
  for (int j = n - 1; j >= 0; --j) {       // L^T x = y
    if (tid == 0) s_col[j] = s_col[j] / A[(size_t)j * n + j];
    __syncthreads();
    const double xj = s_col[j];
    for (int i = tid; i < j; i += nt) s_col[i] -= A[(size_t)j * n + i] * xj;
    __syncthreads();
  }
  for (int i = tid; i < n; i += nt) xq[i] = s_col[i];
}

int dense_factor(femb_handle* h) {
  const int64_t n = h->ndof;
  if (n > 2048) return fail(h, FEMB_ERR_ARG, "dense solver is limited to 2048 DOFs");
  if (h->dense_factored) return FEMB_OK;
  DevBuf<int> status;
  FEMB_CUDA(h, h->denseL.alloc((size_t)n * n));
  FEMB_CUDA(h, status.alloc(1));
  dense_fill_kernel<<<(unsigned)((n + 127) / 128), 128, 0, h->stream>>>(h->rowptr.p, h->colidx.p, h->Kvals.p, h->free_mask.p, h->denseL.p, n, h->bs);
  dense_chol_kernel<<<1, 1024, 0, h->stream>>>(h->denseL.p, (int)n, status.p);
  h->launches += 2;
  FEMB_CUDA(h, cudaGetLastError());
  int* hs = reinterpret_cast<int*>(h->pinned);
  FEMB_CUDA(h, cudaMemcpyAsync(hs, status.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  FEMB_CUDA(h, cudaStreamSynchronize(h->stream));
  if (*hs != 0) return fail(h, FEMB_ERR_SINGULAR, "dense Cholesky hit a non-positive pivot (K_ff not positive definite)");
  h->dense_factored = true;
  return FEMB_OK;
}

int dense_apply(femb_handle* h, const double* d_b, double* d_x, int nrhs, int64_t ld) {
  dense_apply_kernel<<<nrhs, 1024, 0, h->stream>>>(h->denseL.p, d_b, d_x, (int)h->ndof, ld);
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

}  // namespace femb
