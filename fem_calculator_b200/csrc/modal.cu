// Lowest-k modes of K_ff phi = lambda M_ff phi (replaces inv(m_ff) @ k_ff + the unshifted QR
// iteration of BeamSolver.py:440-455,467-481).
//
// Block shift-invert Krylov (sigma = 0) with Rayleigh-Ritz and full M-orthogonalisation:
//   V <- [V, R];  W = K_ff^-1 (M R);  H = V^T M W (symmetric, since V is M-orthonormal);
//   eig(H) -> theta (largest) = 1/lambda (smallest);  R <- M-orthonormalised remainder of W.
// K_ff^-1 is whichever static solver fits the mesh: the persistent chain / dense factors or
// PCG on the masked BSR operator.  M is the 6x6-block-diagonal lumped mass.  All vectors keep
// exact zeros on fixed DOFs, so the pencil is the eliminated one.  Block size >= 2 keeps the
// double eigenvalues of symmetric sections (circular, square) in the basis.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "common.cuh"

namespace femb {

constexpr int kMT = 256;

// y_q = P M P x_q for nv vectors (leading dimension ld); one thread per scalar row and vector.
__global__ void mass_apply_kernel(const double* __restrict__ Mdiag, const uint8_t* __restrict__ mask,
                                  const double* __restrict__ x, double* __restrict__ y, int64_t n, int64_t ld, int nv) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int q = blockIdx.y;
  if (g >= n || q >= nv) return;
  const int64_t node = g / 6;
  const int r = (int)(g - node * 6);
  const double* xq = x + (size_t)q * ld;
  double s = 0.0;
  if (mask[g]) {
    const double* m = Mdiag + (size_t)node * 36 + r * 6;
#pragma unroll
    for (int c = 0; c < 6; ++c) s += m[c] * (mask[node * 6 + c] ? xq[node * 6 + c] : 0.0);
  }
  y[(size_t)q * ld + g] = s;
}

// partial[(j*nb + q)*nchunk + chunk] = sum over the chunk of A_j[i] * Y_q[i]
__global__ void __launch_bounds__(kMT)
multi_dot_kernel(const double* __restrict__ A, int64_t lda, const double* __restrict__ Y, int64_t ldy,
                 int64_t n, int nb, double* __restrict__ partial, int nchunk) {
  __shared__ double s_red[kMT / 32];
  const int j = blockIdx.y, q = blockIdx.z, chunk = blockIdx.x;
  const double* a = A + (size_t)j * lda;
  const double* y = Y + (size_t)q * ldy;
  double acc = 0.0;
  for (int64_t i = (int64_t)chunk * kMT + threadIdx.x; i < n; i += (int64_t)nchunk * kMT) acc += a[i] * y[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double t = threadIdx.x < kMT / 32 ? s_red[threadIdx.x] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) partial[((size_t)j * nb + q) * nchunk + chunk] = t;
  }
}

// out[j*nb+q] = ordered sum of the chunk partials
__global__ void dot_finish_kernel(const double* __restrict__ partial, int nchunk, int total, double* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  double s = 0.0;
  for (int c = 0; c < nchunk; ++c) s += partial[(size_t)t * nchunk + c];
  out[t] = s;
}

// W_q -= sum_j C[j*nb+q] A_j    (q < nb <= 8)
template <int NB>
__global__ void multi_axpy_kernel(const double* __restrict__ A, int64_t lda, int m, const double* __restrict__ C,
                                  double* __restrict__ W, int64_t ldw, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double acc[NB];
#pragma unroll
  for (int q = 0; q < NB; ++q) acc[q] = 0.0;
  for (int j = 0; j < m; ++j) {
    const double a = A[(size_t)j * lda + i];
#pragma unroll
    for (int q = 0; q < NB; ++q) acc[q] += __ldg(C + j * NB + q) * a;
  }
#pragma unroll
  for (int q = 0; q < NB; ++q) W[(size_t)q * ldw + i] -= acc[q];
}

__global__ void scale_kernel(double* __restrict__ x, double s, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] *= s;
}

__global__ void random_fill_kernel(double* __restrict__ x, const uint8_t* __restrict__ mask, int64_t n, uint32_t seed) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t h = (uint32_t)i * 2654435761u ^ (seed * 0x9E3779B9u + 0x85EBCA6Bu);
  h ^= h >> 16; h *= 0x7FEB352Du; h ^= h >> 15; h *= 0x846CA68Bu; h ^= h >> 16;
  x[i] = mask[i] ? ((double)h / 4294967296.0 - 0.5) : 0.0;
}

// Phi_c = sum_j V_j S[j*k + c]  for c < kout (<= 32)
__global__ void ritz_vectors_kernel(const double* __restrict__ V, int64_t ldv, int m, const double* __restrict__ S,
                                    int kout, double* __restrict__ Phi, int64_t ldp, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int c0 = blockIdx.y * 8;
  if (i >= n) return;
  double acc[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] = 0.0;
  for (int j = 0; j < m; ++j) {
    const double v = V[(size_t)j * ldv + i];
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (c0 + c < kout) acc[c] += __ldg(S + (size_t)j * kout + c0 + c) * v;
  }
#pragma unroll
  for (int c = 0; c < 8; ++c)
    if (c0 + c < kout) Phi[(size_t)(c0 + c) * ldp + i] = acc[c];
}

// ---- host: cyclic Jacobi eigen-decomposition of a small symmetric matrix ----------------
static void jacobi_eigh(std::vector<double>& A, int n, std::vector<double>& w, std::vector<double>& Vv) {
  Vv.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) Vv[(size_t)i * n + i] = 1.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < n; ++i) {
      diag += A[(size_t)i * n + i] * A[(size_t)i * n + i];
      for (int j = i + 1; j < n; ++j) off += A[(size_t)i * n + j] * A[(size_t)i * n + j];
    }
    if (off <= 1e-30 * (diag + 1e-300)) break;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = A[(size_t)p * n + q];
        if (apq == 0.0) continue;
        const double app = A[(size_t)p * n + p], aqq = A[(size_t)q * n + q];
        const double tau = (aqq - app) / (2.0 * apq);
        const double t = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
        const double c = 1.0 / std::sqrt(1.0 + t * t), s = t * c;
        for (int k = 0; k < n; ++k) {
          const double akp = A[(size_t)k * n + p], akq = A[(size_t)k * n + q];
          A[(size_t)k * n + p] = c * akp - s * akq;
          A[(size_t)k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {
          const double apk = A[(size_t)p * n + k], aqk = A[(size_t)q * n + k];
          A[(size_t)p * n + k] = c * apk - s * aqk;
          A[(size_t)q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) {
          const double vkp = Vv[(size_t)k * n + p], vkq = Vv[(size_t)k * n + q];
          Vv[(size_t)k * n + p] = c * vkp - s * vkq;
          Vv[(size_t)k * n + q] = s * vkp + c * vkq;
        }
      }
  }
  w.resize(n);
  for (int i = 0; i < n; ++i) w[i] = A[(size_t)i * n + i];
}

// ---- device helpers ---------------------------------------------------------------------
struct ModalWs {
  femb_handle* h;
  int64_t n;
  const uint8_t* mask;   // free-DOF mask; in a row-block partition the ghost tail is cleared
  int64_t n_own;         // rows this rank owns (== n on one GPU)
  int nchunk;
  DevBuf<double> partial, dots;
  std::vector<double> hdots;
};

// out (host) [m*nb] = A^T Y  (A: m vectors, Y: nb vectors)
static int dots(ModalWs& ws, const double* A, int64_t lda, int m, const double* Y, int64_t ldy, int nb, double* out) {
  femb_handle* h = ws.h;
  if (m == 0 || nb == 0) return FEMB_OK;
  FEMB_CUDA(h, ws.partial.ensure((size_t)m * nb * ws.nchunk));
  FEMB_CUDA(h, ws.dots.ensure((size_t)m * nb));
  dim3 grid(ws.nchunk, m, nb);
  multi_dot_kernel<<<grid, kMT, 0, h->stream>>>(A, lda, Y, ldy, ws.n, nb, ws.partial.p, ws.nchunk);
  dot_finish_kernel<<<(m * nb + 127) / 128, 128, 0, h->stream>>>(ws.partial.p, ws.nchunk, m * nb, ws.dots.p);
  h->launches += 2;
  FEMB_CUDA(h, cudaGetLastError());
  int rcd = dist_allreduce(h, ws.dots.p, m * nb);   // row-block partition: ghost tails are zero, ranks sum
  if (rcd) return rcd;
  FEMB_CUDA(h, cudaMemcpyAsync(out, ws.dots.p, sizeof(double) * m * nb, cudaMemcpyDeviceToHost, h->stream));
  FEMB_CUDA(h, cudaStreamSynchronize(h->stream));
  return FEMB_OK;
}

// W_q -= sum_j C[j*nb+q] A_j with host coefficients C (m x nb, row-major)
static int axpy_block(ModalWs& ws, const double* A, int64_t lda, int m, const std::vector<double>& C, double* W,
                      int64_t ldw, int nb, DevBuf<double>& dC) {
  femb_handle* h = ws.h;
  if (m == 0) return FEMB_OK;
  FEMB_CUDA(h, dC.ensure((size_t)m * nb));
  FEMB_CUDA(h, cudaMemcpyAsync(dC.p, C.data(), (size_t)m * nb * 8, cudaMemcpyHostToDevice, h->stream));
  const int grid = (int)((ws.n + kMT - 1) / kMT);
  switch (nb) {
    case 1: multi_axpy_kernel<1><<<grid, kMT, 0, h->stream>>>(A, lda, m, dC.p, W, ldw, ws.n); break;
    case 2: multi_axpy_kernel<2><<<grid, kMT, 0, h->stream>>>(A, lda, m, dC.p, W, ldw, ws.n); break;
    case 3: multi_axpy_kernel<3><<<grid, kMT, 0, h->stream>>>(A, lda, m, dC.p, W, ldw, ws.n); break;
    case 4: multi_axpy_kernel<4><<<grid, kMT, 0, h->stream>>>(A, lda, m, dC.p, W, ldw, ws.n); break;
    default: return fail(h, FEMB_ERR_ARG, "modal block size must be 1..4");
  }
  h->launches++;
  FEMB_CUDA(h, cudaGetLastError());
  FEMB_CUDA(h, cudaStreamSynchronize(h->stream));  // C may be reused by the caller right away
  return FEMB_OK;
}

static int mass_apply(ModalWs& ws, const double* x, double* y, int nv) {
  femb_handle* h = ws.h;
  dim3 grid((unsigned)((ws.n + kMT - 1) / kMT), nv);
  mass_apply_kernel<<<grid, kMT, 0, h->stream>>>(h->Mdiag.p, ws.mask, x, y, ws.n, ws.n, nv);
  h->launches++;
  FEMB_CUDA(h, cudaGetLastError());
  return FEMB_OK;
}

// X (nb vectors) = K_ff^-1 B with the solver that fits the mesh; B must be masked.
static int solve_block(ModalWs& ws, int method, const femb_solve_opts& so, const double* B, double* X, int nb,
                       femb_stats* st) {
  femb_handle* h = ws.h;
  if (dist_active(h)) {
    // one distributed PCG per right-hand side; only the owned rows are meaningful, tails stay zero
    for (int q = 0; q < nb; ++q) {
      femb_stats s1;
      std::memset(&s1, 0, sizeof(s1));
      int rc = dist_solve_rhs(h, so, B + (size_t)q * ws.n, &s1);
      if (rc) return rc;
      st->iterations += s1.iterations;
      st->spmv_launches += s1.spmv_launches;
      st->precond_used = s1.precond_used;
      st->coarse_dim = s1.coarse_dim;
      FEMB_CUDA(h, cudaMemsetAsync(X + (size_t)q * ws.n, 0, ws.n * 8, h->stream));
      FEMB_CUDA(h, cudaMemcpyAsync(X + (size_t)q * ws.n, h->x.p, ws.n_own * 8, cudaMemcpyDeviceToDevice, h->stream));
    }
    return FEMB_OK;
  }
  if (method == FEMB_SOLVER_CHAIN) return chain_apply(h, B, X, nb, ws.n);
  if (method == FEMB_SOLVER_DENSE) return dense_apply(h, B, X, nb, ws.n);
  // two-level preconditioner: one right-hand side at a time through the three-kernel iteration of
  // twolevel.cu (the coarse inverse is built once per K / BC and shared by every solve)
  const bool two_level = lines_applicable(h, so) || twolevel_applicable(h, so);
  if (h->bs == 6 && !two_level && !getenv("FEMB_MODAL_SINGLE_RHS")) {
    // all right-hand sides of the block advance in lockstep and share one matrix pass per iteration
    femb_stats s1;
    std::memset(&s1, 0, sizeof(s1));
    int rc = pcg_solve_multi(h, so, B, ws.n, nb, X, ws.n, &s1);
    st->op_used = s1.op_used;
    st->iterations += s1.iterations;
    st->spmv_launches += s1.spmv_launches;
    // (Tried: a Galerkin start X0 = OPV (OPV^T MV)^-1 OPV^T B from the solves already done, to take the
    // lowest modes out of the residual.  At 1M DOF the 2-norm of the projected right-hand side GROWS
    // 3-100x and every block still needs ~7,300 lockstep iterations (profiles/r01_modal_galerkin_start_negative.log):
    // CG on this operator is not held back by the few lowest modes but by the wide axial / bending
    // stiffness spread, so the idea was dropped.)
    return rc;
  }
  for (int q = 0; q < nb; ++q) {
    femb_stats s1;
    std::memset(&s1, 0, sizeof(s1));
    int rc = pcg_solve_rhs(h, so, B + (size_t)q * ws.n, &s1);
    if (rc) return rc;
    st->op_used = s1.op_used;
    st->coarse_dim = s1.coarse_dim;
    st->precond_used = s1.precond_used;
    st->iterations += s1.iterations;
    st->spmv_launches += s1.spmv_launches;
    FEMB_CUDA(h, cudaMemcpyAsync(X + (size_t)q * ws.n, h->x.p, ws.n * 8, cudaMemcpyDeviceToDevice, h->stream));
  }
  return FEMB_OK;
}

// M-orthogonalise w (device, n) against the m basis vectors (V, MV) twice, then against the
// `na` already accepted candidates (Wc, MWc).  Returns w^T M w in *nrm2 with M w left in mw.
static int orth_candidate(ModalWs& ws, const double* V, const double* MV, int m, const double* Wc, const double* MWc,
                          int na, double* w, double* mw, DevBuf<double>& dC, std::vector<double>& coef, double* nrm2) {
  const int64_t n = ws.n;
  int rc;
  for (int pass = 0; pass < 2; ++pass) {
    if (m > 0) {
      coef.assign(m, 0.0);
      rc = dots(ws, MV, n, m, w, n, 1, coef.data());
      if (rc) return rc;
      rc = axpy_block(ws, V, n, m, coef, w, n, 1, dC);
      if (rc) return rc;
    }
    if (na > 0) {
      coef.assign(na, 0.0);
      rc = dots(ws, MWc, n, na, w, n, 1, coef.data());
      if (rc) return rc;
      rc = axpy_block(ws, Wc, n, na, coef, w, n, 1, dC);
      if (rc) return rc;
    }
  }
  rc = mass_apply(ws, w, mw, 1);
  if (rc) return rc;
  return dots(ws, w, n, 1, mw, n, 1, nrm2);
}

// X <- X S  (X: m vectors -> q vectors, in place through Tmp)
static int compress_basis(ModalWs& ws, double* X, int m, const double* dS, int q, double* Tmp) {
  femb_handle* h = ws.h;
  const int64_t n = ws.n;
  dim3 rg((unsigned)((n + kMT - 1) / kMT), (q + 7) / 8);
  ritz_vectors_kernel<<<rg, kMT, 0, h->stream>>>(X, n, m, dS, q, Tmp, n, n);
  h->launches++;
  FEMB_CUDA(h, cudaGetLastError());
  FEMB_CUDA(h, cudaMemcpyAsync(X, Tmp, (size_t)q * n * 8, cudaMemcpyDeviceToDevice, h->stream));
  return FEMB_OK;
}

int run_modal(femb_handle* h, const femb_eig_opts& o, double* lambda_out, double* phi_out, int32_t* n_found,
              femb_stats* st) {
  const int64_t n = h->ndof;
  const bool dist = dist_active(h);
  const int64_t n_own = dist ? h->n_owned_nodes * h->bs : n;
  int64_t nfree = n - h->n_fixed;
  DevBuf<uint8_t> own_mask;
  DevBuf<double> cnt;
  if (dist) {
    // global number of free DOFs = sum over ranks of the owned free DOFs
    int64_t own_fixed = 0;
    for (int64_t d : h->h_fixed) own_fixed += (d < n_own) ? 1 : 0;
    double* hc = reinterpret_cast<double*>(h->pinned);
    hc[0] = (double)(n_own - own_fixed);
    FEMB_CUDA(h, cnt.alloc(1));
    FEMB_CUDA(h, cudaMemcpyAsync(cnt.p, hc, 8, cudaMemcpyHostToDevice, h->stream));
    int rcc = dist_allreduce(h, cnt.p, 1);
    if (rcc) return rcc;
    FEMB_CUDA(h, cudaMemcpyAsync(hc, cnt.p, 8, cudaMemcpyDeviceToHost, h->stream));
    FEMB_CUDA(h, cudaStreamSynchronize(h->stream));
    nfree = (int64_t)(hc[0] + 0.5);
    // modal vectors keep exact zeros on the ghost tail: their local dot products are the owned parts
    FEMB_CUDA(h, own_mask.alloc((size_t)n));
    FEMB_CUDA(h, cudaMemsetAsync(own_mask.p, 0, (size_t)n, h->stream));
    FEMB_CUDA(h, cudaMemcpyAsync(own_mask.p, h->free_mask.p, (size_t)n_own, cudaMemcpyDeviceToDevice, h->stream));
  }
  *n_found = 0;
  if (nfree <= 0) return FEMB_OK;
  const int k = (int)std::min<int64_t>(o.k, nfree);
  int method = FEMB_SOLVER_PCG;           // static solver behind the shift-invert operator
  if (dist) method = FEMB_SOLVER_PCG;      // row-block partition: distributed PCG
  else if (h->sym.is_chain) method = FEMB_SOLVER_CHAIN;
  else if (n <= 2048) method = FEMB_SOLVER_DENSE;
  // Block size: with a factorisation behind K^-1 extra right-hand sides are almost free, so 4; with
  // PCG every right-hand side costs ~7,000 iterations and the total number of solves grows with the
  // block (1M-DOF frame, 20 modes: 63 / 79 / 116 solves at block 1 / 2 / 4 -> 17.7 / 18.x / 24.2 s), so
  // the default is 2 — the smallest block that still carries the double eigenvalues of symmetric
  // sections and structures in the basis.
  const int block_default = (method == FEMB_SOLVER_PCG) ? 2 : 4;
  const int nb = (int)std::min<int64_t>(4, std::max<int64_t>(1, std::min<int64_t>(o.block > 0 ? o.block : block_default, nfree)));
  int rc = FEMB_OK;
  if (method == FEMB_SOLVER_CHAIN) rc = chain_factor(h);
  else if (method == FEMB_SOLVER_DENSE) rc = dense_factor(h);
  if (rc) return rc;
  femb_solve_opts so;
  std::memset(&so, 0, sizeof(so));
  so.method = FEMB_SOLVER_PCG; so.max_iter = 200000; so.check_every = 50;
  so.precond = (o.precond == FEMB_PRECOND_TWO_LEVEL || o.precond == FEMB_PRECOND_LINES || o.precond == FEMB_PRECOND_AUTO) ? o.precond : FEMB_PRECOND_JACOBI;
  // inner solves 1000x tighter than the wanted pencil residual: measured at 1M DOF, 20 modes, rtol 1e-8 —
  // inner 1e-10 stalls at a pencil residual of 7.7e-8, inner 1e-9 at 3.2e-6 (and both take longer)
  so.rtol = std::min(1e-11, o.rtol * 1e-3);
  so.op = o.op;

  const int mmax = (int)std::min<int64_t>(nfree, std::max(3 * k + 12, 40) + nb);   // basis capacity
  const int keep = (int)std::min<int64_t>(nfree, k + 2 * nb);                      // thick-restart size
  ModalWs ws;
  ws.h = h; ws.n = n; ws.n_own = n_own; ws.mask = dist ? own_mask.p : h->free_mask.p; ws.nchunk = std::max(1, std::min(h->num_sms * 2, (int)((n + kMT - 1) / kMT)));
  DevBuf<double> V, MV, OPV, W, MW, dC, dS, Tmp;
  FEMB_CUDA(h, V.alloc((size_t)mmax * n));
  FEMB_CUDA(h, MV.alloc((size_t)mmax * n));
  FEMB_CUDA(h, OPV.alloc((size_t)mmax * n));
  FEMB_CUDA(h, W.alloc((size_t)nb * n));
  FEMB_CUDA(h, MW.alloc((size_t)nb * n));
  FEMB_CUDA(h, Tmp.alloc((size_t)std::max(std::max(keep, k), 3) * n));
  FEMB_CUDA(h, dS.alloc((size_t)mmax * std::max(keep, k)));
  const int gridn = (int)((n + kMT - 1) / kMT);

  std::vector<double> H((size_t)mmax * mmax, 0.0), coef, evals, evecs, lam_prev, lam, theta;
  std::vector<int> idx;
  int m = 0, steps = 0, restarts = 0;
  uint32_t seed = 1;
  for (int q = 0; q < nb; ++q) {
    random_fill_kernel<<<gridn, kMT, 0, h->stream>>>(W.p + (size_t)q * n, ws.mask, n, seed++);
    h->launches++;
  }
  int ncand = nb;              // candidate vectors currently in W
  double best_res = INFINITY, res_at_restart = INFINITY;
  bool finished = false, ok = false;
  const int max_steps = o.max_iter > 0 ? o.max_iter : 5000;
  while (!finished && steps < max_steps) {
    // ---- 1. M-orthonormalise the candidates against the full basis -> accepted block in W
    int na = 0;
    for (int q = 0; q < ncand && m + na < nfree; ++q) {
      double* wq = W.p + (size_t)q * n;
      double* wa = W.p + (size_t)na * n;      // compacted slot (na <= q)
      double* mwa = MW.p + (size_t)na * n;
      double nrm0 = 0.0, nrm2 = 0.0;
      rc = mass_apply(ws, wq, Tmp.p, 1);
      if (rc) return rc;
      rc = dots(ws, wq, n, 1, Tmp.p, n, 1, &nrm0);
      if (rc) return rc;
      if (wa != wq) FEMB_CUDA(h, cudaMemcpyAsync(wa, wq, n * 8, cudaMemcpyDeviceToDevice, h->stream));
      rc = orth_candidate(ws, V.p, MV.p, m, W.p, MW.p, na, wa, mwa, dC, coef, &nrm2);
      if (rc) return rc;
      if (!(nrm2 > 1e-20 * std::max(nrm0, 1e-300)) || !(nrm2 > 0.0)) {
        // collapsed (invariant subspace): try one fresh random direction
        random_fill_kernel<<<gridn, kMT, 0, h->stream>>>(wa, ws.mask, n, seed++);
        h->launches++;
        rc = orth_candidate(ws, V.p, MV.p, m, W.p, MW.p, na, wa, mwa, dC, coef, &nrm2);
        if (rc) return rc;
        if (!(nrm2 > 0.0)) continue;
      }
      const double inv = 1.0 / std::sqrt(nrm2);
      scale_kernel<<<gridn, kMT, 0, h->stream>>>(wa, inv, n);
      scale_kernel<<<gridn, kMT, 0, h->stream>>>(mwa, inv, n);
      h->launches += 2;
      ++na;
    }
    // ---- 2. thick restart when the block does not fit: keep the best `keep` Ritz vectors
    if (na > 0 && m + na > mmax) {
      const int q = std::min(keep, m);
      std::vector<double> S((size_t)m * q);
      for (int j = 0; j < m; ++j)
        for (int c = 0; c < q; ++c) S[(size_t)j * q + c] = evecs[(size_t)j * m + idx[c]];
      FEMB_CUDA(h, cudaMemcpyAsync(dS.p, S.data(), S.size() * 8, cudaMemcpyHostToDevice, h->stream));
      rc = compress_basis(ws, V.p, m, dS.p, q, Tmp.p);
      if (!rc) rc = compress_basis(ws, MV.p, m, dS.p, q, Tmp.p);
      if (!rc) rc = compress_basis(ws, OPV.p, m, dS.p, q, Tmp.p);
      if (rc) return rc;
      FEMB_CUDA(h, cudaStreamSynchronize(h->stream));
      std::fill(H.begin(), H.end(), 0.0);
      for (int c = 0; c < q; ++c) H[(size_t)c * mmax + c] = evals[idx[c]];
      m = q;
      ++restarts;
      lam_prev.clear();
      if (!(best_res < 0.5 * res_at_restart) && std::isfinite(res_at_restart)) {
        finished = true;  // a whole restart cycle without progress: solver accuracy reached
      }
      res_at_restart = best_res;
    }
    if (finished) break;
    // ---- 3. append, W = K^-1 (M R), H(:, new) = MV^T W
    const int m_old = m;
    if (na > 0) {
      FEMB_CUDA(h, cudaMemcpyAsync(V.p + (size_t)m * n, W.p, (size_t)na * n * 8, cudaMemcpyDeviceToDevice, h->stream));
      FEMB_CUDA(h, cudaMemcpyAsync(MV.p + (size_t)m * n, MW.p, (size_t)na * n * 8, cudaMemcpyDeviceToDevice, h->stream));
      m += na;
      rc = solve_block(ws, method, so, MV.p + (size_t)m_old * n, W.p, na, st);
      if (rc) return rc;
      ++steps;
      FEMB_CUDA(h, cudaMemcpyAsync(OPV.p + (size_t)m_old * n, W.p, (size_t)na * n * 8, cudaMemcpyDeviceToDevice, h->stream));
      std::vector<double> hc((size_t)m * na);
      rc = dots(ws, MV.p, n, m, W.p, n, na, hc.data());
      if (rc) return rc;
      for (int i = 0; i < m; ++i)
        for (int q = 0; q < na; ++q) {
          H[(size_t)i * mmax + (m_old + q)] = hc[(size_t)i * na + q];
          H[(size_t)(m_old + q) * mmax + i] = hc[(size_t)i * na + q];
        }
    }
    ncand = na;
    // ---- 4. Rayleigh-Ritz
    std::vector<double> Hm((size_t)m * m);
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < m; ++j) Hm[(size_t)i * m + j] = 0.5 * (H[(size_t)i * mmax + j] + H[(size_t)j * mmax + i]);
    jacobi_eigh(Hm, m, evals, evecs);
    idx.resize(m);
    for (int i = 0; i < m; ++i) idx[i] = i;
    std::sort(idx.begin(), idx.end(), [&](int a, int b) { return evals[a] > evals[b]; });  // theta descending
    const int kk = std::min(k, m);
    lam.resize(kk); theta.resize(kk);
    for (int c = 0; c < kk; ++c) { theta[c] = evals[idx[c]]; lam[c] = theta[c] > 0.0 ? 1.0 / theta[c] : INFINITY; }
    bool stable = (int)lam_prev.size() == kk && kk == k;
    for (int c = 0; stable && c < kk; ++c) stable = std::fabs(lam[c] - lam_prev[c]) <= 1e-3 * o.rtol * std::fabs(lam[c]);
    lam_prev = lam;
    const bool full = (m + nb > mmax) || (m >= nfree) || (na == 0);
    if (stable || full) {
      // ---- 5. residuals of the k wanted pairs (phi = V s, M phi = MV s, K^-1 M phi = OPV s):
      //   pencil residual        ||K phi - lambda M phi||_2 / ||K phi||_2     (one masked SpMV), and
      //   shift-invert residual  ||K^-1 M phi - theta phi||_M / theta.
      // A pair is accepted when the pencil residual is <= rtol, or — for operators so
      // ill-conditioned that K phi cannot be evaluated to that accuracy (beam chains:
      // cond(K) * eps > rtol) — when the shift-invert residual is <= 1e-3 * rtol.
      std::vector<double> S((size_t)m * kk);
      for (int j = 0; j < m; ++j)
        for (int c = 0; c < kk; ++c) S[(size_t)j * kk + c] = evecs[(size_t)j * m + idx[c]];
      double worst = 0.0;
      double* y = Tmp.p;
      double* my = Tmp.p + (size_t)n;
      double* ky = Tmp.p + (size_t)2 * n;
      for (int c = 0; c < kk; ++c) {
        std::vector<double> s1(m);
        for (int j = 0; j < m; ++j) s1[j] = -S[(size_t)j * kk + c];
        FEMB_CUDA(h, cudaMemsetAsync(y, 0, n * 8, h->stream));
        FEMB_CUDA(h, cudaMemsetAsync(ky, 0, n * 8, h->stream));
        rc = axpy_block(ws, V.p, n, m, s1, y, n, 1, dC);        // y = V s
        if (rc) return rc;
        rc = axpy_block(ws, OPV.p, n, m, s1, ky, n, 1, dC);     // ky = K^-1 M V s
        if (rc) return rc;
        std::vector<double> th1(1, theta[c]);
        rc = axpy_block(ws, y, n, 1, th1, ky, n, 1, dC);        // ky -= theta y
        if (rc) return rc;
        rc = mass_apply(ws, ky, my, 1);
        if (rc) return rc;
        double rn = 0.0;
        rc = dots(ws, ky, n, 1, my, n, 1, &rn);
        if (rc) return rc;
        const double si = std::sqrt(std::max(rn, 0.0)) / std::fabs(theta[c]);
        double res = si <= 1e-3 * o.rtol ? std::min(si, o.rtol) : INFINITY;
        if (!(res <= o.rtol) && si <= 1e2 * o.rtol) {           // close: evaluate the pencil residual
          FEMB_CUDA(h, cudaMemsetAsync(my, 0, n * 8, h->stream));
          rc = axpy_block(ws, MV.p, n, m, s1, my, n, 1, dC);    // my = M V s
          if (rc) return rc;
          if (dist_active(h)) {
            FEMB_CUDA(h, cudaMemsetAsync(ky, 0, n * 8, h->stream));
            rc = dist_spmv_masked(h, y, ky);                    // fills y's ghost tail; y is rebuilt per pair
          } else {
            rc = launch_spmv(h, y, ky, true, nullptr);          // ky = K_ff y (fixed rows: y = 0)
          }
          if (rc) return rc;
          st->spmv_launches++;
          double nk = 0.0, nr = 0.0;
          rc = dots(ws, ky, n, 1, ky, n, 1, &nk);
          if (rc) return rc;
          std::vector<double> l1(1, lam[c]);
          rc = axpy_block(ws, my, n, 1, l1, ky, n, 1, dC);      // ky -= lambda my
          if (rc) return rc;
          rc = dots(ws, ky, n, 1, ky, n, 1, &nr);
          if (rc) return rc;
          res = std::sqrt(std::max(nr, 0.0) / std::max(nk, 1e-300));
        } else if (!(res <= o.rtol)) {
          res = si;
        }
        worst = std::max(worst, res);
      }
      best_res = std::min(best_res, worst);
      st->rel_residual = worst;
      if (worst <= o.rtol || m >= nfree || na == 0) { finished = true; ok = true; }
    }
    if (finished) break;
    for (int q = ncand; q < nb; ++q) {   // pad the candidate block with random directions
      random_fill_kernel<<<gridn, kMT, 0, h->stream>>>(W.p + (size_t)q * n, ws.mask, n, seed++);
      h->launches++;
    }
    ncand = nb;
  }
  // stagnation at the accuracy of the inner solver (ill-conditioned chains): accepted only up to the tolerance the
  // caller allows explicitly (femb_eig_opts.accept_rtol); strict by default
  if (!ok && o.accept_rtol > o.rtol && best_res <= o.accept_rtol) ok = true;
  const int kk = std::min(k, m);
  st->converged = (st->rel_residual <= o.rtol) ? 1 : 0;
  st->method_used = method;
  if (method != FEMB_SOLVER_PCG) st->iterations = steps;
  st->spmv_timed = restarts;
  if (!ok) return fail(h, FEMB_ERR_NOT_CONVERGED, "modal solver did not converge");
  // ---- outputs: eigenvalues > lambda_min (BeamSolver.py:448), M-normalised shapes, zeros on fixed DOFs
  std::vector<double> S((size_t)m * kk);
  for (int j = 0; j < m; ++j)
    for (int c = 0; c < kk; ++c) S[(size_t)j * kk + c] = evecs[(size_t)j * m + idx[c]];
  FEMB_CUDA(h, cudaMemcpyAsync(dS.p, S.data(), S.size() * 8, cudaMemcpyHostToDevice, h->stream));
  dim3 rg((unsigned)gridn, (kk + 7) / 8);
  ritz_vectors_kernel<<<rg, kMT, 0, h->stream>>>(V.p, n, m, dS.p, kk, Tmp.p, n, n);
  h->launches++;
  FEMB_CUDA(h, cudaGetLastError());
  int nout = 0;
  for (int c = 0; c < kk; ++c) {
    if (!(lam[c] > o.lambda_min) || !std::isfinite(lam[c])) continue;
    lambda_out[nout] = lam[c];
    if (phi_out) FEMB_CUDA(h, download(phi_out + (size_t)nout * n_own, Tmp.p + (size_t)c * n, n_own * 8, h->stream));
    ++nout;
  }
  FEMB_CUDA(h, cudaStreamSynchronize(h->stream));
  *n_found = nout;
  return FEMB_OK;
}

}  // namespace femb
