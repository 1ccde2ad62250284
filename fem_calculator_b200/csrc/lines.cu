// Line preconditioner for the matrix-free frame PCG (FEMB_PRECOND_LINES; AUTO picks it for large frames
// whose members form lines).
//
// Why: with the rigid-body coarse space (twolevel.cu) the static solve K_ff u = f (BeamSolver.py:417) still
// needs 1,324 iterations at 1M DOF.  The modes left are rows of collinear members moving along their own
// axis: such a motion costs only the bending energy of the crossing members (12 EI / L^3 per crossing,
// BeamSolver.py:649-652) while the diagonal carries the axial stiffness EA / L (BeamSolver.py:653) — the
// classic anisotropy that point smoothers cannot resolve.  So
//     M^-1 r = omega D^-1 r + sum_lines Q_a (Q_a^T A Q_a)^-1 Q_a^T r + sum_families P_f (P_f^T A P_f)^-1 P_f^T r
//   * Q_a: the axial DOFs of member line a (one per node: the translation along the line's end-to-end
//     direction, rows of fixed DOFs zeroed).  Q_a^T A Q_a is tridiagonal (only consecutive nodes of a line
//     share a member): factored once per assembled K, solved every iteration by one warp per line with two
//     first-order linear recurrences evaluated as warp scans (forward elimination, back substitution);
//   * P_f: one axial translation per BUNDLE of neighbouring lines of family f (coarse.cpp); the Galerkin
//     matrices (a few hundred unknowns per family) are inverted explicitly on the DMMA Cholesky kernels
//     (direct.cu) and applied as dense products.
// CPU study behind the choice (scipy, tests/prototypes/line_precond_study.py): 56x56x54 lattice, rtol 1e-12:
// Jacobi 6,931 iterations, rigid-body two-level 1,324, this form 199 with 2,296 coarse unknowns.
//
// Iteration (Chronopoulos-Gear, linked reductions as in pcg_common.cuh — no float atomics, fixed-order
// sums, bit-reproducible):  operator (ebe.cu) -> ln_update (p, q, x, r; ||r||^2) -> ln_solve (line solves,
// bundle residuals) -> ln_coarse (dense products) -> ln_prolong (z = M^-1 r; (r, z)).
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "pcg_common.cuh"

namespace femb {

constexpr int kLnThreads = 128;        // ln_solve / setup CTAs: 4 warps = 4 lines at a time
constexpr int kLnVecThreads = 256;     // update / prolong kernels
constexpr int kLnCoarseWarps = 8;
constexpr double kLnRidge = 1e-10;     // relative ridge on the diagonal of the bundle Galerkin matrices
constexpr int kLnTargetPerFamily = 768;

struct LnDev {
  const int32_t* line_ptr;
  const int32_t* line_bundle;
  const int32_t* bundle_ptr;
  const int32_t* ent_node;
  const int32_t* ent_blk_diag;
  const int32_t* ent_blk_next;
  const int32_t* node_bundle;   // (F, N)
  const int32_t* bundle_ids;    // coarse index of every bundle_ptr range (null: identity; row-block partition: the local subset)
  const double* node_dir;       // (F, N, 3) caller-provided unit line directions (null: end-to-end from the coordinates)
  int32_t range_off[kLnMaxFam + 1];   // bundle_ptr ranges of each family
  int32_t coarse_blk_off[kLnMaxFam + 1];   // CTAs of ln_coarse_kernel per family (kLnCoarseWarps rows each)
  double* ent_w;                // (n_entries, 3) masked line direction at the entry's node
  double* node_w;               // (F, N, 3)      the same, indexed by (family, node); zero where the node has no line
  double* fac;                  // (n_entries, 3) {1/delta, forward coefficient, backward coefficient}
  double* yl;                   // (F, N) line-solve amplitude of the node's line entry
  double* rb;                   // (n_coarse) bundle residuals
  double* yb;                   // (n_coarse) coarse solution
  const double* inv;            // per-family inverses, family f at inv_off[f], leading dimension fam_pad[f]
  int64_t inv_off[kLnMaxFam];
  int32_t fam_off[kLnMaxFam + 1];
  int32_t fam_pad[kLnMaxFam];
  int32_t n_lines, n_coarse, n_nodes;
  double omega;
};

__device__ __forceinline__ int ln_family_of(const LnDev& T, int c) {
  int f = 0;
#pragma unroll
  for (int k = 1; k < kLnMaxFam; ++k) f += (c >= T.fam_off[k]) ? 1 : 0;
  return f;
}

// ---- numeric setup ----------------------------------------------------------------------------------
// one warp per line: end-to-end unit direction, masked per entry -> ent_w, node_w
__global__ void __launch_bounds__(kLnThreads)
ln_direction_kernel(const LnDev T, const double* __restrict__ xyz, const uint8_t* __restrict__ free_mask) {
  const int line = blockIdx.x * (kLnThreads / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (line >= T.n_lines) return;
  const int lo = T.line_ptr[line], hi = T.line_ptr[line + 1];
  const int f = ln_family_of(T, T.line_bundle[line]);
  const double* pa = xyz + 3 * (size_t)T.ent_node[lo];
  const double* pb = xyz + 3 * (size_t)T.ent_node[hi - 1];
  double d[3] = {pb[0] - pa[0], pb[1] - pa[1], pb[2] - pa[2]};
  const double n2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
  const double inv = n2 > 0.0 ? 1.0 / sqrt(n2) : 0.0;
  for (int e = lo + lane; e < hi; e += 32) {
    const int node = T.ent_node[e];
    const uint8_t* fm = free_mask + 6 * (size_t)node;
    const size_t fn = (size_t)f * T.n_nodes + node;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double dir = T.node_dir ? T.node_dir[3 * fn + c] : d[c] * inv;
      const double w = fm[c] ? dir : 0.0;
      T.ent_w[3 * (size_t)e + c] = w;
      T.node_w[3 * fn + c] = w;
    }
  }
}

// row-block partition: masked direction of EVERY local node that lies on a line (ghosts included — they are the
// columns of the Galerkin rows this rank owns), from the directions the caller computed on the global mesh
__global__ void ln_node_w_kernel(const LnDev T, const uint8_t* __restrict__ free_mask) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)kLnMaxFam * T.n_nodes) return;
  const int node = (int)(t % T.n_nodes);
  const uint8_t* fm = free_mask + 6 * (size_t)node;
  const bool on = T.node_bundle[t] >= 0;
#pragma unroll
  for (int c = 0; c < 3; ++c) T.node_w[3 * t + c] = (on && fm[c]) ? T.node_dir[3 * t + c] : 0.0;
}

__device__ __forceinline__ double ln_quad3(const double* __restrict__ kb, const double* wi, const double* wj) {
  double s = 0.0;
#pragma unroll
  for (int a = 0; a < 3; ++a) s += wi[a] * (kb[a * 6] * wj[0] + kb[a * 6 + 1] * wj[1] + kb[a * 6 + 2] * wj[2]);
  return s;
}

// one warp per line: the tridiagonal Q_a^T A Q_a (diagonal a_k, coupling b_k between entries k and k + 1) from
// the assembled blocks, then its LDL^T factorisation as the coefficients of the two recurrences
//     forward  y_k = r_k / delta_k + f_k y_{k-1},   f_k = -b_{k-1} / delta_k
//     backward x_k = y_k + c_k x_{k+1},             c_k = -b_k / delta_k
__global__ void __launch_bounds__(kLnThreads)
ln_tridiag_kernel(const LnDev T, const double* __restrict__ Kvals) {
  __shared__ double s_a[kLnThreads / 32][kLnMaxLen], s_b[kLnThreads / 32][kLnMaxLen];
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int line = blockIdx.x * (kLnThreads / 32) + wl;
  if (line >= T.n_lines) return;
  const int lo = T.line_ptr[line], len = T.line_ptr[line + 1] - lo;
  for (int k = lane; k < len; k += 32) {
    const int e = lo + k;
    double wi[3], wj[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int c = 0; c < 3; ++c) wi[c] = T.ent_w[3 * (size_t)e + c];
    s_a[wl][k] = ln_quad3(Kvals + 36 * (size_t)T.ent_blk_diag[e], wi, wi);
    const int bn = T.ent_blk_next[e];
    double b = 0.0;
    if (bn >= 0 && k + 1 < len) {
#pragma unroll
      for (int c = 0; c < 3; ++c) wj[c] = T.ent_w[3 * (size_t)(e + 1) + c];
      b = ln_quad3(Kvals + 36 * (size_t)bn, wi, wj);
    }
    s_b[wl][k] = b;
  }
  __syncwarp();
  if (lane == 0) {
    double dprev = 1.0, bprev = 0.0;
    for (int k = 0; k < len; ++k) {
      const double a = s_a[wl][k];
      double delta = a - (k > 0 ? bprev * bprev / dprev : 0.0);
      double bk = bprev;
      if (!(delta > 0.0)) { delta = a > 0.0 ? a : 1.0; bk = 0.0; }      // fixed node / lost positivity: decouple
      const double id = 1.0 / delta;
      double* fc = T.fac + 3 * (size_t)(lo + k);
      fc[0] = id;
      fc[1] = -bk * id;
      fc[2] = 0.0;
      if (k > 0) T.fac[3 * (size_t)(lo + k - 1) + 2] = -bk / dprev;
      dprev = delta;
      bprev = s_b[wl][k];
    }
  }
}

// Galerkin matrix of family f: row A = bundle, K_f[A][B] = sum over entries i of A and stored blocks (i, j) with
// j on a line of the same family in bundle B of w_i^T K_ij[0:3, 0:3] w_j.  One CTA per bundle; the entries of a
// bundle are contiguous.  Groups of 16 entries x 8 block slots produce (value, column) items in shared memory;
// thread t then adds the items whose column is t (mod 128) in item order — one owner per column, fixed order, no
// atomics.  The row lives in shared memory and is written once into the contiguous matrix G.
constexpr int kLnGalEntries = 16, kLnGalSlots = 8;
__global__ void __launch_bounds__(kLnThreads)
ln_galerkin_kernel(const LnDev T, int f, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                   const double* __restrict__ Kvals, double* __restrict__ G, int64_t ld) {
  extern __shared__ double sh[];
  const int nf = T.fam_off[f + 1] - T.fam_off[f];
  double* row = sh;                                     // (nf)
  __shared__ double s_val[kLnThreads];
  __shared__ int s_col[kLnThreads];
  const int rg = T.range_off[f] + blockIdx.x;           // bundle_ptr range
  const int c = T.bundle_ids ? T.bundle_ids[rg] : rg;   // its coarse index
  const int A = c - T.fam_off[f];                       // row inside the family's matrix
  const int e0 = T.line_ptr[T.bundle_ptr[rg]], e1 = T.line_ptr[T.bundle_ptr[rg + 1]];
  for (int k = threadIdx.x; k < nf; k += kLnThreads) row[k] = 0.0;
  __syncthreads();
  const int te = threadIdx.x / kLnGalSlots, ts = threadIdx.x % kLnGalSlots;
  for (int g = e0; g < e1; g += kLnGalEntries) {
    const int e = g + te;
    int i = -1, b0 = 0, deg = 0;
    double wi[3] = {0.0, 0.0, 0.0};
    if (e < e1) {
      i = T.ent_node[e];
      b0 = rowptr[i]; deg = rowptr[i + 1] - b0;
#pragma unroll
      for (int k = 0; k < 3; ++k) wi[k] = T.ent_w[3 * (size_t)e + k];
    }
    for (int s0 = 0;; s0 += kLnGalSlots) {
      const int s = s0 + ts;
      if (!__syncthreads_or(s < deg)) break;      // uniform: every thread's slot is past its row
      double v = 0.0;
      int col = -1;
      if (s < deg) {
        const int b = b0 + s;
        const int j = colidx[b];
        const int cj = T.node_bundle[(size_t)f * T.n_nodes + j];
        if (cj >= 0) {
          const double* wj = T.node_w + ((size_t)f * T.n_nodes + j) * 3;
          const double w3[3] = {wj[0], wj[1], wj[2]};
          v = ln_quad3(Kvals + 36 * (size_t)b, wi, w3);
          col = cj - T.fam_off[f];
        }
      }
      s_val[threadIdx.x] = v;
      s_col[threadIdx.x] = col;
      __syncthreads();
      for (int k = 0; k < kLnThreads; ++k) {
        const int ck = s_col[k];
        if (ck >= 0 && (ck % kLnThreads) == (int)threadIdx.x) row[ck] += s_val[k];
      }
      __syncthreads();
    }
  }
  for (int k = threadIdx.x; k < nf; k += kLnThreads) G[(size_t)A * ld + k] = row[k];
}

// augmented matrix [[G, .], [I, 0]] of coarse_invert from the (n_pad x n_pad) Galerkin matrix G (on a row-block
// partition: the sum over ranks): relative ridge on the diagonal (a bundle without a free axial DOF gets the
// identity), identity on the padding and in the lower left block; everything else was zeroed by the caller
__global__ void ln_aug_fill_kernel(const double* __restrict__ G, double* __restrict__ aug, int64_t n, int64_t n_pad) {
  const int64_t i = blockIdx.y, j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_pad) return;
  const int64_t ld = 2 * n_pad;
  double v = (i < n && j < n) ? G[i * n_pad + j] : 0.0;
  if (i == j) v = (i < n && v > 0.0) ? v * (1.0 + kLnRidge) : 1.0;
  aug[i * ld + j] = v;
  if (i == j) aug[(n_pad + i) * ld + j] = 1.0;
}

// ---- iteration kernels --------------------------------------------------------------------------------
// init: x = 0, r = b, p = q = 0; publishes ||b||^2 into buffer 0 (gamma_0 follows from ln_prolong with wr = 0)
__global__ void __launch_bounds__(kLnVecThreads)
ln_init_kernel(const double* __restrict__ b, double* __restrict__ x, double* __restrict__ r, double* __restrict__ p,
               double* __restrict__ q, int64_t n, const PcgLink L) {
  __shared__ double s_part[kLnVecThreads / 32];
  double v[1] = {0.0};
  for (int64_t g = (int64_t)blockIdx.x * kLnVecThreads + threadIdx.x; g < n; g += (int64_t)gridDim.x * kLnVecThreads) {
    const double bg = b[g];
    x[g] = 0.0; r[g] = bg; p[g] = 0.0; q[g] = 0.0;
    v[0] += bg * bg;
  }
  block_sum_all<kLnVecThreads, 1>(v, s_part);
  if (threadIdx.x == 0) L.upd_partials[L.pstride + blockIdx.x] = v[0];
  if (blockIdx.x == 0 && threadIdx.x == 0) { L.flags[Flag::DONE] = 0; L.flags[Flag::ITERS] = 0; }
}

// update(it): consumes delta (operator) and gamma (ln_prolong of the previous iteration / init) like
// pcg_update_linked_kernel; p = z + beta p, q = s + beta q, x += alpha p, r -= alpha q; publishes ||r||^2 into
// buffer (it + 1) & 1
__global__ void __launch_bounds__(kLnVecThreads)
ln_update_kernel(const double* __restrict__ z, const double* __restrict__ s, double* __restrict__ p, double* __restrict__ q,
                 double* __restrict__ x, double* __restrict__ r, int64_t n, const PcgLink L) {
  __shared__ double s_part[2 * kLnVecThreads / 32];
  if (L.flags[Flag::DONE]) return;
  const int rd = L.it & 1, wr = rd ^ 1;
  double tot[2] = {0.0, 0.0};
  {
    const double* pu = L.upd_partials + (size_t)rd * 2 * L.pstride;
    for (int i = threadIdx.x; i < L.n_op; i += kLnVecThreads) tot[0] += __ldcg(L.op_partials + i);
    for (int i = threadIdx.x; i < L.n_upd; i += kLnVecThreads) tot[1] += __ldcg(pu + i);
  }
  block_sum_all<kLnVecThreads, 2>(tot, s_part);
  const double delta = tot[0], gamma = tot[1];
  const bool first = (L.it == 0);
  const double beta = first ? 0.0 : gamma / L.scal[Scal::RZ0 + rd];
  const double den = first ? delta : delta - beta * gamma / L.scal[Scal::ALPHA + rd];
  const bool bad = !(den > 0.0);           // K_ff (or the preconditioner) not positive definite along p
  const double alpha = bad ? 0.0 : gamma / den;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    L.scal[Scal::RZ0 + wr] = gamma;
    L.scal[Scal::ALPHA + wr] = alpha;
    L.scal[Scal::PQ] = delta;
    if (bad) L.flags[Flag::DONE] = 2;
  }
  if (bad) return;
  double v[1] = {0.0};
  const int64_t n2 = n >> 1;                // ndof = 6 n_nodes is even
  const double2* z2 = reinterpret_cast<const double2*>(z);
  const double2* s2 = reinterpret_cast<const double2*>(s);
  double2* p2 = reinterpret_cast<double2*>(p);
  double2* q2 = reinterpret_cast<double2*>(q);
  double2* x2 = reinterpret_cast<double2*>(x);
  double2* r2 = reinterpret_cast<double2*>(r);
#pragma unroll 2
  for (int64_t i = (int64_t)blockIdx.x * kLnVecThreads + threadIdx.x; i < n2; i += (int64_t)gridDim.x * kLnVecThreads) {
    const double2 zv = z2[i], sv = s2[i];
    double2 pv = p2[i], qv = q2[i], xv = x2[i], rv = r2[i];
    pv.x = zv.x + beta * pv.x; pv.y = zv.y + beta * pv.y;
    qv.x = sv.x + beta * qv.x; qv.y = sv.y + beta * qv.y;
    xv.x += alpha * pv.x; xv.y += alpha * pv.y;
    rv.x -= alpha * qv.x; rv.y -= alpha * qv.y;
    p2[i] = pv; q2[i] = qv; x2[i] = xv; r2[i] = rv;
    v[0] += rv.x * rv.x; v[0] += rv.y * rv.y;
  }
  block_sum_all<kLnVecThreads, 1>(v, s_part);
  if (threadIdx.x == 0) L.upd_partials[(size_t)wr * 2 * L.pstride + L.pstride + blockIdx.x] = v[0];
}

// One CTA per bundle, one warp per line (lines w, w + 4, .. of the bundle): axial residuals a_k = w_k . r_k,
// the line solve (two scans), the line's amplitude per node -> yl, and the bundle residual rb = sum of the
// axial residuals of its lines (per-warp sums in line order, then warp order: fixed).
// A lane holds CH = 4 consecutive entries of the line; affine maps y -> F y + G compose across lanes with a
// Hillis-Steele scan.
__device__ __forceinline__ void ln_scan_affine_up(double& F, double& G, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const double Fp = __shfl_up_sync(0xffffffffu, F, d), Gp = __shfl_up_sync(0xffffffffu, G, d);
    if (lane >= d) { G = fma(F, Gp, G); F *= Fp; }
  }
}
__device__ __forceinline__ void ln_scan_affine_down(double& F, double& G, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const double Fp = __shfl_down_sync(0xffffffffu, F, d), Gp = __shfl_down_sync(0xffffffffu, G, d);
    if (lane + d < 32) { G = fma(F, Gp, G); F *= Fp; }
  }
}

// DIST (row-block partition): the lines are the pieces inside the rank's slab and rb holds the rank's PARTIAL bundle
// residuals; the last CTA (ticket) stores them into every rank's mail area and releases the sequence flag there.
template <bool DIST>
__global__ void __launch_bounds__(kLnThreads)
ln_solve_kernel(const LnDev T, const double* __restrict__ r, const int* __restrict__ flags, const P2PDev* __restrict__ p2p) {
  constexpr int CH = kLnMaxLen / 32;
  __shared__ double s_sum[kLnThreads / 32];
  if (flags[Flag::DONE]) return;
  const int rg = blockIdx.x;
  const int c = T.bundle_ids ? T.bundle_ids[rg] : rg;
  const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int f = ln_family_of(T, c);
  const int l0 = T.bundle_ptr[rg], l1 = T.bundle_ptr[rg + 1];
  double wsum = 0.0;
  for (int line = l0 + wl; line < l1; line += kLnThreads / 32) {
    const int lo = T.line_ptr[line], len = T.line_ptr[line + 1] - lo;
    const int per = (len + 31) >> 5;               // entries per lane (<= CH), consecutive
    const int k0 = lane * per;
    double a[CH], id[CH], ff[CH], cc[CH];
    int node[CH];
    double lsum = 0.0;
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      const int k = k0 + j;
      a[j] = 0.0; id[j] = 0.0; ff[j] = 0.0; cc[j] = 0.0; node[j] = -1;
      if (j < per && k < len) {
        const size_t e = (size_t)(lo + k);
        node[j] = T.ent_node[e];
        const double* w = T.ent_w + 3 * e;
        const double* rn = r + 6 * (size_t)node[j];
        a[j] = w[0] * rn[0] + w[1] * rn[1] + w[2] * rn[2];
        const double* fc = T.fac + 3 * e;
        id[j] = fc[0]; ff[j] = fc[1]; cc[j] = fc[2];
        lsum += a[j];
      }
    }
    // forward: y_k = a_k / delta_k + f_k y_{k-1}
    double F = 1.0, G = 0.0;
#pragma unroll
    for (int j = 0; j < CH; ++j)
      if (j < per) { G = fma(ff[j], G, a[j] * id[j]); F *= ff[j]; }
    ln_scan_affine_up(F, G, lane);
    double yin = __shfl_up_sync(0xffffffffu, G, 1);
    if (lane == 0) yin = 0.0;
    double y[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      y[j] = 0.0;
      if (j < per) { yin = fma(ff[j], yin, a[j] * id[j]); y[j] = yin; }
    }
    // backward: x_k = y_k + c_k x_{k+1}
    F = 1.0; G = 0.0;
#pragma unroll
    for (int j = CH - 1; j >= 0; --j)
      if (j < per) { G = fma(cc[j], G, y[j]); F *= cc[j]; }
    ln_scan_affine_down(F, G, lane);
    double xin = __shfl_down_sync(0xffffffffu, G, 1);
    if (lane == 31) xin = 0.0;
#pragma unroll
    for (int j = CH - 1; j >= 0; --j)
      if (j < per) {
        xin = fma(cc[j], xin, y[j]);
        if (node[j] >= 0) T.yl[(size_t)f * T.n_nodes + node[j]] = xin;
      }
    lsum = warp_sum(lsum);
    wsum += lsum;                                   // valid in lane 0
  }
  if (lane == 0) s_sum[wl] = wsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kLnThreads / 32; ++w) t += s_sum[w];
    T.rb[c] = t;
  }
  if constexpr (DIST) {
    __shared__ int s_last;
    if (threadIdx.x == 0) {
      __threadfence();
      const int tk = atomicAdd(p2p->ticket2, 1);
      s_last = (tk == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const long long seq = p2p->base[0] + flags[Flag::ITERS] + 1;
    const int par = (int)(seq & 1);
    for (int pr = 0; pr < p2p->world; ++pr) {
      double* dst = p2p->peer_rbmail[pr] + (size_t)(p2p->rank * 2 + par) * kLnMaxCoarse;
      for (int k = threadIdx.x; k < T.n_coarse; k += kLnThreads) dst[k] = __ldcg(T.rb + k);
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < p2p->world) st_release_sys(p2p->peer_rbflag[threadIdx.x] + p2p->rank, seq);
    if (threadIdx.x == 0) *p2p->ticket2 = 0;
  }
}

// yb = blockdiag(K_f^-1) rb: one warp per coarse row, the CTA's rows belong to one family (coarse_blk_off).
// DIST: the family's bundle residuals are first summed over the ranks' mail slots in rank order (identical on every
// rank) into shared memory, after waiting for every rank's sequence flag.
template <bool DIST>
__global__ void __launch_bounds__(kLnCoarseWarps * 32)
ln_coarse_kernel(const LnDev T, const int* __restrict__ flags, const P2PDev* __restrict__ p2p) {
  __shared__ double s_rb[DIST ? 1024 : 1];
  if (flags[Flag::DONE]) return;
  int f = 0;
#pragma unroll
  for (int k = 1; k < kLnMaxFam; ++k) f += ((int)blockIdx.x >= T.coarse_blk_off[k]) ? 1 : 0;
  const int off = T.fam_off[f], nf = T.fam_off[f + 1] - off;
  const int lane = threadIdx.x & 31;
  const int row = ((int)blockIdx.x - T.coarse_blk_off[f]) * kLnCoarseWarps + (threadIdx.x >> 5);
  if constexpr (DIST) {
    const long long seq = p2p->base[0] + flags[Flag::ITERS] + 1;
    const int par = (int)(seq & 1);
    if ((int)threadIdx.x < p2p->world) {
      long long spins = 0;
      while (ld_acquire_sys(p2p->my_rbflag + threadIdx.x) != seq) {
        if (++spins > kSpinLimit) { const_cast<int*>(flags)[Flag::DONE] = 4; break; }
      }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < nf; k += kLnCoarseWarps * 32) {
      double t = 0.0;
      for (int pr = 0; pr < p2p->world; ++pr)
        t += *reinterpret_cast<const volatile double*>(p2p->my_rbmail + (size_t)(pr * 2 + par) * kLnMaxCoarse + off + k);
      s_rb[k] = t;
    }
    __syncthreads();
  }
  if (row >= nf) return;
  const double* inv_row = T.inv + T.inv_off[f] + (size_t)row * T.fam_pad[f];
  const double* rb = T.rb + off;
  double acc = 0.0;
#pragma unroll 4
  for (int k = lane; k < nf; k += 32) acc = fma(__ldg(inv_row + k), DIST ? s_rb[k] : __ldcg(rb + k), acc);
  acc = warp_sum(acc);
  if (lane == 0) T.yb[off + row] = acc;
}

// z = omega D^-1 r + sum_f w_f (yl_f + yb[bundle_f]) on the translations of every node; publishes (r, z)
__global__ void __launch_bounds__(kLnVecThreads)
ln_prolong_kernel(const LnDev T, const double* __restrict__ dinv, const double* __restrict__ r, double* __restrict__ z,
                  int wr, const PcgLink L) {
  __shared__ double s_part[kLnVecThreads / 32];
  if (L.flags[Flag::DONE]) return;
  double g[1] = {0.0};
  for (int node = blockIdx.x * kLnVecThreads + threadIdx.x; node < T.n_nodes; node += gridDim.x * kLnVecThreads) {
    const double2* r2 = reinterpret_cast<const double2*>(r + 6 * (size_t)node);
    const double2* d2 = reinterpret_cast<const double2*>(dinv + 6 * (size_t)node);
    const double2 ra = r2[0], rb2 = r2[1], rc = r2[2];
    const double2 da = __ldg(d2), db = __ldg(d2 + 1), dc = __ldg(d2 + 2);
    double zt[3] = {T.omega * da.x * ra.x, T.omega * da.y * ra.y, T.omega * db.x * rb2.x};
#pragma unroll
    for (int f = 0; f < kLnMaxFam; ++f) {
      const size_t fn = (size_t)f * T.n_nodes + node;
      const int cb = __ldg(T.node_bundle + fn);
      if (cb >= 0) {
        const double amp = T.yl[fn] + __ldcg(T.yb + cb);
        const double* w = T.node_w + 3 * fn;
        zt[0] = fma(__ldg(w), amp, zt[0]); zt[1] = fma(__ldg(w + 1), amp, zt[1]); zt[2] = fma(__ldg(w + 2), amp, zt[2]);
      }
    }
    const double z3 = T.omega * db.y * rb2.y, z4 = T.omega * dc.x * rc.x, z5 = T.omega * dc.y * rc.y;
    double2* zo = reinterpret_cast<double2*>(z + 6 * (size_t)node);
    zo[0] = make_double2(zt[0], zt[1]); zo[1] = make_double2(zt[2], z3); zo[2] = make_double2(z4, z5);
    g[0] += ra.x * zt[0] + ra.y * zt[1] + rb2.x * zt[2] + rb2.y * z3 + rc.x * z4 + rc.y * z5;
  }
  block_sum_all<kLnVecThreads, 1>(g, s_part);
  if (threadIdx.x == 0) L.upd_partials[(size_t)wr * 2 * L.pstride + blockIdx.x] = g[0];
}

// ---- row-block partition (dist.cu): the same iteration with the scalars travelling through the peers' mailboxes ----
// update(it) over the owned rows: waits for every rank's {delta, gamma, ||r||^2} (posted by the operator kernels,
// added in rank order: identical on every rank), convergence decision, p, q, x, r; the local ||r||^2 goes to
// red[RR], the iteration counter advances (last CTA of the ordered reduction).
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
dist_ln_update_kernel(const double* __restrict__ z, const double* __restrict__ s, double* __restrict__ p, double* __restrict__ q,
                      double* __restrict__ x, double* __restrict__ r, int64_t n, int first, double rtol, double* partials,
                      int pstride, double* red, int* flags, const P2PDev* __restrict__ p2p) {
  __shared__ double s_glob[3];
  if (flags[Flag::DONE]) return;
  if (threadIdx.x < 32) {
    const long long seq = p2p->base[0] + flags[Flag::ITERS] + 1;
    const int lane = threadIdx.x;
    double v0 = 0.0, v1 = 0.0, v2 = 0.0;
    bool lost = false;
    if (lane < p2p->world) {
      const MailSlot* src = p2p->my_mail + (lane * 2 + (int)(seq & 1));
      long long spins = 0;
      while (ld_acquire_sys(&src->seq) != seq) {
        if (++spins > kSpinLimit) { lost = true; break; }
      }
      v0 = *reinterpret_cast<const volatile double*>(&src->v[0]);
      v1 = *reinterpret_cast<const volatile double*>(&src->v[1]);
      v2 = *reinterpret_cast<const volatile double*>(&src->v[2]);
    }
    lost = __any_sync(0xffffffffu, lost);
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    for (int pr = 0; pr < p2p->world; ++pr) {
      a0 += __shfl_sync(0xffffffffu, v0, pr); a1 += __shfl_sync(0xffffffffu, v1, pr); a2 += __shfl_sync(0xffffffffu, v2, pr);
    }
    if (lane == 0) { s_glob[0] = a0; s_glob[1] = a1; s_glob[2] = lost ? -1.0 : a2; if (lost) flags[Flag::DONE] = 4; }
  }
  __syncthreads();
  const double delta = s_glob[0], gamma = s_glob[1], rr = s_glob[2];
  if (rr < 0.0) return;
  const double tol2 = first ? rtol * rtol * rr : red[Red::TOL2];
  if (first ? (rr == 0.0) : (rr <= tol2)) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      red[Red::RRFINAL] = rr;
      if (first) red[Red::BB] = rr;
      flags[Flag::DONE] = 1;
    }
    return;
  }
  const double beta = first ? 0.0 : gamma / red[Red::GPREV];
  const double den = first ? delta : delta - beta * gamma / red[Red::ALPHA];
  const bool bad = !(den > 0.0);
  const double alpha = bad ? 0.0 : gamma / den;
  double rr_new = 0.0;
  for (int64_t g = (int64_t)blockIdx.x * THREADS + threadIdx.x; g < n; g += (int64_t)gridDim.x * THREADS) {
    const double pg = z[g] + beta * p[g];
    const double qg = s[g] + beta * q[g];
    const double rg = r[g] - alpha * qg;
    p[g] = pg; q[g] = qg;
    x[g] += alpha * pg;
    r[g] = rg;
    rr_new += rg * rg;
  }
  double mine[1], tot[1];
  mine[0] = rr_new;
  if (grid_reduce<THREADS, 1>(mine, partials, pstride, flags + Flag::TICKET1, tot)) {
    if (threadIdx.x == 0) {
      red[Red::GPREV] = gamma;
      red[Red::ALPHA] = alpha;
      if (first) { red[Red::TOL2] = tol2; red[Red::BB] = rr; }
      red[Red::RRFINAL] = rr;
      red[Red::RR] = tot[0];
      flags[Flag::ITERS] = flags[Flag::ITERS] + 1;
      if (bad) flags[Flag::DONE] = 2;
    }
  }
}

// prolongation over the owned nodes: z = omega D^-1 r + sum_f w_f (yl_f + yb[bundle_f]); the local (r, z) goes to
// red[GAMMA].  Phase 1 handles the nodes a neighbour needs: their z goes straight into the neighbours' ghost tails
// (remote stores) and the last CTA to finish the phase (ticket) stores the few nodes with several destinations and
// releases the halo flag of the next operator launch at every neighbour — while phase 2, the interior, is still
// running, so the flag has landed long before the neighbour's operator kernel looks for it.
__device__ __forceinline__ double ln_prolong_node(const LnDev& T, const double* __restrict__ dinv, const double* __restrict__ r,
                                                  double* __restrict__ z, int node, double* zt) {
  const double2* r2 = reinterpret_cast<const double2*>(r + 6 * (size_t)node);
  const double2* d2 = reinterpret_cast<const double2*>(dinv + 6 * (size_t)node);
  const double2 ra = r2[0], rb2 = r2[1], rc = r2[2];
  const double2 da = __ldg(d2), db = __ldg(d2 + 1), dc = __ldg(d2 + 2);
  zt[0] = T.omega * da.x * ra.x; zt[1] = T.omega * da.y * ra.y; zt[2] = T.omega * db.x * rb2.x;
  zt[3] = T.omega * db.y * rb2.y; zt[4] = T.omega * dc.x * rc.x; zt[5] = T.omega * dc.y * rc.y;
#pragma unroll
  for (int f = 0; f < kLnMaxFam; ++f) {
    const size_t fn = (size_t)f * T.n_nodes + node;
    const int cb = __ldg(T.node_bundle + fn);
    if (cb >= 0) {
      const double amp = T.yl[fn] + __ldcg(T.yb + cb);
      const double* w = T.node_w + 3 * fn;
      zt[0] = fma(__ldg(w), amp, zt[0]); zt[1] = fma(__ldg(w + 1), amp, zt[1]); zt[2] = fma(__ldg(w + 2), amp, zt[2]);
    }
  }
  double2* zo = reinterpret_cast<double2*>(z + 6 * (size_t)node);
  zo[0] = make_double2(zt[0], zt[1]); zo[1] = make_double2(zt[2], zt[3]); zo[2] = make_double2(zt[4], zt[5]);
  return ra.x * zt[0] + ra.y * zt[1] + rb2.x * zt[2] + rb2.y * zt[3] + rc.x * zt[4] + rc.y * zt[5];
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS)
dist_ln_prolong_kernel(const LnDev T, const double* __restrict__ dinv, const double* __restrict__ r, double* __restrict__ z,
                       int n_owned, double* partials, int pstride, double* red, int* flags, const P2PDev* __restrict__ p2p) {
  __shared__ int s_last;
  if (flags[Flag::DONE]) return;
  double g = 0.0;
  // phase 1: boundary nodes
  bool pushed = false;
  for (int k = blockIdx.x * THREADS + threadIdx.x; k < p2p->n_bnd; k += gridDim.x * THREADS) {
    const int node = __ldg(p2p->bnd_nodes + k);
    double zt[6];
    g += ln_prolong_node(T, dinv, r, z, node, zt);
    const int sl = __ldg(p2p->send_slot + node);
    double* dst = p2p->peer_z[sl >> 28] + (size_t)(sl & 0xFFFFFFF) * 6;
#pragma unroll
    for (int c = 0; c < 6; ++c) dst[c] = zt[c];
    pushed = true;
  }
  if (pushed) __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const int tk = atomicAdd(p2p->ticket3, 1);
    s_last = (tk == (int)gridDim.x - 1);
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    for (int e = threadIdx.x; e < p2p->n_extra * 6; e += THREADS) {
      const int i = e / 6, c = e - i * 6;
      const int node = p2p->extra[2 * i], sl = p2p->extra[2 * i + 1];
      p2p->peer_z[sl >> 28][(size_t)(sl & 0xFFFFFFF) * 6 + c] = __ldcg(z + (size_t)node * 6 + c);
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < p2p->n_nbr) {
      const long long seq = p2p->base[0] + flags[Flag::ITERS] + 1;
      st_release_sys(p2p->peer_halo_flag[threadIdx.x] + p2p->rank, seq);
    }
    if (threadIdx.x == 0) *p2p->ticket3 = 0;
  }
  // phase 2: interior nodes
  for (int node = blockIdx.x * THREADS + threadIdx.x; node < n_owned; node += gridDim.x * THREADS) {
    if (__ldg(p2p->send_slot + node) >= 0) continue;
    double zt[6];
    g += ln_prolong_node(T, dinv, r, z, node, zt);
  }
  double mine[1], tot[1];
  mine[0] = g;
  if (grid_reduce<THREADS, 1>(mine, partials, pstride, flags + Flag::TICKET2, tot)) {
    if (threadIdx.x == 0) red[Red::GAMMA] = tot[0];
  }
}

// one CTA: the operator-side decision alone (host poll), as pcg_decide_linked_kernel
__global__ void __launch_bounds__(128)
ln_decide_kernel(const PcgLink L) {
  __shared__ double s_part[2 * 128 / 32];
  if (L.flags[Flag::DONE]) return;
  pcg_link_decide<128>(L, s_part);
}

// ---- host side ------------------------------------------------------------------------------------------
static int ln_target_per_family() {
  if (const char* e = getenv("FEMB_LINE_BUNDLES")) { const int v = atoi(e); if (v > 0) return std::min(v, 1024); }
  return kLnTargetPerFamily;
}

// device copies of the line tables in h->line_sym + the buffers sized by them
static int upload_line_tables(femb_handle* h) {
  LineSym& S = h->line_sym;
  h->line_failed = false;
  h->ln_n_ranges = (int32_t)S.bundle_ptr.size() - 1;
  // bundle_ptr ranges per family (ranges are sorted by coarse index, families are contiguous index ranges)
  for (int f = 0; f <= kLnMaxFam; ++f) {
    if (S.bundle_ids.empty()) h->ln_range_off[f] = S.fam_off[f];
    else h->ln_range_off[f] = (int32_t)(std::lower_bound(S.bundle_ids.begin(), S.bundle_ids.end(), S.fam_off[f]) - S.bundle_ids.begin());
  }
  if (S.n_lines > 0) {
    FEMB_CUDA(h, upload(h->ln_line_ptr, S.line_ptr, h->stream));
    FEMB_CUDA(h, upload(h->ln_line_bundle, S.line_bundle, h->stream));
    FEMB_CUDA(h, upload(h->ln_bundle_ptr, S.bundle_ptr, h->stream));
    FEMB_CUDA(h, upload(h->ln_ent_node, S.ent_node, h->stream));
    FEMB_CUDA(h, upload(h->ln_ent_blk_diag, S.ent_blk_diag, h->stream));
    FEMB_CUDA(h, upload(h->ln_ent_blk_next, S.ent_blk_next, h->stream));
    FEMB_CUDA(h, upload(h->ln_node_bundle, S.node_bundle, h->stream));
    if (!S.bundle_ids.empty()) FEMB_CUDA(h, upload(h->ln_bundle_ids, S.bundle_ids, h->stream));
    else h->ln_bundle_ids.release();
    if (!S.node_dir.empty()) FEMB_CUDA(h, upload(h->ln_node_dir, S.node_dir, h->stream));
    else h->ln_node_dir.release();
    FEMB_CUDA(h, h->ln_ent_w.alloc((size_t)S.n_entries * 3));
    FEMB_CUDA(h, h->ln_fac.alloc((size_t)S.n_entries * 3));
    FEMB_CUDA(h, h->ln_node_w.alloc((size_t)kLnMaxFam * h->n_nodes * 3));
    FEMB_CUDA(h, h->ln_yl.alloc((size_t)kLnMaxFam * h->n_nodes));
    FEMB_CUDA(h, h->ln_rb.alloc((size_t)S.n_coarse));
    FEMB_CUDA(h, h->ln_yb.alloc((size_t)S.n_coarse));
    int64_t off = 0, gmax = 1;
    for (int f = 0; f < kLnMaxFam; ++f) {
      const int nf = S.fam_off[f + 1] - S.fam_off[f];
      h->ln_fam_pad[f] = (nf + 63) / 64 * 64;
      h->ln_inv_off[f] = off;
      off += (int64_t)h->ln_fam_pad[f] * h->ln_fam_pad[f];
      gmax = std::max<int64_t>(gmax, (int64_t)h->ln_fam_pad[f] * h->ln_fam_pad[f]);
    }
    FEMB_CUDA(h, h->ln_inv.alloc((size_t)std::max<int64_t>(off, 1)));
    FEMB_CUDA(h, h->ln_gal.alloc((size_t)gmax));
    FEMB_CUDA(h, cudaMemsetAsync(h->ln_node_w.p, 0, h->ln_node_w.bytes(), h->stream));
    FEMB_CUDA(h, cudaMemsetAsync(h->ln_yl.p, 0, h->ln_yl.bytes(), h->stream));
    FEMB_CUDA(h, cudaMemsetAsync(h->ln_rb.p, 0, h->ln_rb.bytes(), h->stream));
    FEMB_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  // the big host tables are not needed again
  std::vector<int32_t>().swap(S.ent_node); std::vector<int32_t>().swap(S.ent_blk_diag);
  std::vector<int32_t>().swap(S.ent_blk_next); std::vector<int32_t>().swap(S.node_bundle);
  std::vector<int32_t>().swap(S.node_ent); std::vector<int32_t>().swap(S.node_line);
  std::vector<double>().swap(S.node_dir);
  h->line_sym_ok = true;
  h->line_num_ok = false;
  return FEMB_OK;
}

static int ensure_line_symbolic(femb_handle* h) {
  if (h->line_sym_ok) return FEMB_OK;
  if (dist_active(h)) return FEMB_OK;              // row-block partition: the tables come from femb_dist_set_lines
  std::vector<double> hx((size_t)h->n_nodes * 3);
  FEMB_CUDA(h, download(hx.data(), h->xyz.p, hx.size() * 8, h->stream));
  FEMB_CUDA(h, cudaStreamSynchronize(h->stream));
  build_line_symbolic(h->sym, h->h_conn.data(), hx.data(), ln_target_per_family(), h->line_sym);
  h->line_dist = false;
  return upload_line_tables(h);
}

// row-block partition: the rank's line tables from the global symbolic phase (femb_dist_set_lines)
int dist_set_lines(femb_handle* h, int32_t n_coarse, const int32_t* fam_off, const int32_t* node_bundle,
                   const int32_t* node_line, const int32_t* node_pos, const double* node_dir) {
  if (!h->have_symbolic || h->n_owned_nodes <= 0) return fail(h, FEMB_ERR_ARG, "call femb_assemble and femb_dist_set_halo before femb_dist_set_lines");
  if (n_coarse <= 0 || n_coarse > kLnMaxCoarse) return fail(h, FEMB_ERR_ARG, "coarse dimension of the line preconditioner out of range");
  for (int f = 0; f < kLnMaxFam; ++f)
    if (fam_off[f + 1] - fam_off[f] > 1024 || fam_off[f + 1] < fam_off[f]) return fail(h, FEMB_ERR_ARG, "at most 1024 bundles per family");
  build_line_symbolic_local(h->sym, h->n_owned_nodes, n_coarse, fam_off, node_bundle, node_line, node_pos, node_dir, h->line_sym);
  h->line_dist = true;
  return upload_line_tables(h);
}

static LnDev ln_dev(const femb_handle* h) {
  LnDev T;
  const LineSym& S = h->line_sym;
  T.line_ptr = h->ln_line_ptr.p; T.line_bundle = h->ln_line_bundle.p; T.bundle_ptr = h->ln_bundle_ptr.p;
  T.ent_node = h->ln_ent_node.p; T.ent_blk_diag = h->ln_ent_blk_diag.p; T.ent_blk_next = h->ln_ent_blk_next.p;
  T.node_bundle = h->ln_node_bundle.p; T.bundle_ids = h->ln_bundle_ids.p; T.node_dir = h->ln_node_dir.p;
  T.ent_w = h->ln_ent_w.p; T.node_w = h->ln_node_w.p; T.fac = h->ln_fac.p;
  T.yl = h->ln_yl.p; T.rb = h->ln_rb.p; T.yb = h->ln_yb.p; T.inv = h->ln_inv.p;
  for (int f = 0; f < kLnMaxFam; ++f) { T.inv_off[f] = h->ln_inv_off[f]; T.fam_pad[f] = h->ln_fam_pad[f]; }
  T.coarse_blk_off[0] = 0;
  for (int f = 0; f <= kLnMaxFam; ++f) {
    T.fam_off[f] = S.fam_off[f];
    T.range_off[f] = h->ln_range_off[f];
    if (f > 0) T.coarse_blk_off[f] = T.coarse_blk_off[f - 1] + (S.fam_off[f] - S.fam_off[f - 1] + kLnCoarseWarps - 1) / kLnCoarseWarps;
  }
  T.n_lines = S.n_lines; T.n_coarse = S.n_coarse; T.n_nodes = (int32_t)h->n_nodes;
  T.omega = 1.0;
  if (const char* e = getenv("FEMB_LN_OMEGA")) { const double v = atof(e); if (v > 0.0) T.omega = v; }
  return T;
}

// directions, line factors, bundle Galerkin matrices and their inverses for the current K and BC mask
static int ensure_line_numeric(femb_handle* h) {
  int rc = ensure_line_symbolic(h);
  if (rc) return rc;
  if (!h->line_sym_ok) { h->line_failed = true; return FEMB_OK; }
  if (h->line_num_ok || h->line_failed) return FEMB_OK;
  const LineSym& S = h->line_sym;
  const bool dist = h->line_dist;
  if (S.n_coarse == 0 || (!dist && S.n_lines == 0)) { h->line_failed = true; return FEMB_OK; }
  const LnDev T = ln_dev(h);
  const int wpc = kLnThreads / 32;
  const int grid_lines = (S.n_lines + wpc - 1) / wpc;
  const bool trace = getenv("FEMB_TRACE") != nullptr;
  cudaEvent_t te[3] = {nullptr, nullptr, nullptr};
  if (trace) { for (auto& e : te) cudaEventCreate(&e); cudaEventRecord(te[0], h->stream); }
  if (dist) {
    const int64_t tot = (int64_t)kLnMaxFam * h->n_nodes;
    ln_node_w_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, h->stream>>>(T, h->free_mask.p);
    h->launches++;
  }
  if (S.n_lines > 0) {
    ln_direction_kernel<<<grid_lines, kLnThreads, 0, h->stream>>>(T, h->xyz.p, h->free_mask.p);
    ln_tridiag_kernel<<<grid_lines, kLnThreads, 0, h->stream>>>(T, h->Kvals.p);
    h->launches += 2;
  }
  FEMB_CUDA(h, cudaGetLastError());
  if (trace) cudaEventRecord(te[1], h->stream);
  bool all_ok = true;
  for (int f = 0; f < kLnMaxFam && all_ok; ++f) {
    const int nf = S.fam_off[f + 1] - S.fam_off[f];
    if (nf == 0) continue;
    const int64_t n_pad = h->ln_fam_pad[f], m = 2 * n_pad;
    FEMB_CUDA(h, h->coarse_aug.ensure((size_t)m * m));
    FEMB_CUDA(h, cudaMemsetAsync(h->coarse_aug.p, 0, (size_t)m * m * sizeof(double), h->stream));
    FEMB_CUDA(h, cudaMemsetAsync(h->ln_gal.p, 0, (size_t)n_pad * n_pad * sizeof(double), h->stream));
    const int n_rg = T.range_off[f + 1] - T.range_off[f];
    if (n_rg > 0) {
      const size_t smem = (size_t)nf * sizeof(double);
      ln_galerkin_kernel<<<n_rg, kLnThreads, smem, h->stream>>>(T, f, h->rowptr.p, h->colidx.p, h->Kvals.p, h->ln_gal.p, n_pad);
      h->launches++;
    }
    if (dist) {                                     // rows of the other ranks' bundles: sum over the ranks
      rc = dist_allreduce(h, h->ln_gal.p, (int)(n_pad * n_pad));
      if (rc) return rc;
    }
    ln_aug_fill_kernel<<<dim3((unsigned)((n_pad + 255) / 256), (unsigned)n_pad), 256, 0, h->stream>>>(h->ln_gal.p, h->coarse_aug.p, nf, n_pad);
    h->launches++;
    FEMB_CUDA(h, cudaGetLastError());
    bool ok = false;
    rc = coarse_invert(h, h->coarse_aug.p, n_pad, h->ln_inv.p + h->ln_inv_off[f], &ok);
    if (rc) return rc;
    all_ok = all_ok && ok;
  }
  if (trace) {
    cudaEventRecord(te[2], h->stream);
    cudaEventSynchronize(te[2]);
    float a = 0.f, b = 0.f;
    cudaEventElapsedTime(&a, te[0], te[1]);
    cudaEventElapsedTime(&b, te[1], te[2]);
    for (auto& e : te) cudaEventDestroy(e);
    fprintf(stderr, "[femb trace] line preconditioner setup: %d lines, %lld entries, coarse dim %d (%d/%d/%d), coverage %.3f; "
            "line factors %.3f ms, Galerkin + inversion %.3f ms\n", S.n_lines, (long long)S.n_entries, S.n_coarse,
            S.fam_off[1] - S.fam_off[0], S.fam_off[2] - S.fam_off[1], S.fam_off[3] - S.fam_off[2], S.coverage, a, b);
  }
  h->line_num_ok = all_ok;
  h->line_failed = !all_ok;
  return FEMB_OK;
}

constexpr int64_t kLnAutoNodes = 50000;     // FEMB_PRECOND_AUTO: below this the setup costs more than it saves
constexpr double kLnAutoCoverage = 0.5;     // ... and at least half of the nodes must lie on a member line

// does this solve use the line preconditioner?  (AUTO decides on the member-line coverage, which needs the
// symbolic phase: it is built on first use and kept with the topology)
bool lines_applicable(femb_handle* h, const femb_solve_opts& o) {
  if (h->bs != 6 || !ebe_selected(h, o.op)) return false;
  if (o.precond != FEMB_PRECOND_LINES && !(o.precond == FEMB_PRECOND_AUTO && h->n_nodes >= kLnAutoNodes)) return false;
  if (ensure_line_symbolic(h) != FEMB_OK) return false;
  if (h->line_sym.n_lines == 0) return false;
  return o.precond == FEMB_PRECOND_LINES || h->line_sym.coverage >= kLnAutoCoverage;
}

int pcg_lines(femb_handle* h, const femb_solve_opts& o, const double* d_b, femb_stats* st) {
  const int pstride = h->num_sms * 8;
  int rc = setup_precond_public(h, FEMB_PRECOND_JACOBI);
  if (rc) return rc;
  rc = ensure_line_numeric(h);
  if (rc) return rc;
  if (!h->line_num_ok) {
    // a bundle matrix could not be factored: plain Jacobi (stats.coarse_dim stays 0)
    femb_solve_opts oj = o;
    oj.precond = FEMB_PRECOND_JACOBI;
    return pcg_solve_rhs(h, oj, d_b, st);
  }
  const int64_t n = h->ndof;
  FEMB_CUDA(h, h->fpartials.ensure((size_t)pstride * 6));
  FEMB_CUDA(h, cudaMemsetAsync(h->flags.p, 0, sizeof(int32_t) * Flag::COUNT, h->stream));
  FEMB_CUDA(h, cudaMemsetAsync(h->scal.p, 0, sizeof(double) * Scal::COUNT, h->stream));
  const LnDev T = ln_dev(h);
  // update and prolong publish into the same partial arrays: one grid size for both
  const int grid_v = std::max(1, std::min((int)((h->n_nodes + kLnVecThreads - 1) / kLnVecThreads), h->num_sms * 4));
  const int grid_c = T.coarse_blk_off[kLnMaxFam];
  const int grid_s = h->ln_n_ranges;
  PcgLink L;
  L.upd_partials = h->fpartials.p; L.op_partials = h->fpartials.p + (size_t)4 * pstride;
  L.scal = h->scal.p; L.flags = h->flags.p;
  L.n_upd = grid_v; L.n_op = ebe_grid(h, 1, h->n_nodes); L.pstride = pstride;
  L.it = 0; L.max_iter = o.max_iter; L.rtol = o.rtol;
  auto precond = [&](int wr_buf) {
    ln_solve_kernel<false><<<grid_s, kLnThreads, 0, h->stream>>>(T, h->r.p, h->flags.p, nullptr);
    ln_coarse_kernel<false><<<grid_c, kLnCoarseWarps * 32, 0, h->stream>>>(T, h->flags.p, nullptr);
    ln_prolong_kernel<<<grid_v, kLnVecThreads, 0, h->stream>>>(T, h->Dinv.p, h->r.p, h->z.p, wr_buf, L);
    h->launches += 3;
  };
  ln_init_kernel<<<grid_v, kLnVecThreads, 0, h->stream>>>(d_b, h->x.p, h->r.p, h->p.p, h->q.p, n, L);
  h->launches++;
  precond(0);
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
      ln_update_kernel<<<grid_v, kLnVecThreads, 0, h->stream>>>(h->z.p, h->s.p, h->p.p, h->q.p, h->x.p, h->r.p, n, L);
      h->launches++;
      precond((it & 1) ^ 1);
      if (timed) { cudaEventRecord(e2, h->stream); evs.push_back(e0); evs.push_back(e1); evs.push_back(e2); }
    }
    L.it = it;
    ln_decide_kernel<<<1, 128, 0, h->stream>>>(L);
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
    st->coarse_dim = T.n_coarse;
    st->precond_used = FEMB_PRECOND_LINES;
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

// ---- row-block partition: the pieces dist.cu's solver loop launches --------------------------------------
// usable when the rank's line tables are set, the exchange runs through peer memory with the exchanges fused into
// the kernels, and the matrix-free operator applies
bool dist_lines_applicable(femb_handle* h, const femb_solve_opts& o, bool fused_p2p) {
  if (!(o.precond == FEMB_PRECOND_LINES || o.precond == FEMB_PRECOND_AUTO)) return false;
  return fused_p2p && h->bs == 6 && h->line_sym_ok && h->line_dist && !h->line_failed && ebe_available_dist(h);
}

int dist_lines_setup(femb_handle* h) {
  int rc = ensure_line_numeric(h);
  if (rc) return rc;
  return h->line_num_ok ? FEMB_OK : 1;          // 1: fall back to Jacobi
}

// z = M^-1 r on the owned rows (+ halo push, (r, z) -> red[GAMMA]); three launches on the handle's stream
int dist_lines_precond(femb_handle* h, double* red) {
  const LnDev T = ln_dev(h);
  const int pstride = h->num_sms * 8;
  const P2PDev* pd = reinterpret_cast<const P2PDev*>(h->p2p_dev_copy.p);
  const int n_owned = (int)h->n_owned_nodes;
  const int grid_v = std::max(1, std::min((n_owned + kLnVecThreads - 1) / kLnVecThreads, h->num_sms * 4));
  if (h->ln_n_ranges > 0) ln_solve_kernel<true><<<h->ln_n_ranges, kLnThreads, 0, h->stream>>>(T, h->r.p, h->flags.p, pd);
  else return fail(h, FEMB_ERR_ARG, "line preconditioner: this rank owns no line piece");
  ln_coarse_kernel<true><<<T.coarse_blk_off[kLnMaxFam], kLnCoarseWarps * 32, 0, h->stream>>>(T, h->flags.p, pd);
  dist_ln_prolong_kernel<kLnVecThreads><<<grid_v, kLnVecThreads, 0, h->stream>>>(T, h->Dinv.p, h->r.p, h->z.p, n_owned,
                                                                                   h->partials.p + 2 * pstride, pstride, red, h->flags.p, pd);
  h->launches += 3;
  FEMB_CUDA(h, cudaGetLastError());
  return FEMB_OK;
}

int dist_lines_update(femb_handle* h, int first, double rtol, double* red) {
  const int pstride = h->num_sms * 8;
  const int64_t n = h->n_owned_nodes * 6;
  const P2PDev* pd = reinterpret_cast<const P2PDev*>(h->p2p_dev_copy.p);
  const int grid = vec_grid(h, n, kLnVecThreads);
  dist_ln_update_kernel<kLnVecThreads><<<grid, kLnVecThreads, 0, h->stream>>>(h->z.p, h->s.p, h->p.p, h->q.p, h->x.p, h->r.p, n, first, rtol,
                                                                               h->partials.p + pstride, pstride, red, h->flags.p, pd);
  h->launches++;
  FEMB_CUDA(h, cudaGetLastError());
  return FEMB_OK;
}

int dist_lines_coarse_dim(const femb_handle* h) { return h->line_sym.n_coarse; }

}  // namespace femb
