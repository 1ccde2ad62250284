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
// Iteration (Chronopoulos-Gear; no float atomics, fixed-order sums, bit-reproducible): ONE persistent cooperative
// kernel, ln_pcg_mega_kernel, with five phases per iteration separated by grid barriers — operator (ebe.cuh) ->
// vector update (p, q, x, r; ||r||^2; axial residuals of the line entries) -> line solves + bundle residuals ->
// coarse products -> prolongation (z = M^-1 r; (r, z)).  On a row-block partition the same kernel exchanges halo,
// reduction scalars and bundle residuals with the other GPUs from inside the launch (flag-in-data stores).
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <limits>

#include "common.cuh"
#include "ebe.cuh"
#include "pcg_common.cuh"

namespace femb {

constexpr int kLnThreads = 128;        // ln_solve / setup CTAs: 4 warps = 4 lines at a time
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
  double* ent_w;                // (n_entries, 3) masked line direction at the entry's node
  double* node_w;               // (F, N, 3)      the same, indexed by (family, node); zero where the node has no line
  double* fac;                  // (n_entries, 3) {1/delta, forward coefficient, backward coefficient}
  double* ae;                   // (n_entries) axial residual w . r of every line entry (written by the vector update)
  double* yle;                  // (n_entries) line-solve amplitude of every line entry
  const int32_t* ent_of;        // (F, N) line entry of the node in family f, -1: none
  double* rb;                   // (n_coarse) bundle residuals (row-block partition: this rank's partial sums)
  double* rbt;                  // (n_coarse) bundle residuals the coarse products read (= rb; partition: the sum over ranks)
  const int32_t* rank_mask;     // (n_coarse) partition: bit p set = rank p owns a piece of the bundle (contributes to its residual)
  double* line_sum;             // (n_lines) axial residual sum of every line (persistent kernel: bundles are summed from these)
  int32_t max_len;              // longest line (entries)
  int32_t* bundle_cnt;          // (bundle_ptr ranges) lines of each bundle finished in the current pass (persistent kernel)
  const int32_t* grp_ptr;       // (line groups of the grid + 1) persistent kernel: the lines of every line group ...
  const int4* grp_lines;        // (n_lines) ... {line, first entry, entries, bundle range}, balanced by length on the host (ln_assign_lines)
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
// Line solves: the two first-order recurrences of a line run as scans of affine maps y -> F y + G, composed across
// the lanes of a line group with a Hillis-Steele scan.
template <int LW>
__device__ __forceinline__ void ln_scan_affine_up(double& F, double& G, int lane) {
#pragma unroll
  for (int d = 1; d < LW; d <<= 1) {
    const double Fp = __shfl_up_sync(0xffffffffu, F, d, LW), Gp = __shfl_up_sync(0xffffffffu, G, d, LW);
    if (lane >= d) { G = fma(F, Gp, G); F *= Fp; }
  }
}
template <int LW>
__device__ __forceinline__ void ln_scan_affine_down(double& F, double& G, int lane) {
#pragma unroll
  for (int d = 1; d < LW; d <<= 1) {
    const double Fp = __shfl_down_sync(0xffffffffu, F, d, LW), Gp = __shfl_down_sync(0xffffffffu, G, d, LW);
    if (lane + d < LW) { G = fma(F, Gp, G); F *= Fp; }
  }
}

// persistent kernel: ALL lines dealt round-robin to the line groups (LW lanes) of the whole grid — every group gets
// n_lines / (groups in the grid) lines instead of a CTA walking its bundles one after the other; the per-line
// residual sums go to line_sum and the bundles are summed from them (bundle order, line order) where they are used.
// Lane l of a group holds entries l, l + LW, l + 2 LW, .. (coalesced table reads — with consecutive entries per lane
// the 40 narrow loads of a lane each touched their own sector and the phase was bound by L1 wavefronts); the two
// recurrences run as CH rounds of LW-wide affine scans with the carry handed from round to round.
// Everything the phase touches is indexed by line ENTRY (contiguous per line): the axial residuals ae = w . r were
// scattered there by the vector update (which has r in registers and a thread per node), the amplitudes go to yle and
// are gathered per node by the prolongation — no dependent node-id -> r gather sits on the scans' critical path.
// POST (row-block partition): the group that completes a bundle stores the rank's partial residual straight into every
// rank's flag-in-data slot [this rank][parity][bundle] (lane p -> rank p) instead of into rb.
template <int THREADS, int LW, bool POST>
__device__ __forceinline__ void ln_solve_lines_flat(const LnDev& T, int cta, int ncta,
                                                    const P2PDev* pd = nullptr, long long seq = 0) {
  constexpr int CH = kLnMaxLen / LW;                  // rounds (LW = 16: 8 rounds cover 128 entries)
  constexpr int NG = THREADS / LW;
  const int lane = threadIdx.x & (LW - 1);
  const int gg = cta * NG + threadIdx.x / LW;
  // the group's lines: host-side longest-processing-time assignment (ln_assign_lines): lines differ in length — on a
  // row-block partition the pieces of cut lines are a fraction of the uncut ones — and dealt round-robin the groups
  // that drew the long lines kept everybody else at the barrier
  const int g0 = __ldg(T.grp_ptr + gg), g1 = __ldg(T.grp_ptr + gg + 1);
  int trips = g1 - g0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) trips = max(trips, __shfl_xor_sync(0xffffffffu, trips, o));   // warp-uniform (shuffles)
  for (int it = 0; it < trips; ++it) {
    const bool on = g0 + it < g1;
    const int4 lrec = on ? __ldg(T.grp_lines + g0 + it) : make_int4(0, 0, 0, 0);     // {line, first entry, entries, bundle range}
    const int line = lrec.x, lo = lrec.y, len = lrec.z;
    int nround = (len + LW - 1) / LW;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nround = max(nround, __shfl_xor_sync(0xffffffffu, nround, o));   // warp-uniform
    double g[CH], ff[CH], cc[CH];
    double lsum = 0.0;
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      g[j] = 0.0; ff[j] = 0.0; cc[j] = 0.0;
      const int k = j * LW + lane;
      if (j < nround && k < len) {
        const size_t e = (size_t)(lo + k);
        const double a = __ldcg(T.ae + e);            // written by the vector update of this launch
        const double* fc = T.fac + 3 * e;
        g[j] = a * __ldg(fc); ff[j] = __ldg(fc + 1); cc[j] = __ldg(fc + 2);
        lsum += a;
      }
    }
    // forward: y_k = g_k + f_k y_{k-1}
    double carry = 0.0;
#pragma unroll
    for (int j = 0; j < CH; ++j)
      if (j < nround) {
        double F = ff[j], G = g[j];
        ln_scan_affine_up<LW>(F, G, lane);
        g[j] = fma(F, carry, G);                      // y of this entry
        carry = __shfl_sync(0xffffffffu, g[j], LW - 1, LW);
      }
    // backward: x_k = y_k + c_k x_{k+1}
    carry = 0.0;
#pragma unroll
    for (int j = CH - 1; j >= 0; --j)
      if (j < nround) {
        double F = cc[j], G = g[j];
        ln_scan_affine_down<LW>(F, G, lane);
        const double xk = fma(F, carry, G);
        carry = __shfl_sync(0xffffffffu, xk, 0, LW);
        if (j * LW + lane < len) T.yle[(size_t)lo + j * LW + lane] = xk;
      }
    // line sum in a fixed order: per lane round by round, then the shuffle tree
#pragma unroll
    for (int o = LW / 2; o > 0; o >>= 1) lsum += __shfl_down_sync(0xffffffffu, lsum, o, LW);
    double t_post = 0.0;
    int c_post = -1;
    if (on && lane == 0) {
      // the group that completes a bundle (integer ticket) adds the bundle's line sums in line order: the value does
      // not depend on which group that is
      T.line_sum[line] = lsum;
      __threadfence();
      const int c = __ldg(T.line_bundle + line);
      const int rg = lrec.w;
      const int b0 = __ldg(T.bundle_ptr + rg), b1 = __ldg(T.bundle_ptr + rg + 1);
      if (atomicAdd(T.bundle_cnt + rg, 1) == b1 - b0 - 1) {
        __threadfence();
        double t = 0.0;
        for (int l = b0; l < b1; ++l) t += __ldcg(T.line_sum + l);
        if constexpr (POST) { t_post = t; c_post = c; }
        else T.rb[c] = t;
        T.bundle_cnt[rg] = 0;
      }
    }
    if constexpr (POST) {
      c_post = __shfl_sync(0xffffffffu, c_post, 0, LW);
      t_post = __shfl_sync(0xffffffffu, t_post, 0, LW);
      if (c_post >= 0 && lane < pd->world)
        ll_store(pd->peer_ll_rb[lane] + (size_t)(pd->rank * 2 + (int)(seq & 1)) * kLnMaxCoarse + c_post, t_post, (unsigned)seq);
    }
  }
}

// ---- the whole iteration as ONE persistent kernel (single GPU) --------------------------------------------------
// Five dependent kernels per iteration (operator, update, line solves, coarse products, prolongation) of 10-30 us
// each leave a fifth of the iteration in launch gaps and reduction tails (ncu launch list: 75 us of kernels in a
// 95 us iteration at 1M DOF).  ln_pcg_mega_kernel runs `n_iters` iterations in one cooperative launch: every SM
// keeps its CTAs resident, the five phases are separated by grid-wide barriers (one atomic arrive + a generation
// flag, acquire/release at device scope), and every reduction is a "publish partials, barrier, everybody adds them
// in the same fixed order" — no last-CTA tail, still no float atomics, bit-reproducible.  The host launches again
// every `check_every` iterations and reads the flags in between.
constexpr int kMegaThreads = 128;      // the operator phase needs 128 registers: 4 CTAs of 128 threads per SM

struct MegaArgs {
  LnDev T;
  FrameParams P;
  const int4* pair_rec;
  const int4* node_rec;
  const double4* pair_aux;   // (n_pairs) {1/L, 1/sqrt(cx^2+cy^2), w_z, w_y} of every pair's element (ebe.cuh)
  const uint8_t* free_mask;
  const double* dinv;
  const double* b;
  double *x, *r, *z, *p, *q, *s;
  double* part;          // [3][pstride] published partial sums: delta, ||r||^2, gamma
  double* scal;          // Scal:: slots (carried across launches and mirrored for the host)
  int* flags;
  unsigned int* bar;     // grid barrier words (mega_barrier)
  unsigned long long* phase_ns;   // [8] time spent per phase (CTA 0), accumulated
  unsigned long long* cta_ns;     // FEMB_TRACE: [grid][8] working time per phase of every CTA (null otherwise)
  double* glob;          // [4] DIST: {delta, gamma, ||r||^2} of the world, written by CTA 0 before the barrier
  int64_t n;             // ndof
  int n_nodes, n_ranges, pstride;
  int it0, n_iters, max_iter, init;
  double rtol;
  const P2PDev* p2p;     // row-block partition (DIST): n / n_nodes count the OWNED rows, T.n_nodes all local nodes
};

// Grid barrier words (uint32): [0] arrive counter; generation flag replicated kBarCopies times, copy k at word
// 32 (1 + k) — each on its own 128-byte line.  A CTA polls copy (cta mod kBarCopies): with one shared word the ~600
// spinning CTAs kept one L2 slice busy and the arrive atomics of the stragglers queued behind their reads (measured:
// phase wall minus mean working time per CTA ~5 us per barrier, profiles/r02_persistent_pcg_experiments.log).
constexpr int kBarCopies = 16;
constexpr int kBarWords = 32 * (1 + kBarCopies);

__device__ __forceinline__ void mega_barrier_release(unsigned int* bar) {
  bar[0] = 0;
  __threadfence();
#pragma unroll
  for (int k = 0; k < kBarCopies; ++k) atomicAdd(bar + 32 * (1 + k), 1u);
}

__device__ __forceinline__ void mega_barrier(unsigned int* bar, unsigned int nb) {
  __syncthreads();
  if (threadIdx.x == 0) {
    volatile unsigned int* vgen = bar + 32 * (1 + (blockIdx.x % kBarCopies));
    const unsigned int gen = *vgen;            // read before arriving
    __threadfence();
    if (atomicAdd(bar, 1u) == nb - 1) mega_barrier_release(bar);
    else while (*vgen == gen) { }
    __threadfence();
  }
  __syncthreads();
}

// the same barrier; the LAST CTA to arrive runs `hook` (all its threads) before it releases the others: everything the
// other CTAs published before arriving is visible to the hook, everything the hook writes is visible behind the barrier
template <class F>
__device__ __forceinline__ void mega_barrier_hook(unsigned int* bar, unsigned int nb, int* s_last, F hook) {
  __syncthreads();
  unsigned int gen = 0;
  volatile unsigned int* vgen = bar + 32 * (1 + (blockIdx.x % kBarCopies));
  if (threadIdx.x == 0) {
    gen = *vgen;
    __threadfence();
    *s_last = (atomicAdd(bar, 1u) == nb - 1) ? 1 : 0;
    if (*s_last) __threadfence();
  }
  __syncthreads();
  if (*s_last) {
    hook();
    __syncthreads();
    if (threadIdx.x == 0) mega_barrier_release(bar);
  } else if (threadIdx.x == 0) {
    while (*vgen == gen) { }
  }
  if (threadIdx.x == 0) __threadfence();
  __syncthreads();
}

__device__ __forceinline__ unsigned long long mega_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// fixed-order sum of `cnt` published partials, identical in every thread of every CTA
template <int THREADS>
__device__ __forceinline__ double mega_total(const double* part, int cnt, double* s_part) {
  double v[1] = {0.0};
  for (int i = threadIdx.x; i < cnt; i += THREADS) v[0] += __ldcg(part + i);
  block_sum_all<THREADS, 1>(v, s_part);
  return v[0];
}

// DIST (one rank of a row-block partition, dist.cu): the same kernel on the owned rows; the three exchanges with the
// other GPUs happen INSIDE the launch, through peer memory, as flag-in-data stores (P2PDev in pcg_common.cuh: every
// value is one 16-byte store that carries its own sequence flag, the receiver polls the value's slot — no system
// fence, no separate flag, no extra grid barrier to hand a "it has landed" on):
//   halo        the prolongation handles the nodes a neighbour needs first and stores their z into the neighbour's
//               halo slots; after its interior nodes every CTA unpacks a share of the rank's own halo slots into the
//               ghost tail of z (by then the values have usually landed);
//   scalars     after the operator's barrier CTA 0 stores the rank's {delta, gamma, ||r||^2} into every rank's
//               slots, adds the world's entries in rank order (identical on every rank) and leaves the totals
//               behind the next barrier;
//   coarse      after the line phase CTA p stores the rank's partial bundle residuals into rank p's slots; the next
//               CTAs add the world's partials in rank order as they land.
// Sequence numbers = per-solve base + iteration; waits are bounded (a lost peer sets DONE = 4).
template <bool DIST>
__global__ void __launch_bounds__(kMegaThreads, 4)
ln_pcg_mega_kernel(const MegaArgs A) {
  __shared__ double s_part[2 * kMegaThreads / 32];
  __shared__ int s_last;
  __shared__ double2 s_rt[kMegaThreads / 32][96];
  const LnDev& T = A.T;
  const int cta = blockIdx.x, ncta = gridDim.x;
  const unsigned int nb = gridDim.x;
  double* part_delta = A.part;
  double* part_rr = A.part + A.pstride;
  double* part_gamma = A.part + 2 * (size_t)A.pstride;
  const bool clock = (cta == 0 && threadIdx.x == 0);
  unsigned long long t_last = clock ? mega_now() : 0ull;
  // FEMB_TRACE: every CTA's own working time per phase (barrier waits excluded) -> cta_ns[cta][8]
  unsigned long long t_work = (A.cta_ns && threadIdx.x == 0) ? mega_now() : 0ull;
  auto work_mark = [&](int ph) { if (A.cta_ns && threadIdx.x == 0) A.cta_ns[(size_t)cta * 8 + ph] += mega_now() - t_work; };
  auto work_start = [&]() { if (A.cta_ns && threadIdx.x == 0) t_work = mega_now(); };
  auto lap = [&](int ph) {
    if (clock) { const unsigned long long t = mega_now(); A.phase_ns[ph] += t - t_last; t_last = t; }
  };
  // Vector update p = z + beta p, q = s + beta q, x += alpha p, r -= alpha q (init: x = 0, r = b, p = q = 0), the
  // ||r||^2 partial, and the axial residuals of the new r for the line phase.  A warp takes 32 consecutive nodes = 96
  // double2 per vector (three coalesced loads per lane); the new r is parked in shared memory so that lane n can pick
  // up the three translations of node n and scatter w_f . r to the node's line entries (ae).
  auto update_vectors = [&](bool init, double alpha, double beta) {
    constexpr int NW = kMegaThreads / 32;
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const double2* z2 = reinterpret_cast<const double2*>(A.z);
    const double2* s2 = reinterpret_cast<const double2*>(A.s);
    const double2* b2 = reinterpret_cast<const double2*>(A.b);
    double2* p2 = reinterpret_cast<double2*>(A.p);
    double2* q2 = reinterpret_cast<double2*>(A.q);
    double2* x2 = reinterpret_cast<double2*>(A.x);
    double2* r2 = reinterpret_cast<double2*>(A.r);
    double v[1] = {0.0};
    // equal contiguous node ranges per warp (the phase is bandwidth bound: equal bytes = equal time; chunks of 32 nodes
    // dealt round-robin left 2 or 3 chunks per warp — measured max 15.2 against mean 12.2 us of working time per CTA)
    const int64_t gw = (int64_t)cta * NW + wl, nw_tot = (int64_t)ncta * NW;
    const int64_t n_lo = A.n_nodes * gw / nw_tot, n_hi = A.n_nodes * (gw + 1) / nw_tot;
    for (int64_t n0 = n_lo; n0 < n_hi; n0 += 32) {
      const int64_t i0 = n0 * 3, i_hi = n_hi * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int64_t i = i0 + c * 32 + lane;
        if (i < i_hi) {
          double2 rv;
          if (init) {
            rv = b2[i];
            const double2 zero = make_double2(0.0, 0.0);
            x2[i] = zero; p2[i] = zero; q2[i] = zero; r2[i] = rv;
          } else {
            const double2 zv = z2[i], sv = s2[i];
            double2 pv = p2[i], qv = q2[i], xv = x2[i];
            rv = r2[i];
            pv.x = zv.x + beta * pv.x; pv.y = zv.y + beta * pv.y;
            qv.x = sv.x + beta * qv.x; qv.y = sv.y + beta * qv.y;
            xv.x += alpha * pv.x; xv.y += alpha * pv.y;
            rv.x -= alpha * qv.x; rv.y -= alpha * qv.y;
            p2[i] = pv; q2[i] = qv; x2[i] = xv; r2[i] = rv;
          }
          v[0] += rv.x * rv.x; v[0] += rv.y * rv.y;
          s_rt[wl][c * 32 + lane] = rv;
        }
      }
      __syncwarp();
      const int64_t node = n0 + lane;
      if (node < n_hi) {
        const double2 ra = s_rt[wl][3 * lane], rb2 = s_rt[wl][3 * lane + 1];
#pragma unroll
        for (int f = 0; f < kLnMaxFam; ++f) {
          const size_t fn = (size_t)f * T.n_nodes + node;
          const int e = __ldg(T.ent_of + fn);
          if (e >= 0) {
            const double* w = T.node_w + 3 * fn;
            T.ae[e] = __ldg(w) * ra.x + __ldg(w + 1) * ra.y + __ldg(w + 2) * rb2.x;
          }
        }
      }
      __syncwarp();
    }
    block_sum_all<kMegaThreads, 1>(v, s_part);
    if (threadIdx.x == 0) part_rr[cta] = v[0];
  };
  // one application of the preconditioner: z = M^-1 r, publishes the (r, z) partials.  DIST: `seq` numbers the
  // coarse-residual exchange of this application, seq + 1 the halo the next operator phase waits for.
  auto precond = [&](long long seq) {
    if constexpr (DIST) {
      // bundle residuals of the world: the line groups store this rank's partial sums straight into every rank's
      // flag-in-data slots; right behind its own lines every CTA adds the contributing ranks' slots of a share of the
      // bundles in rank order as they land -> rbt.  No barrier in between: the slots synchronise themselves.
      const P2PDev* pd = A.p2p;
      ln_solve_lines_flat<kMegaThreads, 16, true>(T, cta, ncta, pd, seq);
      const int par = (int)(seq & 1);
      const unsigned flag = (unsigned)seq;
      // (the grid's LAST CTAs take the polls: the first ones hold the longest lines of the length-balanced assignment)
      const bool clk1 = (A.cta_ns && cta == ncta - 2 && threadIdx.x == 0);       // FEMB_TRACE: time the polls of one thread
      const unsigned long long tp0 = clk1 ? mega_now() : 0ull;
      for (int k = (ncta - 1 - cta) * kMegaThreads + threadIdx.x; k < T.n_coarse; k += ncta * kMegaThreads) {
        const int mask = __ldg(T.rank_mask + k);
        double t = 0.0;
        for (int pr = 0; pr < pd->world; ++pr)
          if ((mask >> pr) & 1) {
            double v;
            if (!ll_wait(pd->my_ll_rb + (size_t)(pr * 2 + par) * kLnMaxCoarse + k, flag, v)) A.flags[Flag::DONE] = 4;
            t += v;
          }
        T.rbt[k] = t;
      }
      if (clk1) A.phase_ns[5] += mega_now() - tp0;
    } else {
      ln_solve_lines_flat<kMegaThreads, 16, false>(T, cta, ncta);
    }
    work_mark(2); mega_barrier(A.bar, nb); work_start();
    lap(2);
    {   // coarse products: one warp per row, rows dealt to the grid's warps (partition: only the rows of bundles this
        // rank owns a piece of — the prolongation reads no others)
      const int lane = threadIdx.x & 31;
      const int nwarp = ncta * (kMegaThreads / 32);
      const int n_rows = DIST ? A.n_ranges : T.n_coarse;
      for (int row = cta * (kMegaThreads / 32) + (threadIdx.x >> 5); row < n_rows; row += nwarp) {
        const int c = DIST ? __ldg(T.bundle_ids + row) : row;
        const int f = ln_family_of(T, c);
        const int off = T.fam_off[f], nf = T.fam_off[f + 1] - off;
        const double* inv_row = T.inv + T.inv_off[f] + (size_t)(c - off) * T.fam_pad[f];
        const double* rb = T.rbt + off;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int k = lane;
#pragma unroll 2
        for (; k + 96 < nf; k += 128) {
          const double m0 = __ldg(inv_row + k), m1 = __ldg(inv_row + k + 32), m2 = __ldg(inv_row + k + 64), m3 = __ldg(inv_row + k + 96);
          a0 = fma(m0, __ldcg(rb + k), a0); a1 = fma(m1, __ldcg(rb + k + 32), a1);
          a2 = fma(m2, __ldcg(rb + k + 64), a2); a3 = fma(m3, __ldcg(rb + k + 96), a3);
        }
        for (; k < nf; k += 32) a0 = fma(__ldg(inv_row + k), __ldcg(rb + k), a0);
        const double acc = warp_sum((a0 + a1) + (a2 + a3));
        if (lane == 0) T.yb[c] = acc;
      }
    }
    work_mark(3); mega_barrier(A.bar, nb); work_start();
    lap(3);
    double g = 0.0;
    // equal contiguous node ranges per CTA (same reasoning as in the vector update)
    const int p_lo = (int)((int64_t)A.n_nodes * cta / ncta), p_hi = (int)((int64_t)A.n_nodes * (cta + 1) / ncta);
    auto prolong_node = [&](int node, double* zt) {
      // every load of the node issued before the first dependent one (the yb gather): two latencies per node
      const double2* r2 = reinterpret_cast<const double2*>(A.r + 6 * (size_t)node);
      const double2* d2 = reinterpret_cast<const double2*>(A.dinv + 6 * (size_t)node);
      const double2 ra = r2[0], rb2 = r2[1], rc = r2[2];
      const double2 da = __ldg(d2), db = __ldg(d2 + 1), dc = __ldg(d2 + 2);
      int cb[kLnMaxFam];
      double yl[kLnMaxFam], w[kLnMaxFam][3];
#pragma unroll
      for (int f = 0; f < kLnMaxFam; ++f) {
        const size_t fn = (size_t)f * T.n_nodes + node;
        cb[f] = __ldg(T.node_bundle + fn);
        const int e = __ldg(T.ent_of + fn);
        yl[f] = e >= 0 ? __ldcg(T.yle + e) : 0.0;     // the node's line amplitude in the family
        w[f][0] = __ldg(T.node_w + 3 * fn); w[f][1] = __ldg(T.node_w + 3 * fn + 1); w[f][2] = __ldg(T.node_w + 3 * fn + 2);
      }
      zt[0] = T.omega * da.x * ra.x; zt[1] = T.omega * da.y * ra.y; zt[2] = T.omega * db.x * rb2.x;
#pragma unroll
      for (int f = 0; f < kLnMaxFam; ++f) {
        const double amp = yl[f] + __ldcg(T.yb + max(cb[f], 0));      // w is zero where cb < 0
        zt[0] = fma(w[f][0], amp, zt[0]); zt[1] = fma(w[f][1], amp, zt[1]); zt[2] = fma(w[f][2], amp, zt[2]);
      }
      zt[3] = T.omega * db.y * rb2.y; zt[4] = T.omega * dc.x * rc.x; zt[5] = T.omega * dc.y * rc.y;
      double2* zo = reinterpret_cast<double2*>(A.z + 6 * (size_t)node);
      zo[0] = make_double2(zt[0], zt[1]); zo[1] = make_double2(zt[2], zt[3]); zo[2] = make_double2(zt[4], zt[5]);
      g += ra.x * zt[0] + ra.y * zt[1] + rb2.x * zt[2] + rb2.y * zt[3] + rc.x * zt[4] + rc.y * zt[5];
    };
    if constexpr (DIST) {
      // the nodes a neighbour needs first: their z goes out as flag-in-data stores into the neighbours' halo slots and
      // is on the wire while the interior runs
      const P2PDev* pd = A.p2p;
      const unsigned hflag = (unsigned)(seq + 1);
      // A warp takes 32 consecutive boundary nodes; their z is parked in shared memory so that every store instruction
      // writes 32 CONSECUTIVE 16-byte slots (full 128-byte NVLink writes): with one node per lane (6 slots each, 96 bytes
      // apart) every slot travelled as its own packet and the 145k packets of an interior rank's 24,200 boundary nodes
      // cost ~10 us per iteration on 8 GPUs (profiles/r02_dist_8gpu_phases.log).  Runs of nodes whose destination
      // offsets are consecutive at one neighbour go out this way, anything else (a node with several destinations,
      // a run that changes neighbour) falls back to per-node stores.
      {
        constexpr int NW = kMegaThreads / 32;
        const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
        double* sz = reinterpret_cast<double*>(&s_rt[wl][0]);          // 192 doubles per warp
        for (int k0 = (cta * NW + wl) * 32; k0 < pd->n_bnd; k0 += ncta * NW * 32) {
          const int k = k0 + lane;
          const bool on = k < pd->n_bnd;
          int d0 = 0, d1 = 0, sl0 = -1;
          double zt[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
          if (on) {
            const int node = __ldg(pd->bnd_nodes + k);
            prolong_node(node, zt);
            d0 = __ldg(pd->bnd_dst_ptr + k); d1 = __ldg(pd->bnd_dst_ptr + k + 1);
            sl0 = __ldg(pd->bnd_dst + d0);
          }
          // uniform run: every lane has exactly one destination, same neighbour, consecutive offsets
          const int base = __shfl_sync(0xffffffffu, sl0, 0);
          const bool fits = on ? (d1 - d0 == 1 && sl0 == base + lane) : true;
          const int n_on = min(32, pd->n_bnd - k0);
          if (__all_sync(0xffffffffu, fits) && base >= 0) {
#pragma unroll
            for (int c = 0; c < 6; ++c) sz[lane * 6 + c] = zt[c];
            __syncwarp();
            uint4* dst = pd->peer_ll_halo[base >> 28] + (size_t)(base & 0xFFFFFFF) * 6;
            for (int e = lane; e < n_on * 6; e += 32) ll_store(dst + e, sz[e], hflag);
            __syncwarp();
          } else if (on) {
            for (int d = d0; d < d1; ++d) {
              const int sl = __ldg(pd->bnd_dst + d);
              uint4* dst = pd->peer_ll_halo[sl >> 28] + (size_t)(sl & 0xFFFFFFF) * 6;
#pragma unroll
              for (int c = 0; c < 6; ++c) ll_store(dst + c, zt[c], hflag);
            }
          }
        }
      }
      for (int node = p_lo + threadIdx.x; node < p_hi; node += kMegaThreads) {
        if (__ldg(pd->send_slot + node) >= 0) continue;
        double zt[6];
        prolong_node(node, zt);
      }
    } else {
      for (int node = p_lo + threadIdx.x; node < p_hi; node += kMegaThreads) {
        double zt[6];
        prolong_node(node, zt);
      }
    }
    double v[1] = {g};
    block_sum_all<kMegaThreads, 1>(v, s_part);
    if (threadIdx.x == 0) part_gamma[cta] = v[0];
    if constexpr (DIST) {
      // the neighbours' halo for the next operator phase: unpack the slots into the ghost tail of z as they land (the
      // grid starts from its last CTA: the first ones carry the boundary nodes above)
      const P2PDev* pd = A.p2p;
      const unsigned hflag = (unsigned)(seq + 1);
      const bool clk1 = (A.cta_ns && cta == ncta - 1 && threadIdx.x == 0);
      const unsigned long long tp0 = clk1 ? mega_now() : 0ull;
      for (int kn = 0; kn < pd->n_nbr; ++kn) {
        const long long first = pd->recv_start[kn] * 6, cnt = pd->recv_count[kn] * 6;
        const uint4* src = pd->my_ll_halo + (first - pd->n_owned * 6);
        for (long long e = (long long)(ncta - 1 - cta) * kMegaThreads + threadIdx.x; e < cnt; e += (long long)ncta * kMegaThreads) {
          double v;
          if (!ll_wait(src + e, hflag, v)) A.flags[Flag::DONE] = 4;
          A.z[first + e] = v;
        }
      }
      if (clk1) A.phase_ns[6] += mega_now() - tp0;
    }
    work_mark(4); mega_barrier(A.bar, nb); work_start();
    lap(4);
  };

  double gamma_prev = A.scal[Scal::RZ0], alpha_prev = A.scal[Scal::ALPHA], tol2 = A.scal[Scal::TOL2];
  if (A.init) {
    // x = 0, r = b, p = q = 0; ||b||^2; z_0 = M^-1 r_0
    update_vectors(true, 0.0, 0.0);
    mega_barrier(A.bar, nb);
    precond(DIST ? A.p2p->base[0] : 0);
  }
  int it = A.it0, done = 0;
  double rr = 0.0;
  for (int k = 0; k < A.n_iters; ++k, ++it) {
    // scalars of the state after `it` updates: (r, z) and ||r||^2, same value in every thread of every CTA
    double gamma = mega_total<kMegaThreads>(part_gamma, ncta, s_part);
    rr = mega_total<kMegaThreads>(part_rr, ncta, s_part);
    double delta;
    if constexpr (!DIST) {
      if (it == 0) {
        tol2 = A.rtol * A.rtol * rr;
        if (clock) { A.scal[Scal::BB] = rr; A.scal[Scal::TOL2] = tol2; }
        if (rr == 0.0) done = 1;               // zero load: u = 0 is the answer
      } else if (rr <= tol2) done = 1;
      else if (it >= A.max_iter) done = 3;
      if (done) break;
    } else {
      // (the halo of z has landed: CTA 0 waited for the neighbours' flags before the barrier that closed the prolongation)
    }
    // operator: s = A z, publishes (z, s)
    {
      double v[1];
      v[0] = ebe_nodes_phase<kMegaThreads>(A.P, A.pair_rec, A.node_rec, A.n_nodes, A.free_mask, A.z, A.s, cta, ncta, A.pair_aux);
      block_sum_all<kMegaThreads, 1>(v, s_part);
      if (threadIdx.x == 0) part_delta[cta] = v[0];
    }
    if constexpr (!DIST) {
      work_mark(0); mega_barrier(A.bar, nb); work_start();
      lap(0);
      delta = mega_total<kMegaThreads>(part_delta, ncta, s_part);
    } else {
      // {delta, gamma, ||r||^2} of the world inside the operator's barrier: the last CTA to arrive adds the rank's
      // delta partials, stores the rank's three scalars into every rank's slots (flag-in-data), adds the world's in
      // rank order as they land (identical on every rank) and leaves the totals for everybody behind the barrier
      const P2PDev* pd = A.p2p;
      const long long seq = pd->base[0] + it + 1;
      work_mark(0); mega_barrier_hook(A.bar, nb, &s_last, [&]() {
        const double dl = mega_total<kMegaThreads>(part_delta, ncta, s_part);
        if (threadIdx.x < 32) {
          const int lane = threadIdx.x;
          const int par = (int)(seq & 1);
          const unsigned flag = (unsigned)seq;
          double v0 = 0.0, v1 = 0.0, v2 = 0.0;
          bool ok = true;
          if (lane < pd->world) {
            uint4* dst = pd->peer_ll_scal[lane] + (pd->rank * 2 + par) * 4;
            ll_store(dst, dl, flag); ll_store(dst + 1, gamma, flag); ll_store(dst + 2, rr, flag);
            const uint4* src = pd->my_ll_scal + (lane * 2 + par) * 4;
            ok = ll_wait(src, flag, v0) & ll_wait(src + 1, flag, v1) & ll_wait(src + 2, flag, v2);
          }
          const bool lost = __any_sync(0xffffffffu, !ok) || (*reinterpret_cast<volatile int*>(A.flags + Flag::DONE) == 4);
          double a0 = 0.0, a1 = 0.0, a2 = 0.0;
          for (int pr = 0; pr < pd->world; ++pr) {
            a0 += __shfl_sync(0xffffffffu, v0, pr); a1 += __shfl_sync(0xffffffffu, v1, pr); a2 += __shfl_sync(0xffffffffu, v2, pr);
          }
          if (lane == 0) { A.glob[0] = a0; A.glob[1] = a1; A.glob[2] = lost ? -1.0 : a2; }
        }
      });
      work_start();
      lap(0);
      delta = __ldcg(A.glob); gamma = __ldcg(A.glob + 1); rr = __ldcg(A.glob + 2);
      if (rr < 0.0) { done = 4; break; }          // a peer never answered
      if (it == 0) {
        tol2 = A.rtol * A.rtol * rr;
        if (clock) { A.scal[Scal::BB] = rr; A.scal[Scal::TOL2] = tol2; }
        if (rr == 0.0) done = 1;
      } else if (rr <= tol2) done = 1;
      else if (it >= A.max_iter) done = 3;
      if (done) break;
    }
    const double beta = (it == 0) ? 0.0 : gamma / gamma_prev;
    const double den = (it == 0) ? delta : delta - beta * gamma / alpha_prev;
    if (!(den > 0.0)) { done = 2; break; }     // K_ff (or the preconditioner) not positive definite along p
    const double alpha = gamma / den;
    gamma_prev = gamma; alpha_prev = alpha;
    update_vectors(false, alpha, beta);
    work_mark(1); mega_barrier(A.bar, nb); work_start();
    lap(1);
    precond(DIST ? A.p2p->base[0] + it + 1 : 0);
  }
  if (clock) {
    A.scal[Scal::RZ0] = gamma_prev; A.scal[Scal::ALPHA] = alpha_prev;
    if (done) A.scal[Scal::RR] = rr;
    A.flags[Flag::ITERS] = it;
    if (done) A.flags[Flag::DONE] = done;
  }
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
    FEMB_CUDA(h, h->ln_ae.alloc((size_t)S.n_entries));
    FEMB_CUDA(h, h->ln_yle.alloc((size_t)S.n_entries));
    {
      // entry of every (family, node): families are contiguous coarse-index ranges, lines carry their bundle
      std::vector<int32_t> eo((size_t)kLnMaxFam * h->n_nodes, -1);
      for (int32_t l = 0; l < S.n_lines; ++l) {
        int f = 0;
        for (int k = 1; k < kLnMaxFam; ++k) f += (S.line_bundle[l] >= S.fam_off[k]) ? 1 : 0;
        for (int32_t e = S.line_ptr[l]; e < S.line_ptr[l + 1]; ++e) eo[(size_t)f * h->n_nodes + S.ent_node[e]] = e;
      }
      FEMB_CUDA(h, upload(h->ln_ent_of, eo, h->stream));
      FEMB_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    FEMB_CUDA(h, h->ln_rb.alloc((size_t)S.n_coarse));
    FEMB_CUDA(h, h->ln_rbt.alloc((size_t)S.n_coarse));
    FEMB_CUDA(h, h->ln_yb.alloc((size_t)S.n_coarse));
    FEMB_CUDA(h, h->ln_line_sum.alloc((size_t)S.n_lines));
    FEMB_CUDA(h, h->ln_bundle_cnt.alloc((size_t)std::max<size_t>(1, S.bundle_ptr.size())));
    FEMB_CUDA(h, cudaMemsetAsync(h->ln_bundle_cnt.p, 0, h->ln_bundle_cnt.bytes(), h->stream));
    h->ln_max_len = 0;
    for (int32_t a = 0; a < S.n_lines; ++a) h->ln_max_len = std::max(h->ln_max_len, S.line_ptr[a + 1] - S.line_ptr[a]);
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
    FEMB_CUDA(h, cudaMemsetAsync(h->ln_ae.p, 0, h->ln_ae.bytes(), h->stream));
    FEMB_CUDA(h, cudaMemsetAsync(h->ln_yle.p, 0, h->ln_yle.bytes(), h->stream));
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
  h->ln_mask_ok = false;
  h->ln_grp_count = 0;
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
  T.ae = h->ln_ae.p; T.yle = h->ln_yle.p; T.ent_of = h->ln_ent_of.p; T.rb = h->ln_rb.p; T.rbt = h->line_dist ? h->ln_rbt.p : h->ln_rb.p; T.yb = h->ln_yb.p; T.inv = h->ln_inv.p;
  T.grp_ptr = h->ln_grp_ptr.p; T.grp_lines = reinterpret_cast<const int4*>(h->ln_grp_lines.p); T.rank_mask = h->ln_rank_mask.p; T.line_sum = h->ln_line_sum.p; T.max_len = h->ln_max_len; T.bundle_cnt = h->ln_bundle_cnt.p;
  for (int f = 0; f < kLnMaxFam; ++f) { T.inv_off[f] = h->ln_inv_off[f]; T.fam_pad[f] = h->ln_fam_pad[f]; }
  for (int f = 0; f <= kLnMaxFam; ++f) {
    T.fam_off[f] = S.fam_off[f];
    T.range_off[f] = h->ln_range_off[f];
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
  if (dist && !h->ln_mask_ok) {
    // which ranks own a piece of which bundle (bit p = rank p): 2^rank per local bundle, summed over the ranks
    std::vector<double> md((size_t)S.n_coarse, 0.0);
    for (int32_t id : S.bundle_ids) md[id] = (double)(1 << h->dist_rank);
    DevBuf<double> tmp;
    FEMB_CUDA(h, upload(tmp, md, h->stream));
    rc = dist_allreduce(h, tmp.p, S.n_coarse);
    if (rc) return rc;
    FEMB_CUDA(h, download(md.data(), tmp.p, md.size() * sizeof(double), h->stream));
    FEMB_CUDA(h, cudaStreamSynchronize(h->stream));
    std::vector<int32_t> mi(md.size());
    for (size_t k = 0; k < md.size(); ++k) mi[k] = (int32_t)md[k];
    FEMB_CUDA(h, upload(h->ln_rank_mask, mi, h->stream));
    FEMB_CUDA(h, cudaStreamSynchronize(h->stream));
    h->ln_mask_ok = true;
  }
  if (dist) {
    const int64_t tot = (int64_t)kLnMaxFam * h->n_nodes;
    ln_node_w_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, h->stream>>>(T, h->free_mask.p);
    h->launches++;
  }
  rc = ebe_pair_aux(h);                              // per-pair constants of the operator phase (current coordinates)
  if (rc) return rc;
  if (S.n_lines > 0) {
    ln_direction_kernel<<<grid_lines, kLnThreads, 0, h->stream>>>(T, h->xyz.p, h->free_mask.p);
    ln_tridiag_kernel<<<grid_lines, kLnThreads, 0, h->stream>>>(T, h->Kvals.p);
    h->launches += 2;
  }
  FEMB_CUDA(h, cudaGetLastError());
  if (trace) cudaEventRecord(te[1], h->stream);
  bool all_ok = true;
  // One GPU: the three families' chains (Galerkin matrix -> augmented matrix -> 12 panel steps of three small kernels
  // -> extraction) are independent and latency bound (~0.6 ms each, launch after launch): they run on three streams
  // forked from the handle's and joined back into it; one host round trip reads the three pivot counters.
  // FEMB_LN_SETUP_STREAMS=0, FEMB_TRACE and the row-block partition (NCCL sums in between) keep the sequential path.
  static const bool concurrent_ok = !(getenv("FEMB_LN_SETUP_STREAMS") && atoi(getenv("FEMB_LN_SETUP_STREAMS")) == 0);
  if (!dist && !trace && concurrent_ok) {
    for (int f = 0; f < kLnMaxFam; ++f) {
      if (!h->ln_stream[f]) FEMB_CUDA(h, cudaStreamCreateWithFlags(&h->ln_stream[f], cudaStreamNonBlocking));
      if (!h->ln_ev_done[f]) FEMB_CUDA(h, cudaEventCreateWithFlags(&h->ln_ev_done[f], cudaEventDisableTiming));
    }
    if (!h->ln_ev_start) FEMB_CUDA(h, cudaEventCreateWithFlags(&h->ln_ev_start, cudaEventDisableTiming));
    FEMB_CUDA(h, h->ln_status.ensure(kLnMaxFam));
    FEMB_CUDA(h, cudaMemsetAsync(h->ln_status.p, 0, sizeof(int) * kLnMaxFam, h->stream));
    for (int f = 0; f < kLnMaxFam; ++f) {                      // allocations before the fork (cudaMalloc synchronises)
      const int64_t n_pad = h->ln_fam_pad[f], m = 2 * n_pad;
      if (S.fam_off[f + 1] == S.fam_off[f]) continue;
      FEMB_CUDA(h, h->ln_aug_f[f].ensure((size_t)m * m));
      FEMB_CUDA(h, h->ln_gal_f[f].ensure((size_t)n_pad * n_pad));
    }
    FEMB_CUDA(h, cudaEventRecord(h->ln_ev_start, h->stream));
    for (int f = 0; f < kLnMaxFam; ++f) {
      const int nf = S.fam_off[f + 1] - S.fam_off[f];
      if (nf == 0) continue;
      cudaStream_t st = h->ln_stream[f];
      const int64_t n_pad = h->ln_fam_pad[f], m = 2 * n_pad;
      FEMB_CUDA(h, cudaStreamWaitEvent(st, h->ln_ev_start, 0));
      FEMB_CUDA(h, cudaMemsetAsync(h->ln_aug_f[f].p, 0, (size_t)m * m * sizeof(double), st));
      FEMB_CUDA(h, cudaMemsetAsync(h->ln_gal_f[f].p, 0, (size_t)n_pad * n_pad * sizeof(double), st));
      const int n_rg = T.range_off[f + 1] - T.range_off[f];
      if (n_rg > 0) {
        ln_galerkin_kernel<<<n_rg, kLnThreads, (size_t)nf * sizeof(double), st>>>(T, f, h->rowptr.p, h->colidx.p, h->Kvals.p, h->ln_gal_f[f].p, n_pad);
        h->launches++;
      }
      ln_aug_fill_kernel<<<dim3((unsigned)((n_pad + 255) / 256), (unsigned)n_pad), 256, 0, st>>>(h->ln_gal_f[f].p, h->ln_aug_f[f].p, nf, n_pad);
      h->launches++;
      rc = coarse_invert_async(h, st, h->ln_aug_f[f].p, n_pad, h->ln_inv.p + h->ln_inv_off[f], h->ln_status.p + f);
      if (rc) return rc;
      FEMB_CUDA(h, cudaEventRecord(h->ln_ev_done[f], st));
      FEMB_CUDA(h, cudaStreamWaitEvent(h->stream, h->ln_ev_done[f], 0));
    }
    int* hs = reinterpret_cast<int*>(reinterpret_cast<char*>(h->pinned) + 1024);
    FEMB_CUDA(h, cudaMemcpyAsync(hs, h->ln_status.p, sizeof(int) * kLnMaxFam, cudaMemcpyDeviceToHost, h->stream));
    FEMB_CUDA(h, cudaStreamSynchronize(h->stream));
    for (int f = 0; f < kLnMaxFam; ++f) all_ok = all_ok && (hs[f] == 0);
    h->line_num_ok = all_ok;
    h->line_failed = !all_ok;
    return FEMB_OK;
  }
  std::vector<cudaEvent_t> tev;                     // FEMB_TRACE: Galerkin kernel | sum over ranks | inversion, per family
  auto mark = [&]() { if (trace) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, h->stream); tev.push_back(e); } };
  for (int f = 0; f < kLnMaxFam && (all_ok || dist); ++f) {
    const int nf = S.fam_off[f + 1] - S.fam_off[f];
    if (nf == 0) continue;
    const int64_t n_pad = h->ln_fam_pad[f], m = 2 * n_pad;
    FEMB_CUDA(h, h->coarse_aug.ensure((size_t)m * m));
    FEMB_CUDA(h, cudaMemsetAsync(h->coarse_aug.p, 0, (size_t)m * m * sizeof(double), h->stream));
    FEMB_CUDA(h, cudaMemsetAsync(h->ln_gal.p, 0, (size_t)n_pad * n_pad * sizeof(double), h->stream));
    const int n_rg = T.range_off[f + 1] - T.range_off[f];
    mark();
    if (n_rg > 0) {
      const size_t smem = (size_t)nf * sizeof(double);
      ln_galerkin_kernel<<<n_rg, kLnThreads, smem, h->stream>>>(T, f, h->rowptr.p, h->colidx.p, h->Kvals.p, h->ln_gal.p, n_pad);
      h->launches++;
    }
    mark();
    if (dist) {                                     // rows of the other ranks' bundles: sum over the ranks
      rc = dist_allreduce(h, h->ln_gal.p, (int)(n_pad * n_pad));
      if (rc) return rc;
    }
    mark();
    ln_aug_fill_kernel<<<dim3((unsigned)((n_pad + 255) / 256), (unsigned)n_pad), 256, 0, h->stream>>>(h->ln_gal.p, h->coarse_aug.p, nf, n_pad);
    h->launches++;
    FEMB_CUDA(h, cudaGetLastError());
    // row-block partition: every rank holds the same summed Galerkin matrix, so the inversions are dealt out (family
    // f to rank f mod world) instead of repeated on every rank; the inverses travel with one sum over the ranks below
    // (the others contribute zeros), a failed factorisation as a NaN in the first entry
    double* inv_f = h->ln_inv.p + h->ln_inv_off[f];
    if (dist && (f % h->dist_world) != h->dist_rank) {
      FEMB_CUDA(h, cudaMemsetAsync(inv_f, 0, (size_t)n_pad * n_pad * sizeof(double), h->stream));
      mark();
      continue;
    }
    bool ok = false;
    rc = coarse_invert(h, h->coarse_aug.p, n_pad, inv_f, &ok);
    if (rc) return rc;
    if (dist && !ok) {
      const double nan = std::numeric_limits<double>::quiet_NaN();
      FEMB_CUDA(h, cudaMemcpyAsync(inv_f, &nan, sizeof(double), cudaMemcpyHostToDevice, h->stream));
      FEMB_CUDA(h, cudaStreamSynchronize(h->stream));
    } else {
      all_ok = all_ok && ok;
    }
    mark();
  }
  if (dist) {
    int64_t tot = 0;
    for (int f = 0; f < kLnMaxFam; ++f) tot = std::max<int64_t>(tot, h->ln_inv_off[f] + (int64_t)h->ln_fam_pad[f] * h->ln_fam_pad[f]);
    rc = dist_allreduce(h, h->ln_inv.p, (int)tot);
    if (rc) return rc;
    double* first = reinterpret_cast<double*>(reinterpret_cast<char*>(h->pinned) + 1024);
    for (int f = 0; f < kLnMaxFam; ++f) {
      first[f] = 0.0;
      if (S.fam_off[f + 1] > S.fam_off[f])
        FEMB_CUDA(h, cudaMemcpyAsync(first + f, h->ln_inv.p + h->ln_inv_off[f], sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    }
    FEMB_CUDA(h, cudaStreamSynchronize(h->stream));
    for (int f = 0; f < kLnMaxFam; ++f)
      if (!(first[f] == first[f])) all_ok = false;
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
    double part[3] = {0.0, 0.0, 0.0};
    for (size_t k = 0; k + 3 < tev.size(); k += 4)
      for (int j = 0; j < 3; ++j) { float ms = 0.f; cudaEventElapsedTime(&ms, tev[k + j], tev[k + j + 1]); part[j] += ms; }
    fprintf(stderr, "[femb trace]   Galerkin kernels %.3f ms, their sums over the ranks %.3f ms, inversions on this rank %.3f ms\n", part[0], part[1], part[2]);
    for (auto e : tev) cudaEventDestroy(e);
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

// Lines -> line groups of the persistent kernel's grid (`n_groups` = CTAs x groups per CTA): longest processing time
// first onto the least loaded group.  Cost model: a fixed part per line (pointer loads, ticket, fence) + one unit per
// scan round of 16 entries.  Deterministic; any assignment gives the same numbers (the bundle sums are taken in line
// order by whichever group completes a bundle).
static int ln_assign_lines(femb_handle* h, int n_groups) {
  if (h->ln_grp_count == n_groups && h->ln_grp_ptr.p) return FEMB_OK;
  const LineSym& S = h->line_sym;
  std::vector<int32_t> order((size_t)S.n_lines);
  for (int32_t l = 0; l < S.n_lines; ++l) order[l] = l;
  auto cost = [&](int32_t l) { return 2 + (S.line_ptr[l + 1] - S.line_ptr[l] + 15) / 16; };
  std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return cost(a) > cost(b); });
  // least-loaded group: a binary heap of (load, group)
  std::vector<std::pair<int64_t, int32_t>> heap((size_t)n_groups);
  for (int g = 0; g < n_groups; ++g) heap[g] = {0, g};
  auto cmp = [](const std::pair<int64_t, int32_t>& a, const std::pair<int64_t, int32_t>& b) { return a > b; };   // min-heap
  std::make_heap(heap.begin(), heap.end(), cmp);
  std::vector<int32_t> of((size_t)S.n_lines), cnt((size_t)n_groups + 1, 0);
  for (int32_t l : order) {
    std::pop_heap(heap.begin(), heap.end(), cmp);
    auto& top = heap.back();
    of[l] = top.second;
    cnt[top.second + 1]++;
    top.first += cost(l);
    std::push_heap(heap.begin(), heap.end(), cmp);
  }
  for (int g = 0; g < n_groups; ++g) cnt[g + 1] += cnt[g];
  std::vector<int32_t> range_of((size_t)S.n_lines, 0);
  for (size_t rg = 0; rg + 1 < S.bundle_ptr.size(); ++rg)
    for (int32_t l = S.bundle_ptr[rg]; l < S.bundle_ptr[rg + 1]; ++l) range_of[l] = (int32_t)rg;
  std::vector<int32_t> lines((size_t)std::max(1, S.n_lines) * 4, 0), fill(cnt.begin(), cnt.end() - 1);
  for (int32_t l = 0; l < S.n_lines; ++l) {                              // ascending line index inside a group
    int32_t* r = lines.data() + 4 * (size_t)fill[of[l]]++;
    r[0] = l; r[1] = S.line_ptr[l]; r[2] = S.line_ptr[l + 1] - S.line_ptr[l]; r[3] = range_of[l];
  }
  FEMB_CUDA(h, upload(h->ln_grp_ptr, cnt, h->stream));
  FEMB_CUDA(h, upload(h->ln_grp_lines, lines, h->stream));
  FEMB_CUDA(h, cudaStreamSynchronize(h->stream));
  h->ln_grp_count = n_groups;
  return FEMB_OK;
}

// launches of the persistent kernel until the solve is over; `dist`: this rank's part of a row-block partition
static int run_mega(femb_handle* h, const femb_solve_opts& o, const double* d_b, femb_stats* st, bool dist) {
  const auto t_host0 = std::chrono::steady_clock::now();
  const int pstride = h->num_sms * 8;
  FEMB_CUDA(h, h->fpartials.ensure((size_t)pstride * 6));
  // u64 words: [2..10) phase clocks, [16..20) world scalars, [64..) the grid barrier's words (kBarWords x u32)
  FEMB_CUDA(h, h->mega_state.ensure(64 + kBarWords / 2));
  FEMB_CUDA(h, cudaMemsetAsync(h->mega_state.p, 0, h->mega_state.bytes(), h->stream));
  FEMB_CUDA(h, cudaMemsetAsync(h->flags.p, 0, sizeof(int32_t) * Flag::COUNT, h->stream));
  FEMB_CUDA(h, cudaMemsetAsync(h->scal.p, 0, sizeof(double) * Scal::COUNT, h->stream));
  // (5 or 6 CTAs of 96 / 80 registers per SM measured slower: 82.2 / 96.6 against 80.7 us per iteration,
  // profiles/r02_persistent_pcg_experiments.log)
  const void* fn = dist ? (const void*)ln_pcg_mega_kernel<true> : (const void*)ln_pcg_mega_kernel<false>;
  static int per_sm[2] = {0, 0};
  if (!per_sm[dist]) {
    FEMB_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[dist], fn, kMegaThreads, 0));
    per_sm[dist] = std::max(1, std::min(per_sm[dist], 8));      // the partial arrays hold num_sms * 8 entries
  }
  {
    const int rc = ln_assign_lines(h, h->num_sms * per_sm[dist] * (kMegaThreads / 16));
    if (rc) return rc;
  }
  MegaArgs A;
  A.T = ln_dev(h);
  A.P.xyz = h->xyz.p; A.P.conn = h->conn.p; A.P.elem_sec = h->elem_sec.p; A.P.sec_props = h->sec_props.p;
  A.P.E = h->E; A.P.G = h->G; A.P.rho = h->rho;
  A.pair_rec = reinterpret_cast<const int4*>(h->pair_rec.p);
  A.node_rec = reinterpret_cast<const int4*>(h->pair_node_rec.p);
  A.pair_aux = reinterpret_cast<const double4*>(h->pair_aux.p);
  A.free_mask = h->free_mask.p; A.dinv = h->Dinv.p; A.b = d_b;
  A.x = h->x.p; A.r = h->r.p; A.z = h->z.p; A.p = h->p.p; A.q = h->q.p; A.s = h->s.p;
  A.part = h->fpartials.p; A.scal = h->scal.p; A.flags = h->flags.p;
  A.bar = reinterpret_cast<unsigned int*>(h->mega_state.p + 64);
  A.phase_ns = h->mega_state.p + 2;
  A.glob = reinterpret_cast<double*>(h->mega_state.p + 16);
  A.cta_ns = nullptr;
  A.n_nodes = (int)(dist ? h->n_owned_nodes : h->n_nodes);
  A.n = (int64_t)A.n_nodes * 6;
  A.n_ranges = h->ln_n_ranges; A.pstride = pstride;
  A.max_iter = o.max_iter; A.rtol = o.rtol;
  A.p2p = dist ? reinterpret_cast<const P2PDev*>(h->p2p_dev_copy.p) : nullptr;
  const int grid = h->num_sms * per_sm[dist];
  DevBuf<unsigned long long> cta_ns;
  if (getenv("FEMB_TRACE")) {
    FEMB_CUDA(h, cta_ns.alloc((size_t)grid * 8));
    FEMB_CUDA(h, cudaMemsetAsync(cta_ns.p, 0, cta_ns.bytes(), h->stream));
    A.cta_ns = cta_ns.p;
  }
  // Krylov vectors pinned in the L2 for the duration of the solve: persisting lines inside the vector pool, streaming
  // (evict-first) for what does not fit the set-aside.  FEMB_L2_PERSIST=0 switches it off.
  bool l2_window = false;
  {
    static int want = -1;
    static size_t max_persist = 0, max_window = 0;
    size_t set_aside = 0;
    if (want < 0) {
      want = 1;
      if (const char* e = getenv("FEMB_L2_PERSIST")) want = atoi(e) != 0;
      cudaDeviceProp prop;
      if (want && cudaGetDeviceProperties(&prop, h->device) == cudaSuccess && prop.persistingL2CacheMaxSize > 0) {
        max_persist = (size_t)prop.persistingL2CacheMaxSize;
        max_window = (size_t)prop.accessPolicyMaxWindowSize;
      } else want = 0;
    }
    // Only when the whole pool fits the device's persisting maximum: the set-aside is taken from everybody else, and
    // with a pool several times the L2 (8M DOF: 383 MB) a partial window made the iteration TWICE as slow (measured:
    // 345 instead of 173 ms per step).  Set-aside = the pool: at 1M DOF 48 MB pinned 73.9 us per iteration, the device
    // maximum 74.6, no window 76.9.
    const bool fits = want && h->vec_pool.p && h->vec_pool.bytes() <= max_persist && h->vec_pool.bytes() <= max_window;
    const size_t need = fits ? h->vec_pool.bytes() : 0;
    if (need > 0) {
      if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, need) == cudaSuccess) set_aside = need;
      else cudaGetLastError();
    }
    if (fits && set_aside > 0) {
      cudaStreamAttrValue av;
      std::memset(&av, 0, sizeof(av));
      const size_t bytes = std::min(h->vec_pool.bytes(), max_window);
      av.accessPolicyWindow.base_ptr = h->vec_pool.p;
      av.accessPolicyWindow.num_bytes = bytes;
      av.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)set_aside / (double)bytes);
      av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
      if (cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &av) == cudaSuccess) l2_window = true;
      else cudaGetLastError();
    }
  }
  struct Peek { int32_t flags[Flag::COUNT]; double scal[Scal::COUNT]; unsigned long long ns[8]; };
  Peek* peek = reinterpret_cast<Peek*>(h->pinned);
  const int check = o.check_every > 0 ? o.check_every : 50;
  int it = 0, done = 0, launches = 0;
  while (!done && it <= o.max_iter) {
    // one more pass than updates: the pass that finds ||r|| <= tol (or the cap) leaves through the decision
    A.it0 = it; A.n_iters = std::min(check, o.max_iter + 1 - it); A.init = (it == 0) ? 1 : 0;
    void* args[] = {(void*)&A};
    FEMB_CUDA(h, cudaLaunchCooperativeKernel(fn, dim3((unsigned)grid), dim3(kMegaThreads), args, 0, h->stream));
    h->launches++;
    ++launches;
    FEMB_CUDA(h, cudaMemcpyAsync(peek->flags, h->flags.p, sizeof(peek->flags), cudaMemcpyDeviceToHost, h->stream));
    FEMB_CUDA(h, cudaMemcpyAsync(peek->scal, h->scal.p, sizeof(peek->scal), cudaMemcpyDeviceToHost, h->stream));
    FEMB_CUDA(h, cudaStreamSynchronize(h->stream));
    done = peek->flags[Flag::DONE];
    it = peek->flags[Flag::ITERS];
    if (!done && A.n_iters <= 0) break;
  }
  FEMB_CUDA(h, cudaMemcpyAsync(peek->ns, h->mega_state.p + 2, sizeof(peek->ns), cudaMemcpyDeviceToHost, h->stream));
  FEMB_CUDA(h, cudaStreamSynchronize(h->stream));
  if (l2_window) {
    cudaStreamAttrValue av;
    std::memset(&av, 0, sizeof(av));
    cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &av);     // num_bytes = 0: no window
    cudaCtxResetPersistingL2Cache();
    // give the set-aside back: left in place it slowed everything that streams through the L2 afterwards (measured:
    // fused assembly 0.102 -> 0.179 ms, assembled SpMV in-loop 0.68 -> 0.52 of the HBM peak)
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
  }
  if (st) {
    st->method_used = FEMB_SOLVER_PCG;
    st->op_used = FEMB_OP_EBE;
    st->coarse_dim = A.T.n_coarse;
    st->precond_used = FEMB_PRECOND_LINES;
    st->iterations = peek->flags[Flag::ITERS];
    st->converged = (done == 1);
    st->spmv_launches = peek->flags[Flag::ITERS];      // operator passes (phases of the persistent kernel)
    const double bb = peek->scal[Scal::BB];
    st->rel_residual = bb > 0.0 ? sqrt(peek->scal[Scal::RR] / bb) : 0.0;
    // phase clocks of CTA 0 (globaltimer): operator | update + line solves + coarse products + prolongation
    st->spmv_ms = (double)peek->ns[0] * 1e-6;
    st->update_ms = (double)(peek->ns[1] + peek->ns[2] + peek->ns[3] + peek->ns[4]) * 1e-6;
    st->spmv_timed = peek->flags[Flag::ITERS];
    if (getenv("FEMB_TRACE"))
      fprintf(stderr, "[femb trace] persistent PCG: %d iterations in %d launches; per-iteration phase times (CTA 0): operator %.2f us, "
              "update %.2f, line solves %.2f, coarse %.2f, prolongation %.2f; host wall of the call %.3f ms\n", st->iterations, launches,
              peek->ns[0] * 1e-3 / std::max(1, st->iterations), peek->ns[1] * 1e-3 / std::max(1, st->iterations),
              peek->ns[2] * 1e-3 / std::max(1, st->iterations), peek->ns[3] * 1e-3 / std::max(1, st->iterations),
              peek->ns[4] * 1e-3 / std::max(1, st->iterations),
              std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_host0).count());
    if (dist && A.cta_ns)
      fprintf(stderr, "[femb trace]   one thread's polls per iteration: bundle residuals of the world %.2f us, halo unpack %.2f us\n",
              peek->ns[5] * 1e-3 / std::max(1, st->iterations), peek->ns[6] * 1e-3 / std::max(1, st->iterations));
    if (A.cta_ns && st->iterations > 0) {
      std::vector<unsigned long long> w((size_t)grid * 8);
      cudaMemcpy(w.data(), cta_ns.p, w.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
      const char* nm[5] = {"operator", "update", "line solves", "coarse", "prolongation"};
      for (int ph = 0; ph < 5; ++ph) {
        double sum = 0.0, mx = 0.0, mn = 1e300, mx_hi = 0.0;
        for (int c = 0; c < grid; ++c) {
          const double v = (double)w[(size_t)c * 8 + ph];
          sum += v; mx = std::max(mx, v); mn = std::min(mn, v);
          if (c >= 64) mx_hi = std::max(mx_hi, v);      // the first CTAs also poll the other ranks' slots (partition)
        }
        fprintf(stderr, "[femb trace]   %-12s working time per CTA and iteration (barrier waits excluded): min %.2f us, mean %.2f, max %.2f (CTAs >= 64: %.2f)\n",
                nm[ph], mn * 1e-3 / st->iterations, sum / grid * 1e-3 / st->iterations, mx * 1e-3 / st->iterations, mx_hi * 1e-3 / st->iterations);
      }
    }
  }
  if (done == 4) return fail(h, FEMB_ERR_CUDA, "peer-memory exchange timed out waiting for another rank");
  if (done == 2) return fail(h, FEMB_ERR_SINGULAR, "PCG breakdown: p^T K p <= 0 (K_ff is not positive definite — unconstrained rigid-body motion or zero section properties?)");
  if (done != 1) return fail(h, FEMB_ERR_NOT_CONVERGED, "PCG did not reach rtol within max_iter");
  return FEMB_OK;
}

// compulsory bytes of one iteration of the persistent kernel (every array once per phase that must touch it):
//   operator   pair + node records, coordinates, z read, s written, mask (ebe_bytes) + 32 B per pair of constants
//   update     z, s, p, q, x, r read; p, q, x, r written (10 x 8 B / DOF) + per family and node: entry id 4,
//              direction 24 read, axial residual 8 written
//   lines      per line entry: axial residual 8 + factors 24 read, amplitude 8 written
//   coarse     the per-family inverses                                              8 B x sum nf^2
//   prolong    r, 1/diag read, z written (48 B each) + per family bundle id 4, entry id 4, amplitude 8, direction 24
double lines_iteration_bytes(const femb_handle* h) {
  if (!h->line_sym_ok || h->line_sym.n_coarse == 0) return 0.0;
  const LineSym& S = h->line_sym;
  const double nodes = (double)(h->line_dist ? h->n_owned_nodes : h->n_nodes);
  double inv = 0.0;
  for (int f = 0; f < kLnMaxFam; ++f) { const double nf = S.fam_off[f + 1] - S.fam_off[f]; inv += 8.0 * nf * nf; }
  return ebe_bytes(h, 1) + 32.0 * (double)h->sym.pair_code.size() + (480.0 + 36.0 * kLnMaxFam) * nodes + 40.0 * (double)S.n_entries + inv + (144.0 + 40.0 * kLnMaxFam) * nodes;
}

int pcg_lines(femb_handle* h, const femb_solve_opts& o, const double* d_b, femb_stats* st) {
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
  return run_mega(h, o, d_b, st, false);
}

// ---- row-block partition (dist.cu) -------------------------------------------------------------------------
// usable when the rank's line tables are set, the exchange runs through peer memory and the matrix-free operator applies
bool dist_lines_applicable(femb_handle* h, const femb_solve_opts& o, bool fused_p2p) {
  if (!(o.precond == FEMB_PRECOND_LINES || o.precond == FEMB_PRECOND_AUTO)) return false;
  return fused_p2p && h->bs == 6 && h->line_sym_ok && h->line_dist && !h->line_failed && ebe_available_dist(h);
}

int dist_lines_setup(femb_handle* h) {
  int rc = ensure_line_numeric(h);
  if (rc) return rc;
  return h->line_num_ok ? FEMB_OK : 1;          // 1: fall back to Jacobi
}

// the whole distributed solve of one right-hand side (owned rows of d_b); the caller has zeroed the ghost tails of the
// Krylov vectors, uploaded the sequence base and lined the ranks up
int dist_lines_solve(femb_handle* h, const femb_solve_opts& o, const double* d_b, femb_stats* st) {
  if (h->ln_n_ranges <= 0) return fail(h, FEMB_ERR_ARG, "line preconditioner: this rank owns no line piece");
  return run_mega(h, o, d_b, st, true);
}

}  // namespace femb
