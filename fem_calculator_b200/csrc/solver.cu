// BSR SpMV, boundary-condition masking, block-Jacobi PCG and reaction recovery.
//
// Operator.  BCs are eliminated by masking (free_mask): A = P K P + (I - P), P = diag(free).
// Every Krylov vector keeps exact zeros on the fixed DOFs, so K's fixed COLUMNS multiply
// zeros and only the fixed ROWS need overriding in the SpMV epilogue — K_ff is never
// materialised and the un-eliminated K stays available for r = K u - f
// (BeamSolver.py:412-418, ReactionSolver.py:199-205).
//
// SpMV.  BSR with bs x bs row-major blocks, one thread per scalar row: the bs threads of a
// block row read one contiguous 8*bs*bs-byte block per step with 16-byte loads, x comes
// through the read-only path (L2 resident: 8 B/DOF).  HBM-bound: 8 B/nnz + 4 B/block.
//
// Reductions are two-stage and ordered (per-CTA partial -> last CTA sums the partials in
// index order), so dot products — and therefore the whole PCG trajectory — are
// bit-reproducible run to run; float atomics are never used.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "common.cuh"
#include "pcg_common.cuh"

namespace femb {

// ------------------------------------------------------------------------------- SpMV
// y = A x over block rows; MASKED: rows with free_mask==0 return x (identity rows).
// DOT: also accumulates sum(x_i * y_i) into `dot` through the ordered grid reduction and,
// being the PCG's first kernel of an iteration, honours the done flag.
template <int BS, bool MASKED, bool DOT, int THREADS, int UNR = 2>
__global__ void __launch_bounds__(THREADS)
bsr_spmv_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                const double* __restrict__ vals, const uint8_t* __restrict__ free_mask,
                const double* __restrict__ x, double* __restrict__ y, int64_t nrows,
                double* partials, int pstride, double* scal, int* flags,
                const uint8_t* __restrict__ skip_node, const int32_t* __restrict__ node_list,
                const P2PDev* __restrict__ p2p) {
  // skip_node != nullptr: block rows flagged there are left to a later launch (rows that read
  // ghost columns wait for the halo); node_list != nullptr: row g belongs to node_list[g / BS].
  static_assert(BS == 6 || BS == 3, "block size");
  if (DOT && flags[Flag::DONE]) return;
  if (DOT && p2p) {
    // fused peer-memory mode: the neighbours' update kernels stored the ghost entries of x straight
    // into this rank's vector and released a sequence flag — wait for it before gathering
    if ((int)threadIdx.x < p2p->n_nbr) {
      const long long seq = p2p->base[0] + flags[Flag::ITERS] + 1;
      long long spins = 0;
      while (ld_acquire_sys(p2p->my_halo_flag + p2p->nbr[threadIdx.x]) < seq) {
        if (++spins > kSpinLimit) { flags[Flag::DONE] = 4; break; }
      }
    }
    __syncthreads();
  }
  double dot = 0.0;
  const int n_owned_p2p = (DOT && p2p) ? (int)(nrows / BS) : 0;
  for (int64_t g = (int64_t)blockIdx.x * THREADS + threadIdx.x; g < nrows; g += (int64_t)gridDim.x * THREADS) {
    int node = (int)(g / BS);
    const int r = (int)(g - (int64_t)node * BS);
    if (node_list) node = __ldg(node_list + node);
    if (skip_node && skip_node[node]) continue;
    const int64_t go = (int64_t)node * BS + r;   // == g unless list-driven
    const int b0 = __ldg(rowptr + node), b1 = __ldg(rowptr + node + 1);
    double acc = 0.0;
    if (BS == 6) {
      // three 16-byte loads per 48-byte row, two blocks in flight per thread.  Measured inside
      // PCG at 1M DOF on one B200 (profiles/r01_spmv_variants.log): unroll 2 -> 73 us, unroll 4 -> 85,
      // unroll 8 -> 78, unroll 1 -> 83; a sector-exact 32 B + 16 B split per row 10 % slower; two rows
      // per thread with 256-bit loads 78-85 us; TMA-staged tiles (spmv_tma.cu) 99-133 us.  Reading
      // only the diagonal + upper blocks from HBM and the lower ones as L2-resident transposes cut
      // DRAM traffic to 237 MB (ncu) but not the time (81.5 vs 82.3 us): the kernel is bound by the
      // dependent colidx -> x gather steps per row, not by bandwidth.
#pragma unroll UNR
      for (int b = b0; b < b1; ++b) {
        const int col = __ldg(colidx + b);
        const double2* a2 = reinterpret_cast<const double2*>(vals + (size_t)b * 36 + r * 6);
        const double2* x2 = reinterpret_cast<const double2*>(x + (size_t)col * 6);
        const double2 a0 = __ldcs(a2), a1 = __ldcs(a2 + 1), a2v = __ldcs(a2 + 2);
        double2 x0, x1, x2v;
        if (p2p && col >= n_owned_p2p) {
          // ghost entries are stored by peer GPUs, possibly while CTAs of this kernel are already
          // resident: read them L2-coherently (the non-coherent path served stale L1 lines — measured
          // as 15,319 instead of 6,931 iterations at 1M DOF on 2 GPUs)
          x0 = __ldcg(x2); x1 = __ldcg(x2 + 1); x2v = __ldcg(x2 + 2);
        } else {
          x0 = __ldg(x2); x1 = __ldg(x2 + 1); x2v = __ldg(x2 + 2);
        }
        acc += a0.x * x0.x; acc += a0.y * x0.y; acc += a1.x * x1.x;
        acc += a1.y * x1.y; acc += a2v.x * x2v.x; acc += a2v.y * x2v.y;
      }
    } else {
#pragma unroll 4
      for (int b = b0; b < b1; ++b) {
        const int col = __ldg(colidx + b);
        const double* a = vals + (size_t)b * 9 + r * 3;
        const double* xc = x + (size_t)col * 3;
        acc += __ldcs(a) * __ldg(xc); acc += __ldcs(a + 1) * __ldg(xc + 1); acc += __ldcs(a + 2) * __ldg(xc + 2);
      }
    }
    const double xg = x[go];
    if (MASKED && !free_mask[go]) acc = xg;
    y[go] = acc;
    if (DOT) dot += xg * acc;
  }
  if (DOT) {
    double mine[1], tot[1];
    mine[0] = dot;
    if (grid_reduce<THREADS, 1>(mine, partials, pstride, flags + Flag::TICKET0, tot)) {
      if (threadIdx.x == 0) scal[Scal::PQ] = tot[0];
      if (p2p && (int)threadIdx.x < p2p->world) {
        // post {delta, gamma, ||r||^2} of this rank in every peer's mailbox (scal = the solver's red[])
        const long long seq = p2p->base[0] + flags[Flag::ITERS] + 1;
        MailSlot* dst = p2p->peer_mail[threadIdx.x] + (p2p->rank * 2 + (int)(seq & 1));
        dst->v[0] = tot[0]; dst->v[1] = scal[1]; dst->v[2] = scal[2]; dst->v[3] = 0.0;
        __threadfence_system();
        st_release_sys(&dst->seq, seq);
      }
    }
  }
}

// ------------------------------------------------------------------------------- BC
__global__ void fill_mask_kernel(uint8_t* mask, int64_t n, uint8_t v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) mask[i] = v;
}
__global__ void clear_fixed_kernel(uint8_t* mask, const int64_t* fixed, int64_t nf) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nf; i += (int64_t)gridDim.x * blockDim.x) mask[fixed[i]] = 0;
}
// u0m = prescribed values on fixed DOFs, 0 elsewhere
__global__ void mask_prescribed_kernel(const double* u0, const uint8_t* mask, double* out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = mask[i] ? 0.0 : u0[i];
}
// b = free ? f - (K u0)_i : 0     (f_f - k_fs u_s, BeamSolver.py:416)
__global__ void rhs_kernel(const double* f, const double* Ku0, const uint8_t* mask, double* b, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    b[i] = mask[i] ? (Ku0 ? f[i] - Ku0[i] : f[i]) : 0.0;
}

// ------------------------------------------------------------------- preconditioner
// Block-Jacobi: invert the masked diagonal block (fixed rows/cols -> identity) by
// Gauss-Jordan without pivoting (SPD); a non-positive pivot degrades that row to identity.
// mode 1 (Jacobi) keeps only the reciprocal diagonal inside the same block layout.
template <int BS>
__global__ void precond_setup_kernel(const double* __restrict__ vals, const int32_t* __restrict__ diag_blk,
                                     const uint8_t* __restrict__ mask, double* __restrict__ Dinv,
                                     int64_t n_nodes, int mode) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_nodes) return;
  double A[BS][BS], B[BS][BS];
  const double* d = vals + (size_t)diag_blk[i] * BS * BS;
#pragma unroll
  for (int r = 0; r < BS; ++r)
#pragma unroll
    for (int c = 0; c < BS; ++c) {
      const bool fr = mask[i * BS + r] && mask[i * BS + c];
      double v = fr ? d[r * BS + c] : (r == c ? 1.0 : 0.0);
      if (mode != 2 && r != c) v = 0.0;
      if (mode == 0) v = (r == c) ? 1.0 : 0.0;
      A[r][c] = v;
      B[r][c] = (r == c) ? 1.0 : 0.0;
    }
#pragma unroll
  for (int k = 0; k < BS; ++k) {
    double piv = A[k][k];
    if (!(piv > 0.0)) {  // not SPD here (e.g. an unconnected point): fall back to identity row
#pragma unroll
      for (int c = 0; c < BS; ++c) { A[k][c] = (c == k) ? 1.0 : 0.0; A[c][k] = (c == k) ? 1.0 : 0.0; B[k][c] = (c == k) ? 1.0 : 0.0; }
      piv = 1.0;
    }
    const double ip = 1.0 / piv;
#pragma unroll
    for (int c = 0; c < BS; ++c) { A[k][c] *= ip; B[k][c] *= ip; }
#pragma unroll
    for (int r = 0; r < BS; ++r) {
      if (r == k) continue;
      const double m = A[r][k];
#pragma unroll
      for (int c = 0; c < BS; ++c) { A[r][c] -= m * A[k][c]; B[r][c] -= m * B[k][c]; }
    }
  }
  double* o = Dinv + (size_t)i * BS * BS;
#pragma unroll
  for (int r = 0; r < BS; ++r)
#pragma unroll
    for (int c = 0; c < BS; ++c) o[r * BS + c] = 0.5 * (B[r][c] + B[c][r]);
}

// scalar Jacobi: reciprocal of the masked diagonal (identity on fixed DOFs / non-positive pivots)
template <int BS>
__global__ void jacobi_setup_kernel(const double* __restrict__ vals, const int32_t* __restrict__ diag_blk,
                                    const uint8_t* __restrict__ mask, double* __restrict__ dinv, int64_t n,
                                    int identity) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  const int64_t node = g / BS;
  const int r = (int)(g - node * BS);
  const double d = vals[(size_t)diag_blk[node] * BS * BS + r * BS + r];
  dinv[g] = (!identity && mask[g] && d > 0.0) ? 1.0 / d : 1.0;
}

// -------------------------------------------------------------------------------- PCG
// Chronopoulos-Gear form of preconditioned CG: two kernels per iteration,
//   (1) s = A z with delta = (z, s)                      [bsr_spmv_kernel<.., DOT>]
//   (2) beta = gamma/gamma_prev, alpha = gamma / (delta - beta*gamma/alpha_prev),
//       p = z + beta p, q = s + beta q (= A p), x += alpha p, r -= alpha q, z = Dinv r,
//       gamma_new = (r, z), rr = (r, r), convergence / breakdown decision  [pcg_update_kernel]
// Every vector is read and written once per iteration by one fused kernel; all dot products go
// through the ordered two-stage reduction (bit-reproducible).
// init: x = 0, r = b, z = Dinv r, p = q = 0; gamma -> RZ1 (the slot iteration 0 reads), bb
template <int BS, int THREADS, bool BLOCKJ>
__global__ void __launch_bounds__(THREADS)
pcg_init_kernel(const double* __restrict__ b, const double* __restrict__ Dinv, double* __restrict__ x,
                double* __restrict__ r, double* __restrict__ z, double* __restrict__ p, double* __restrict__ q,
                int64_t n, double rtol, double* partials, int pstride, double* scal, int* flags) {
  double rz = 0.0, bb = 0.0;
  for (int64_t g = (int64_t)blockIdx.x * THREADS + threadIdx.x; g < n; g += (int64_t)gridDim.x * THREADS) {
    const double bg = b[g];
    double zg;
    if (BLOCKJ) {
      const int64_t node = g / BS;
      double rn[BS];
#pragma unroll
      for (int c = 0; c < BS; ++c) rn[c] = b[node * BS + c];
      zg = apply_dinv_row<BS>(Dinv, g, rn);
    } else {
      zg = Dinv[g] * bg;
    }
    x[g] = 0.0; r[g] = bg; z[g] = zg; p[g] = 0.0; q[g] = 0.0;
    rz += bg * zg; bb += bg * bg;
  }
  double mine[2], tot[2];
  mine[0] = rz;
  mine[1] = bb;
  if (grid_reduce<THREADS, 2>(mine, partials, pstride, flags + Flag::TICKET1, tot)) {
    if (threadIdx.x == 0) {
      scal[Scal::RZ0] = tot[0]; scal[Scal::RZ1] = tot[0];
      scal[Scal::BB] = tot[1]; scal[Scal::RR] = tot[1];
      scal[Scal::TOL2] = rtol * rtol * tot[1];
      scal[Scal::ALPHA] = 1.0;
      flags[Flag::ITERS] = 0;
      flags[Flag::DONE] = (tot[1] == 0.0) ? 1 : 0;  // zero load: u = 0 is the answer
    }
  }
}

// iteration `it` (parity = it & 1): gamma_it lives in RZ0 + (parity ^ 1), gamma_{it+1} goes to RZ0 + parity
template <int BS, int THREADS, bool BLOCKJ>
__global__ void __launch_bounds__(THREADS)
pcg_update_kernel(const double* __restrict__ Dinv, const double* __restrict__ s, double* __restrict__ p,
                  double* __restrict__ q, double* __restrict__ x, double* __restrict__ r, double* __restrict__ z,
                  int64_t n, int parity, int first, int max_iter, double* partials, int pstride, double* scal,
                  int* flags) {
  if (flags[Flag::DONE]) return;
  const double delta = scal[Scal::PQ];
  const double gamma = scal[Scal::RZ0 + (parity ^ 1)];
  const double beta = first ? 0.0 : gamma / scal[Scal::RZ0 + parity];
  const double den = first ? delta : delta - beta * gamma / scal[Scal::ALPHA];
  const bool bad = !(den > 0.0);           // K_ff not positive definite along p
  const double alpha = bad ? 0.0 : gamma / den;
  double rz = 0.0, rr = 0.0;
  if (BLOCKJ) {
    for (int64_t g = (int64_t)blockIdx.x * THREADS + threadIdx.x; g < n; g += (int64_t)gridDim.x * THREADS) {
      const int64_t nb = (g / BS) * BS;
      double rn[BS];
#pragma unroll
      for (int c = 0; c < BS; ++c) rn[c] = r[nb + c] - alpha * (s[nb + c] + beta * q[nb + c]);
      const int rloc = (int)(g - nb);
      double rg = rn[0];
#pragma unroll
      for (int c = 1; c < BS; ++c) rg = (rloc == c) ? rn[c] : rg;
      const double zg = apply_dinv_row<BS>(Dinv, g, rn);
      const double pg = z[g] + beta * p[g];
      p[g] = pg;
      x[g] += alpha * pg;
      z[g] = zg;
      rz += rg * zg; rr += rg * rg;
    }
    // q[g] and r[g] are also read by the other rows of their node, so they are committed in a
    // second pass.  All rows of a node live in the same CTA and grid-stride step
    // (THREADS % BS == 0), hence the barrier is enough; the re-reads hit L1.
    __syncthreads();
    for (int64_t g = (int64_t)blockIdx.x * THREADS + threadIdx.x; g < n; g += (int64_t)gridDim.x * THREADS) {
      const double qg = s[g] + beta * q[g];
      q[g] = qg;
      r[g] -= alpha * qg;
    }
  } else {
    // scalar Jacobi: purely element-wise, one pass, two rows per thread and step (16-byte accesses)
    const int64_t n2 = n >> 1;
    const double2* z2 = reinterpret_cast<const double2*>(z);
    const double2* s2 = reinterpret_cast<const double2*>(s);
    const double2* d2 = reinterpret_cast<const double2*>(Dinv);
    double2* p2 = reinterpret_cast<double2*>(p);
    double2* q2 = reinterpret_cast<double2*>(q);
    double2* x2 = reinterpret_cast<double2*>(x);
    double2* r2 = reinterpret_cast<double2*>(r);
    double2* zo = reinterpret_cast<double2*>(z);
#pragma unroll 2
    for (int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x; i < n2; i += (int64_t)gridDim.x * THREADS) {
      const double2 zv = z2[i], sv = s2[i], dv = __ldg(d2 + i);
      double2 pv = p2[i], qv = q2[i], xv = x2[i], rv = r2[i];
      pv.x = zv.x + beta * pv.x; pv.y = zv.y + beta * pv.y;
      qv.x = sv.x + beta * qv.x; qv.y = sv.y + beta * qv.y;
      xv.x += alpha * pv.x; xv.y += alpha * pv.y;
      rv.x -= alpha * qv.x; rv.y -= alpha * qv.y;
      const double2 zn = make_double2(dv.x * rv.x, dv.y * rv.y);
      p2[i] = pv; q2[i] = qv; x2[i] = xv; r2[i] = rv; zo[i] = zn;
      rz += rv.x * zn.x; rz += rv.y * zn.y;
      rr += rv.x * rv.x; rr += rv.y * rv.y;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {   // odd tail (3-DOF meshes with an odd node count)
      const int64_t g = n - 1;
      const double pg = z[g] + beta * p[g];
      const double qg = s[g] + beta * q[g];
      const double rg = r[g] - alpha * qg;
      const double zg = Dinv[g] * rg;
      p[g] = pg; q[g] = qg; x[g] += alpha * pg; r[g] = rg; z[g] = zg;
      rz += rg * zg; rr += rg * rg;
    }
  }
  double mine[2], tot[2];
  mine[0] = rz;
  mine[1] = rr;
  if (grid_reduce<THREADS, 2>(mine, partials, pstride, flags + Flag::TICKET1, tot)) {
    if (threadIdx.x == 0) {
      scal[Scal::RZ0 + parity] = tot[0];
      scal[Scal::RR] = tot[1];
      scal[Scal::ALPHA] = alpha;
      const int it = flags[Flag::ITERS] + 1;
      flags[Flag::ITERS] = it;
      if (bad) flags[Flag::DONE] = 2;
      else if (tot[1] <= scal[Scal::TOL2]) flags[Flag::DONE] = 1;
      else if (it >= max_iter) flags[Flag::DONE] = 3;
    }
  }
}

// ---- linked single-vector PCG (matrix-free operator, scalar Jacobi): see PcgLink -------------
// init: x = 0, r = b, z = Dinv r, p = q = 0; publishes {gamma_0, ||b||^2} into buffer 0
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
pcg_init_linked_kernel(const double* __restrict__ b, const double* __restrict__ dinv, double* __restrict__ x,
                       double* __restrict__ r, double* __restrict__ z, double* __restrict__ p, double* __restrict__ q,
                       int64_t n, const PcgLink L) {
  __shared__ double s_part[2 * THREADS / 32];
  double v[2] = {0.0, 0.0};
  for (int64_t g = (int64_t)blockIdx.x * THREADS + threadIdx.x; g < n; g += (int64_t)gridDim.x * THREADS) {
    const double bg = b[g];
    const double zg = dinv[g] * bg;
    x[g] = 0.0; r[g] = bg; z[g] = zg; p[g] = 0.0; q[g] = 0.0;
    v[0] += bg * zg; v[1] += bg * bg;
  }
  block_sum_all<THREADS, 2>(v, s_part);
  if (threadIdx.x == 0) { L.upd_partials[blockIdx.x] = v[0]; L.upd_partials[L.pstride + blockIdx.x] = v[1]; }
  if (blockIdx.x == 0 && threadIdx.x == 0) { L.flags[Flag::DONE] = 0; L.flags[Flag::ITERS] = 0; }
}

// update(it): consumes delta (operator) and {gamma, rr} (previous update / init), then the same
// element-wise pass as pcg_update_kernel's scalar-Jacobi branch; publishes the new {gamma, rr}
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
pcg_update_linked_kernel(const double* __restrict__ dinv, const double* __restrict__ s, double* __restrict__ p,
                         double* __restrict__ q, double* __restrict__ x, double* __restrict__ r, double* __restrict__ z,
                         int64_t n, const PcgLink L) {
  __shared__ double s_part[2 * THREADS / 32];
  if (L.flags[Flag::DONE]) return;
  const int rd = L.it & 1, wr = rd ^ 1;
  double tot[2] = {0.0, 0.0};
  {
    const double* pu = L.upd_partials + (size_t)rd * 2 * L.pstride;
    for (int i = threadIdx.x; i < L.n_op; i += THREADS) tot[0] += __ldcg(L.op_partials + i);
    for (int i = threadIdx.x; i < L.n_upd; i += THREADS) tot[1] += __ldcg(pu + i);
  }
  block_sum_all<THREADS, 2>(tot, s_part);
  const double delta = tot[0], gamma = tot[1];
  const bool first = (L.it == 0);
  const double beta = first ? 0.0 : gamma / L.scal[Scal::RZ0 + rd];
  const double den = first ? delta : delta - beta * gamma / L.scal[Scal::ALPHA + rd];
  const bool bad = !(den > 0.0);           // K_ff not positive definite along p
  const double alpha = bad ? 0.0 : gamma / den;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    L.scal[Scal::RZ0 + wr] = gamma;
    L.scal[Scal::ALPHA + wr] = alpha;
    L.scal[Scal::PQ] = delta;
    if (bad) L.flags[Flag::DONE] = 2;
  }
  if (bad) return;
  double v[2] = {0.0, 0.0};
  const int64_t n2 = n >> 1;
  const double2* z2 = reinterpret_cast<const double2*>(z);
  const double2* s2 = reinterpret_cast<const double2*>(s);
  const double2* d2 = reinterpret_cast<const double2*>(dinv);
  double2* p2 = reinterpret_cast<double2*>(p);
  double2* q2 = reinterpret_cast<double2*>(q);
  double2* x2 = reinterpret_cast<double2*>(x);
  double2* r2 = reinterpret_cast<double2*>(r);
  double2* zo = reinterpret_cast<double2*>(z);
#pragma unroll 2
  for (int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x; i < n2; i += (int64_t)gridDim.x * THREADS) {
    const double2 zv = z2[i], sv = s2[i], dv = __ldg(d2 + i);
    double2 pv = p2[i], qv = q2[i], xv = x2[i], rv = r2[i];
    pv.x = zv.x + beta * pv.x; pv.y = zv.y + beta * pv.y;
    qv.x = sv.x + beta * qv.x; qv.y = sv.y + beta * qv.y;
    xv.x += alpha * pv.x; xv.y += alpha * pv.y;
    rv.x -= alpha * qv.x; rv.y -= alpha * qv.y;
    const double2 zn = make_double2(dv.x * rv.x, dv.y * rv.y);
    p2[i] = pv; q2[i] = qv; x2[i] = xv; r2[i] = rv; zo[i] = zn;
    v[0] += rv.x * zn.x; v[0] += rv.y * zn.y;
    v[1] += rv.x * rv.x; v[1] += rv.y * rv.y;
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {   // odd tail
    const int64_t g = n - 1;
    const double pg = z[g] + beta * p[g];
    const double qg = s[g] + beta * q[g];
    const double rg = r[g] - alpha * qg;
    const double zg = dinv[g] * rg;
    p[g] = pg; q[g] = qg; x[g] += alpha * pg; r[g] = rg; z[g] = zg;
    v[0] += rg * zg; v[1] += rg * rg;
  }
  block_sum_all<THREADS, 2>(v, s_part);
  if (threadIdx.x == 0) {
    double* po = L.upd_partials + (size_t)wr * 2 * L.pstride;
    po[blockIdx.x] = v[0]; po[L.pstride + blockIdx.x] = v[1];
  }
}

// one CTA: the operator-side decision alone, so the host's poll sees the state after the last update
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
pcg_decide_linked_kernel(const PcgLink L) {
  __shared__ double s_part[2 * THREADS / 32];
  if (L.flags[Flag::DONE]) return;
  pcg_link_decide<THREADS>(L, s_part);
}

// out = K u - f (minus_f) or K u
__global__ void axpy_sub_kernel(double* out, const double* f, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] -= f[i];
}
// x[fixed] = prescribed
__global__ void set_prescribed_kernel(double* x, const double* u0, const uint8_t* mask, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (!mask[i]) x[i] = u0[i];
}

// ---------------------------------------------------------------------------- host side
int launch_spmv(femb_handle* h, const double* x, double* y, bool masked, double* dot_partials) {
  return launch_spmv_rows(h, x, y, h->ndof, masked, dot_partials, h->scal.p, nullptr, nullptr, nullptr);
}

// rows [0, n) only (the distributed solver owns a prefix of the local rows); the fused dot goes
// to scal_out[Scal::PQ]
int launch_spmv_rows(femb_handle* h, const double* x, double* y, int64_t n, bool masked, double* dot_partials,
                     double* scal_out, const uint8_t* skip_node, const int32_t* node_list, const void* p2p_dev) {
  // row-block distributed PCG on a frame: the matrix-free operator over the owned nodes (dist.cu passes
  // scal_out = red + DELTA, whose neighbours red[1], red[2] the peer-memory epilogue posts, and the
  // reduction's flags block)
  if (masked && dot_partials && !skip_node && !node_list && h->bs == 6 && n == h->n_owned_nodes * 6 && ebe_available_dist(h))
    return launch_ebe(h, x, y, 1, true, dot_partials, scal_out, h->flags.p + Flag::TICKET0, h->flags.p + Flag::DONE, nullptr,
                      p2p_dev, h->n_owned_nodes);
  const int pstride = h->num_sms * 8;
  const int grid = vec_grid(h, n, kRowThreads);
#define SPMV(BS, M, D)                                                                         \
  bsr_spmv_kernel<BS, M, D, kRowThreads><<<grid, kRowThreads, 0, h->stream>>>(                  \
      h->rowptr.p, h->colidx.p, h->Kvals.p, h->free_mask.p, x, y, n, dot_partials, pstride,    \
      scal_out, h->flags.p, skip_node, node_list, reinterpret_cast<const P2PDev*>(p2p_dev))
  const bool dot = dot_partials != nullptr;
  if (h->bs == 6) {
    if (masked && dot) SPMV(6, true, true);
    else if (masked) SPMV(6, true, false);
    else SPMV(6, false, false);
  } else {
    if (masked && dot) SPMV(3, true, true);
    else if (masked) SPMV(3, true, false);
    else SPMV(3, false, false);
  }
#undef SPMV
  h->launches++;
  FEMB_CUDA(h, cudaGetLastError());
  return FEMB_OK;
}

int setup_bc_vectors(femb_handle* h) {
  const int64_t n = h->ndof;
  FEMB_CUDA(h, h->b.alloc(n));
  // the six Krylov vectors share ONE allocation: the persistent PCG kernel pins that range in the L2 (access policy
  // window, lines.cu) so that the vectors survive the table streams of the other phases
  const size_t stride = ((size_t)n + 31) / 32 * 32;
  if (h->vec_pool.n != 6 * stride || !h->vec_pool.p) {
    for (DevBuf<double>* v : {&h->x, &h->r, &h->z, &h->p, &h->q, &h->s})
      if (!(v == &h->z && h->p2p_z_exported && h->z.p == h->p2p_z_exported && h->z.n == (size_t)n)) v->release();
    FEMB_CUDA(h, h->vec_pool.alloc(6 * stride));
  }
  {
    int k = 0;
    for (DevBuf<double>* v : {&h->x, &h->r, &h->z, &h->p, &h->q, &h->s}) {
      double* slot = h->vec_pool.p + stride * (size_t)(k++);
      // (a row-block rank's z lives in its exported peer-memory allocation, dist.cu: it stays there)
      if (v == &h->z && h->p2p_z_exported && h->z.p == h->p2p_z_exported && h->z.n == (size_t)n) continue;
      v->adopt(slot, (size_t)n);
    }
  }
  FEMB_CUDA(h, h->partials.alloc((size_t)h->num_sms * 8 * 4));
  FEMB_CUDA(h, h->scal.alloc(Scal::COUNT));
  FEMB_CUDA(h, h->flags.alloc(Flag::COUNT));
  FEMB_CUDA(h, cudaMemsetAsync(h->flags.p, 0, sizeof(int32_t) * Flag::COUNT, h->stream));
  FEMB_CUDA(h, cudaMemsetAsync(h->scal.p, 0, sizeof(double) * Scal::COUNT, h->stream));
  return FEMB_OK;
}

// b = masked(f - K u0)   (u0 == nullptr -> b = masked f)
static int build_rhs(femb_handle* h, bool have_u0) {
  const int64_t n = h->ndof;
  const int g = vec_grid(h, n, kVecThreads);
  if (have_u0) {
    mask_prescribed_kernel<<<g, kVecThreads, 0, h->stream>>>(h->u0.p, h->free_mask.p, h->p.p, n);
    h->launches++;
    int rc = launch_spmv(h, h->p.p, h->q.p, false, nullptr);
    if (rc) return rc;
    rhs_kernel<<<g, kVecThreads, 0, h->stream>>>(h->f.p, h->q.p, h->free_mask.p, h->b.p, n);
  } else {
    rhs_kernel<<<g, kVecThreads, 0, h->stream>>>(h->f.p, nullptr, h->free_mask.p, h->b.p, n);
  }
  h->launches++;
  FEMB_CUDA(h, cudaGetLastError());
  return FEMB_OK;
}

static int setup_precond(femb_handle* h, int mode) {
  if (mode != FEMB_PRECOND_BLOCK_JACOBI) {
    FEMB_CUDA(h, h->Dinv.alloc((size_t)h->ndof));
    const int g = (int)((h->ndof + 255) / 256);
    const int ident = (mode == FEMB_PRECOND_NONE);
    if (h->bs == 6) jacobi_setup_kernel<6><<<g, 256, 0, h->stream>>>(h->Kvals.p, h->diag_blk.p, h->free_mask.p, h->Dinv.p, h->ndof, ident);
    else jacobi_setup_kernel<3><<<g, 256, 0, h->stream>>>(h->Kvals.p, h->diag_blk.p, h->free_mask.p, h->Dinv.p, h->ndof, ident);
    h->launches++;
    FEMB_CUDA(h, cudaGetLastError());
    return FEMB_OK;
  }
  FEMB_CUDA(h, h->Dinv.alloc((size_t)h->n_nodes * h->bs * h->bs));
  const int grid = (int)((h->n_nodes + 127) / 128);
  if (h->bs == 6)
    precond_setup_kernel<6><<<grid, 128, 0, h->stream>>>(h->Kvals.p, h->diag_blk.p, h->free_mask.p, h->Dinv.p, h->n_nodes, mode);
  else
    precond_setup_kernel<3><<<grid, 128, 0, h->stream>>>(h->Kvals.p, h->diag_blk.p, h->free_mask.p, h->Dinv.p, h->n_nodes, mode);
  h->launches++;
  FEMB_CUDA(h, cudaGetLastError());
  return FEMB_OK;
}

static int pcg_core(femb_handle* h, const femb_solve_opts& o, const double* d_b, femb_stats* st);

int run_pcg(femb_handle* h, const femb_solve_opts& o, femb_stats* st) {
  int rc = build_rhs(h, h->u0.p != nullptr);
  if (rc) return rc;
  return pcg_core(h, o, h->b.p, st);
}

// K_ff x = b for an arbitrary (already masked) device right-hand side; solution in h->x
int pcg_solve_rhs(femb_handle* h, const femb_solve_opts& o, const double* d_b, femb_stats* st) {
  return pcg_core(h, o, d_b, st);
}

struct PcgPeek {
  int32_t flags[Flag::COUNT];
  double scal[Scal::COUNT];
};

// Matrix-free operator + scalar Jacobi: two kernels per iteration with linked reductions (PcgLink).
static int pcg_core_linked(femb_handle* h, const femb_solve_opts& o, const double* d_b, femb_stats* st) {
  const int64_t n = h->ndof;
  const int pstride = h->num_sms * 8;
  int rc = setup_precond(h, o.precond);
  if (rc) return rc;
  FEMB_CUDA(h, h->fpartials.ensure((size_t)pstride * 6));
  FEMB_CUDA(h, cudaMemsetAsync(h->flags.p, 0, sizeof(int32_t) * Flag::COUNT, h->stream));
  FEMB_CUDA(h, cudaMemsetAsync(h->scal.p, 0, sizeof(double) * Scal::COUNT, h->stream));
  constexpr int UT = 256;
  const int grid_u = occ_grid(h, pcg_update_linked_kernel<UT>, (n + 1) / 2, UT);
  PcgLink L;
  L.upd_partials = h->fpartials.p; L.op_partials = h->fpartials.p + (size_t)4 * pstride;
  L.scal = h->scal.p; L.flags = h->flags.p;
  L.n_upd = grid_u; L.n_op = ebe_grid(h, 1, h->n_nodes); L.pstride = pstride;
  L.it = 0; L.max_iter = o.max_iter; L.rtol = o.rtol;
  pcg_init_linked_kernel<UT><<<grid_u, UT, 0, h->stream>>>(d_b, h->Dinv.p, h->x.p, h->r.p, h->z.p, h->p.p, h->q.p, n, L);
  h->launches++;
  FEMB_CUDA(h, cudaGetLastError());
  PcgPeek* peek = reinterpret_cast<PcgPeek*>(h->pinned);
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
      pcg_update_linked_kernel<UT><<<grid_u, UT, 0, h->stream>>>(h->Dinv.p, h->s.p, h->p.p, h->q.p, h->x.p, h->r.p, h->z.p, n, L);
      if (timed) { cudaEventRecord(e2, h->stream); evs.push_back(e0); evs.push_back(e1); evs.push_back(e2); }
      h->launches++;
    }
    L.it = it;
    pcg_decide_linked_kernel<128><<<1, 128, 0, h->stream>>>(L);
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
    st->precond_used = (o.precond == FEMB_PRECOND_NONE) ? FEMB_PRECOND_NONE : FEMB_PRECOND_JACOBI;
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

static bool linked_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("FEMB_PCG_LINKED"); on = (e && e[0] == '0') ? 0 : 1; }
  return on != 0;
}

static int pcg_core(femb_handle* h, const femb_solve_opts& o, const double* d_b, femb_stats* st) {
  if (lines_applicable(h, o)) return pcg_lines(h, o, d_b, st);
  if (twolevel_applicable(h, o)) return pcg_twolevel(h, o, d_b, st);
  if (h->bs == 6 && o.precond != FEMB_PRECOND_BLOCK_JACOBI && ebe_selected(h, o.op) && linked_enabled())
    return pcg_core_linked(h, o, d_b, st);
  const int64_t n = h->ndof;
  const int gridv = vec_grid(h, n, kRowThreads);
  const int pstride = h->num_sms * 8;
  int rc = setup_precond(h, o.precond);
  if (rc) return rc;
  FEMB_CUDA(h, cudaMemsetAsync(h->flags.p, 0, sizeof(int32_t) * Flag::COUNT, h->stream));
  const bool blockj = (o.precond == FEMB_PRECOND_BLOCK_JACOBI);
#define INIT(BS, BJ) pcg_init_kernel<BS, kRowThreads, BJ><<<gridv, kRowThreads, 0, h->stream>>>(d_b, h->Dinv.p, h->x.p, h->r.p, h->z.p, h->p.p, h->q.p, n, o.rtol, h->partials.p, pstride, h->scal.p, h->flags.p)
  if (h->bs == 6) { if (blockj) INIT(6, true); else INIT(6, false); }
  else { if (blockj) INIT(3, true); else INIT(3, false); }
#undef INIT
  h->launches++;
  FEMB_CUDA(h, cudaGetLastError());

  PcgPeek* peek = reinterpret_cast<PcgPeek*>(h->pinned);
  const int check = o.check_every > 0 ? o.check_every : 50;
  const bool prof = o.profile != 0;
  const bool ebe = ebe_selected(h, o.op);   // matrix-free frame operator (ebe.cu) instead of the BSR SpMV
  std::vector<cudaEvent_t> evs;
  int spmv_launches = 0;
  int it = 0;
  int done = 0;
  while (!done && it < o.max_iter) {
    const int batch = (o.max_iter - it) < check ? (o.max_iter - it) : check;
    for (int k = 0; k < batch; ++k, ++it) {
      const int parity = it & 1;
      const bool timed = prof && (it % o.profile) == 0;
      cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
      if (timed) {
        // events come from a pool that persists in the handle: after the warm-up solves no event
        // is created inside a timed step
        if (h->ev_pool.size() < evs.size() + 3) {
          const size_t old = h->ev_pool.size();
          h->ev_pool.resize(old + 768);
          for (size_t k = old; k < h->ev_pool.size(); ++k) cudaEventCreate(&h->ev_pool[k]);
        }
        e0 = h->ev_pool[evs.size()]; e1 = h->ev_pool[evs.size() + 1]; e2 = h->ev_pool[evs.size() + 2];
        cudaEventRecord(e0, h->stream);
      }
      if (ebe) rc = launch_ebe(h, h->z.p, h->s.p, 1, true, h->partials.p, h->scal.p + Scal::PQ, h->flags.p + Flag::TICKET0,
                               h->flags.p + Flag::DONE, nullptr);
      else rc = launch_spmv(h, h->z.p, h->s.p, true, h->partials.p);
      if (timed) cudaEventRecord(e1, h->stream);
      if (rc) return rc;
      ++spmv_launches;
#define UPD(BS, BJ) pcg_update_kernel<BS, kRowThreads, BJ><<<occ_grid(h, pcg_update_kernel<BS, kRowThreads, BJ>, BJ ? n : (n + 1) / 2, kRowThreads), kRowThreads, 0, h->stream>>>(h->Dinv.p, h->s.p, h->p.p, h->q.p, h->x.p, h->r.p, h->z.p, n, parity, it == 0 ? 1 : 0, o.max_iter, h->partials.p + pstride, pstride, h->scal.p, h->flags.p)
      if (h->bs == 6) { if (blockj) UPD(6, true); else UPD(6, false); }
      else { if (blockj) UPD(3, true); else UPD(3, false); }
#undef UPD
      if (timed) { cudaEventRecord(e2, h->stream); evs.push_back(e0); evs.push_back(e1); evs.push_back(e2); }
      h->launches += 1;
    }
    FEMB_CUDA(h, cudaGetLastError());
    FEMB_CUDA(h, cudaMemcpyAsync(peek->flags, h->flags.p, sizeof(peek->flags), cudaMemcpyDeviceToHost, h->stream));
    FEMB_CUDA(h, cudaMemcpyAsync(peek->scal, h->scal.p, sizeof(peek->scal), cudaMemcpyDeviceToHost, h->stream));
    FEMB_CUDA(h, cudaStreamSynchronize(h->stream));
    done = peek->flags[Flag::DONE];
  }
  FEMB_CUDA(h, cudaMemcpyAsync(peek->flags, h->flags.p, sizeof(peek->flags), cudaMemcpyDeviceToHost, h->stream));
  FEMB_CUDA(h, cudaMemcpyAsync(peek->scal, h->scal.p, sizeof(peek->scal), cudaMemcpyDeviceToHost, h->stream));
  FEMB_CUDA(h, cudaStreamSynchronize(h->stream));
  done = peek->flags[Flag::DONE];
  if (st) {
    st->method_used = FEMB_SOLVER_PCG;
    st->op_used = ebe ? FEMB_OP_EBE : FEMB_OP_BSR;
    st->precond_used = (o.precond == FEMB_PRECOND_NONE || o.precond == FEMB_PRECOND_BLOCK_JACOBI) ? o.precond : FEMB_PRECOND_JACOBI;
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
    st->spmv_timed = (int32_t)(evs.size() / 3);  // number of iterations whose two kernels were timed
  }
  if (done == 2) return fail(h, FEMB_ERR_SINGULAR, "PCG breakdown: p^T K p <= 0 (K_ff is not positive definite — unconstrained rigid-body motion or zero section properties?)");
  if (done != 1) return fail(h, FEMB_ERR_NOT_CONVERGED, "PCG did not reach rtol within max_iter");
  return FEMB_OK;
}

int setup_rhs_for_direct(femb_handle* h) { return build_rhs(h, h->u0.p != nullptr); }
int setup_precond_public(femb_handle* h, int mode) { return setup_precond(h, mode); }

// x[fixed] = prescribed value (the reference always prescribes 0: BeamSolver.py:413,418)
int apply_prescribed(femb_handle* h) {
  if (!h->u0.p) return FEMB_OK;
  set_prescribed_kernel<<<vec_grid(h, h->ndof, kVecThreads), kVecThreads, 0, h->stream>>>(h->x.p, h->u0.p, h->free_mask.p, h->ndof);
  h->launches++;
  FEMB_CUDA(h, cudaGetLastError());
  return FEMB_OK;
}

int launch_reactions(femb_handle* h, bool minus_f, double* d_out) {
  int rc = launch_spmv(h, h->x.p, d_out, false, nullptr);
  if (rc) return rc;
  if (minus_f) {
    axpy_sub_kernel<<<vec_grid(h, h->ndof, kVecThreads), kVecThreads, 0, h->stream>>>(d_out, h->f.p, h->ndof);
    h->launches++;
    FEMB_CUDA(h, cudaGetLastError());
  }
  return FEMB_OK;
}

int bc_build_mask(femb_handle* h, const int64_t* d_fixed, int64_t n_fixed) {
  const int64_t n = h->ndof;
  fill_mask_kernel<<<vec_grid(h, n, kVecThreads), kVecThreads, 0, h->stream>>>(h->free_mask.p, n, 1);
  if (n_fixed > 0)
    clear_fixed_kernel<<<vec_grid(h, n_fixed, kVecThreads), kVecThreads, 0, h->stream>>>(h->free_mask.p, d_fixed, n_fixed);
  h->launches += 2;
  FEMB_CUDA(h, cudaGetLastError());
  return FEMB_OK;
}

}  // namespace femb

// ------------------------------------------------------------- multi-RHS PCG (NB = 2 or 4)
// NB right-hand sides advance in lockstep through independent CG recurrences that share ONE
// operator pass per iteration: the matrix-free kernel rebuilds each element record once for NB
// vectors (or, with the assembled operator, the 359 MB of K are read once: SpMM).  Work vectors are
// interleaved by right-hand side, v[g*NB + q], so the gather of a block column is one contiguous read
// and every vector kernel is a plain 16- / 32-byte stream.  z = Dinv r is formed on the fly (scalar
// Jacobi).  Used by the modal solver's shift-invert steps (block Krylov, block size 2 or 4).
namespace femb {

struct MScal {  // doubles, each [4] (NB = 2 uses the first two slots)
  enum { PQ = 0, RZ = 4, RR = 8, BB = 12, TOL2 = 16, BETA = 20, COUNT = 24 };
};
struct MFlag {  // ints
  enum { DONEQ = 0, ALLDONE = 4, ITERS = 5, BAD = 6, TICKET0 = 7, TICKET1 = 8, COUNT = 12 };
};

template <int NB>
__device__ __forceinline__ void ldv(const double* p, double (&v)[NB]) {
  if constexpr (NB == 4) {
    const double4 t = *reinterpret_cast<const double4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
    const double2 t = *reinterpret_cast<const double2*>(p);
    v[0] = t.x; v[1] = t.y;
  }
}
template <int NB>
__device__ __forceinline__ void stv(double* p, const double (&v)[NB]) {
  if constexpr (NB == 4) *reinterpret_cast<double4*>(p) = make_double4(v[0], v[1], v[2], v[3]);
  else *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
}
template <int NB>
__device__ __forceinline__ void ldgv(const double* p, double (&v)[NB]) {   // read-only path (LDG.CONSTANT)
  if constexpr (NB == 4) {
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
  } else {
    asm("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "l"(p));
  }
}

// x = 0, r = b (interleaved from the nb column vectors; missing columns are zero), p = Dinv r
template <int NB, int THREADS>
__global__ void __launch_bounds__(THREADS)
mpcg_init_kernel(const double* __restrict__ B, int64_t ldb, int nb, const double* __restrict__ dinv,
                 double* __restrict__ x, double* __restrict__ r, double* __restrict__ p, int64_t n, double rtol,
                 double* partials, int pstride, double* scal, int* flags) {
  double mine[2 * NB];
#pragma unroll
  for (int q = 0; q < 2 * NB; ++q) mine[q] = 0.0;
  for (int64_t g = (int64_t)blockIdx.x * THREADS + threadIdx.x; g < n; g += (int64_t)gridDim.x * THREADS) {
    const double d = dinv[g];
    double b[NB], z0[NB], pz[NB];
#pragma unroll
    for (int q = 0; q < NB; ++q) { b[q] = q < nb ? B[(size_t)q * ldb + g] : 0.0; z0[q] = 0.0; pz[q] = d * b[q]; }
    stv<NB>(x + g * NB, z0);
    stv<NB>(r + g * NB, b);
    stv<NB>(p + g * NB, pz);
#pragma unroll
    for (int q = 0; q < NB; ++q) { mine[q] += b[q] * d * b[q]; mine[NB + q] += b[q] * b[q]; }
  }
  double tot[2 * NB];
  if (grid_reduce<THREADS, 2 * NB>(mine, partials, pstride, flags + MFlag::TICKET1, tot)) {
    if (threadIdx.x == 0) {
      int all = 1;
      for (int q = 0; q < 4; ++q) flags[MFlag::DONEQ + q] = 1;   // slots beyond NB stay "done"
      for (int q = 0; q < NB; ++q) {
        scal[MScal::RZ + q] = tot[q];
        scal[MScal::BB + q] = tot[NB + q];
        scal[MScal::RR + q] = tot[NB + q];
        scal[MScal::TOL2 + q] = rtol * rtol * tot[NB + q];
        scal[MScal::BETA + q] = 0.0;
        const int dq = (tot[NB + q] == 0.0) ? 1 : 0;
        flags[MFlag::DONEQ + q] = dq;
        all &= dq;
      }
      flags[MFlag::ALLDONE] = all;
      flags[MFlag::ITERS] = 0;
      flags[MFlag::BAD] = 0;
    }
  }
}

// q = A p for the NB interleaved vectors (masked assembled operator), pq_j = (p_j, q_j)
template <int NB, int THREADS>
__global__ void __launch_bounds__(THREADS)
mpcg_spmm_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                 const double* __restrict__ vals, const uint8_t* __restrict__ free_mask,
                 const double* __restrict__ p, double* __restrict__ qv, int64_t n,
                 double* partials, int pstride, double* scal, int* flags) {
  if (flags[MFlag::ALLDONE]) return;
  double dot[NB];
#pragma unroll
  for (int q = 0; q < NB; ++q) dot[q] = 0.0;
  for (int64_t g = (int64_t)blockIdx.x * THREADS + threadIdx.x; g < n; g += (int64_t)gridDim.x * THREADS) {
    const int node = (int)(g / 6);
    const int r = (int)(g - (int64_t)node * 6);
    const int b0 = __ldg(rowptr + node), b1 = __ldg(rowptr + node + 1);
    double acc[NB];
#pragma unroll
    for (int q = 0; q < NB; ++q) acc[q] = 0.0;
#pragma unroll 2
    for (int b = b0; b < b1; ++b) {
      const int col = __ldg(colidx + b);
      const double2* a2 = reinterpret_cast<const double2*>(vals + (size_t)b * 36 + r * 6);
      const double2 a01 = __ldcs(a2), a23 = __ldcs(a2 + 1), a45 = __ldcs(a2 + 2);
      const double a[6] = {a01.x, a01.y, a23.x, a23.y, a45.x, a45.y};
      const double* pc = p + (size_t)col * 6 * NB;
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        double v[NB];
        ldgv<NB>(pc + c * NB, v);
#pragma unroll
        for (int q = 0; q < NB; ++q) acc[q] += a[c] * v[q];
      }
    }
    double pg[NB];
    ldv<NB>(p + g * NB, pg);
    if (!free_mask[g]) {
#pragma unroll
      for (int q = 0; q < NB; ++q) acc[q] = pg[q];
    }
    stv<NB>(qv + g * NB, acc);
#pragma unroll
    for (int q = 0; q < NB; ++q) dot[q] += pg[q] * acc[q];
  }
  double tot[NB];
  if (grid_reduce<THREADS, NB>(dot, partials, pstride, flags + MFlag::TICKET0, tot)) {
    if (threadIdx.x == 0)
      for (int q = 0; q < NB; ++q) scal[MScal::PQ + q] = tot[q];
  }
}

// x += alpha p, r -= alpha q; rz_new = (r, Dinv r), rr = (r, r); per-vector convergence, beta
template <int NB, int THREADS>
__global__ void __launch_bounds__(THREADS)
mpcg_update_xr_kernel(const double* __restrict__ dinv, const double* __restrict__ p, const double* __restrict__ qv,
                      double* __restrict__ x, double* __restrict__ r, int64_t n, int max_iter,
                      double* partials, int pstride, double* scal, int* flags) {
  if (flags[MFlag::ALLDONE]) return;
  double alpha[NB];
  bool bad = false;
#pragma unroll
  for (int q = 0; q < NB; ++q) {
    const double pq = scal[MScal::PQ + q];
    const bool dq = flags[MFlag::DONEQ + q] != 0;
    if (!dq && !(pq > 0.0)) bad = true;
    alpha[q] = (dq || !(pq > 0.0)) ? 0.0 : scal[MScal::RZ + q] / pq;
  }
  double mine[2 * NB];
#pragma unroll
  for (int q = 0; q < 2 * NB; ++q) mine[q] = 0.0;
  for (int64_t g = (int64_t)blockIdx.x * THREADS + threadIdx.x; g < n; g += (int64_t)gridDim.x * THREADS) {
    const double d = __ldg(dinv + g);
    double pv[NB], qq[NB], xv[NB], rv[NB];
    ldv<NB>(p + g * NB, pv); ldv<NB>(qv + g * NB, qq); ldv<NB>(x + g * NB, xv); ldv<NB>(r + g * NB, rv);
#pragma unroll
    for (int q = 0; q < NB; ++q) { xv[q] += alpha[q] * pv[q]; rv[q] -= alpha[q] * qq[q]; }
    stv<NB>(x + g * NB, xv);
    stv<NB>(r + g * NB, rv);
#pragma unroll
    for (int q = 0; q < NB; ++q) { mine[q] += rv[q] * d * rv[q]; mine[NB + q] += rv[q] * rv[q]; }
  }
  double tot[2 * NB];
  if (grid_reduce<THREADS, 2 * NB>(mine, partials, pstride, flags + MFlag::TICKET1, tot)) {
    if (threadIdx.x == 0) {
      const int it = flags[MFlag::ITERS] + 1;
      flags[MFlag::ITERS] = it;
      int all = 1;
      for (int q = 0; q < NB; ++q) {
        if (flags[MFlag::DONEQ + q]) continue;     // frozen: its scalars keep the converged values
        const double rz_old = scal[MScal::RZ + q];
        scal[MScal::RZ + q] = tot[q];
        scal[MScal::RR + q] = tot[NB + q];
        scal[MScal::BETA + q] = rz_old > 0.0 ? tot[q] / rz_old : 0.0;
        if (tot[NB + q] <= scal[MScal::TOL2 + q]) flags[MFlag::DONEQ + q] = 1;
        else all = 0;
      }
      if (bad) { flags[MFlag::BAD] = 1; all = 1; }
      if (it >= max_iter) all = 1;
      flags[MFlag::ALLDONE] = all;
    }
  }
}

// p = Dinv r + beta p (vectors that are done keep their p: it is never used again)
template <int NB, int THREADS>
__global__ void __launch_bounds__(THREADS)
mpcg_update_p_kernel(const double* __restrict__ dinv, const double* __restrict__ r, double* __restrict__ p, int64_t n,
                     const double* scal, const int* flags) {
  if (flags[MFlag::ALLDONE]) return;
  double beta[NB];
#pragma unroll
  for (int q = 0; q < NB; ++q) beta[q] = scal[MScal::BETA + q];
  for (int64_t g = (int64_t)blockIdx.x * THREADS + threadIdx.x; g < n; g += (int64_t)gridDim.x * THREADS) {
    const double d = __ldg(dinv + g);
    double rv[NB], pv[NB];
    ldv<NB>(r + g * NB, rv); ldv<NB>(p + g * NB, pv);
#pragma unroll
    for (int q = 0; q < NB; ++q) pv[q] = d * rv[q] + beta[q] * pv[q];
    stv<NB>(p + g * NB, pv);
  }
}

// ---- linked lockstep iteration (matrix-free operator): the operator publishes its (p_q, A p_q)
// partials, the x/r kernel consumes them and publishes {rz_q, rr_q}, the p kernel consumes those,
// decides per-vector convergence and beta.  No ticket / last-CTA tails; rz and the per-vector done
// flags travel through parity-indexed slots (slot it & 1 is read in iteration it, (it+1) & 1 written).
struct MLink {
  double* op_partials;    // [NB][pstride]
  double* xr_partials;    // [2 buffers][2 NB][pstride]
  double* scal;           // MScal layout; RZ lives in two parity slots: RZ + 4 * parity ... see MLScal
  int* flags;
  int n_op, n_xr, pstride;
  int it, max_iter;
};
struct MLScal { enum { RZ0 = 0, RZ1 = 4, RR = 8, BB = 12, TOL2 = 16, COUNT = 24 }; };
struct MLFlag { enum { DONEQ0 = 0, DONEQ1 = 4, ALLDONE = 8, ITERS = 9, BAD = 10, COUNT = 12 }; };

template <int NB, int THREADS>
__global__ void __launch_bounds__(THREADS)
mpcg_init_linked_kernel(const double* __restrict__ B, int64_t ldb, int nb, const double* __restrict__ dinv,
                        double* __restrict__ x, double* __restrict__ r, double* __restrict__ p, int64_t n, double rtol,
                        double* partials, int pstride, int* ticket, double* scal, int* flags) {
  double mine[2 * NB];
#pragma unroll
  for (int q = 0; q < 2 * NB; ++q) mine[q] = 0.0;
  for (int64_t g = (int64_t)blockIdx.x * THREADS + threadIdx.x; g < n; g += (int64_t)gridDim.x * THREADS) {
    const double d = dinv[g];
    double b[NB], z0[NB], pz[NB];
#pragma unroll
    for (int q = 0; q < NB; ++q) { b[q] = q < nb ? B[(size_t)q * ldb + g] : 0.0; z0[q] = 0.0; pz[q] = d * b[q]; }
    stv<NB>(x + g * NB, z0);
    stv<NB>(r + g * NB, b);
    stv<NB>(p + g * NB, pz);
#pragma unroll
    for (int q = 0; q < NB; ++q) { mine[q] += b[q] * d * b[q]; mine[NB + q] += b[q] * b[q]; }
  }
  double tot[2 * NB];
  if (grid_reduce<THREADS, 2 * NB>(mine, partials, pstride, ticket, tot)) {   // once per solve
    if (threadIdx.x == 0) {
      int all = 1;
      for (int q = 0; q < 4; ++q) { flags[MLFlag::DONEQ0 + q] = 1; flags[MLFlag::DONEQ1 + q] = 1; }
      for (int q = 0; q < NB; ++q) {
        scal[MLScal::RZ0 + q] = tot[q];
        scal[MLScal::BB + q] = tot[NB + q];
        scal[MLScal::RR + q] = tot[NB + q];
        scal[MLScal::TOL2 + q] = rtol * rtol * tot[NB + q];
        const int dq = (tot[NB + q] == 0.0) ? 1 : 0;
        flags[MLFlag::DONEQ0 + q] = dq;
        all &= dq;
      }
      flags[MLFlag::ALLDONE] = all;
      flags[MLFlag::ITERS] = 0;
      flags[MLFlag::BAD] = 0;
    }
  }
}

template <int NB, int THREADS>
__global__ void __launch_bounds__(THREADS)
mpcg_update_xr_linked_kernel(const double* __restrict__ dinv, const double* __restrict__ p, const double* __restrict__ qv,
                             double* __restrict__ x, double* __restrict__ r, int64_t n, const MLink L) {
  __shared__ double s_part[2 * NB * THREADS / 32];
  if (L.flags[MLFlag::ALLDONE]) return;
  const int par = L.it & 1;
  double pq[NB];
#pragma unroll
  for (int q = 0; q < NB; ++q) pq[q] = 0.0;
  for (int i = threadIdx.x; i < L.n_op; i += THREADS) {
#pragma unroll
    for (int q = 0; q < NB; ++q) pq[q] += __ldcg(L.op_partials + (size_t)q * L.pstride + i);
  }
  block_sum_all<THREADS, NB>(pq, s_part);
  double alpha[NB];
  bool bad = false;
#pragma unroll
  for (int q = 0; q < NB; ++q) {
    const bool dq = L.flags[MLFlag::DONEQ0 + 4 * par + q] != 0;
    if (!dq && !(pq[q] > 0.0)) bad = true;
    alpha[q] = (dq || !(pq[q] > 0.0)) ? 0.0 : L.scal[MLScal::RZ0 + 4 * par + q] / pq[q];
  }
  if (bad) {                                  // K_ff not positive definite along some p_q
    if (blockIdx.x == 0 && threadIdx.x == 0) { L.flags[MLFlag::BAD] = 1; L.flags[MLFlag::ALLDONE] = 1; }
    return;
  }
  double mine[2 * NB];
#pragma unroll
  for (int q = 0; q < 2 * NB; ++q) mine[q] = 0.0;
  for (int64_t g = (int64_t)blockIdx.x * THREADS + threadIdx.x; g < n; g += (int64_t)gridDim.x * THREADS) {
    const double d = __ldg(dinv + g);
    double pv[NB], qq[NB], xv[NB], rv[NB];
    ldv<NB>(p + g * NB, pv); ldv<NB>(qv + g * NB, qq); ldv<NB>(x + g * NB, xv); ldv<NB>(r + g * NB, rv);
#pragma unroll
    for (int q = 0; q < NB; ++q) { xv[q] += alpha[q] * pv[q]; rv[q] -= alpha[q] * qq[q]; }
    stv<NB>(x + g * NB, xv);
    stv<NB>(r + g * NB, rv);
#pragma unroll
    for (int q = 0; q < NB; ++q) { mine[q] += rv[q] * d * rv[q]; mine[NB + q] += rv[q] * rv[q]; }
  }
  block_sum_all<THREADS, 2 * NB>(mine, s_part);
  if (threadIdx.x == 0) {
    double* po = L.xr_partials + (size_t)par * 2 * NB * L.pstride;
#pragma unroll
    for (int v = 0; v < 2 * NB; ++v) po[(size_t)v * L.pstride + blockIdx.x] = mine[v];
  }
}

template <int NB, int THREADS>
__global__ void __launch_bounds__(THREADS)
mpcg_update_p_linked_kernel(const double* __restrict__ dinv, const double* __restrict__ r, double* __restrict__ p, int64_t n,
                            const MLink L) {
  __shared__ double s_part[2 * NB * THREADS / 32];
  if (L.flags[MLFlag::ALLDONE]) return;
  const int par = L.it & 1;
  double tot[2 * NB];
#pragma unroll
  for (int v = 0; v < 2 * NB; ++v) tot[v] = 0.0;
  {
    const double* pi = L.xr_partials + (size_t)par * 2 * NB * L.pstride;
    for (int i = threadIdx.x; i < L.n_xr; i += THREADS) {
#pragma unroll
      for (int v = 0; v < 2 * NB; ++v) tot[v] += __ldcg(pi + (size_t)v * L.pstride + i);
    }
  }
  block_sum_all<THREADS, 2 * NB>(tot, s_part);
  double beta[NB];
  int all = 1;
  const bool lead = blockIdx.x == 0 && threadIdx.x == 0;
#pragma unroll
  for (int q = 0; q < NB; ++q) {
    const bool dq = L.flags[MLFlag::DONEQ0 + 4 * par + q] != 0;
    const double rz_old = L.scal[MLScal::RZ0 + 4 * par + q];
    beta[q] = (!dq && rz_old > 0.0) ? tot[q] / rz_old : 0.0;
    const int ndq = (dq || tot[NB + q] <= L.scal[MLScal::TOL2 + q]) ? 1 : 0;
    all &= ndq;
    if (lead) {   // frozen vectors keep their converged scalars
      L.scal[MLScal::RZ0 + 4 * (par ^ 1) + q] = dq ? rz_old : tot[q];
      if (!dq) L.scal[MLScal::RR + q] = tot[NB + q];
      L.flags[MLFlag::DONEQ0 + 4 * (par ^ 1) + q] = ndq;
    }
  }
  if (L.it + 1 >= L.max_iter) all = 1;
  if (lead) { L.flags[MLFlag::ITERS] = L.it + 1; if (all) L.flags[MLFlag::ALLDONE] = 1; }
  for (int64_t g = (int64_t)blockIdx.x * THREADS + threadIdx.x; g < n; g += (int64_t)gridDim.x * THREADS) {
    const double d = __ldg(dinv + g);
    double rv[NB], pv[NB];
    ldv<NB>(r + g * NB, rv); ldv<NB>(p + g * NB, pv);
#pragma unroll
    for (int q = 0; q < NB; ++q) pv[q] = d * rv[q] + beta[q] * pv[q];
    stv<NB>(p + g * NB, pv);
  }
}

template <int NB>
__global__ void mpcg_extract_kernel(const double* __restrict__ x, double* __restrict__ X, int64_t ldx, int nb, int64_t n) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  double e[NB];
  ldv<NB>(x + g * NB, e);
  for (int q = 0; q < nb && q < NB; ++q) X[(size_t)q * ldx + g] = e[q];
}

// K_ff X = B for nb <= NB (already masked) right-hand sides, scalar-Jacobi PCG in lockstep.
// st->iterations receives the lockstep iteration count, st->spmv_launches the operator launches.
template <int NB>
static int pcg_solve_multi_linked(femb_handle* h, const femb_solve_opts& o, const double* d_B, int64_t ldb, int nb, double* d_X,
                                  int64_t ldx, femb_stats* st) {
  const int64_t n = h->ndof;
  const int gridv = vec_grid(h, n, kRowThreads);
  const int pstride = h->num_sms * 8;
  int rc = setup_precond(h, FEMB_PRECOND_JACOBI);
  if (rc) return rc;
  FEMB_CUDA(h, h->mx.ensure((size_t)n * NB));
  FEMB_CUDA(h, h->mr.ensure((size_t)n * NB));
  FEMB_CUDA(h, h->mp.ensure((size_t)n * NB));
  FEMB_CUDA(h, h->mq.ensure((size_t)n * NB));
  FEMB_CUDA(h, h->mpartials.ensure((size_t)pstride * (4 + 16 + 8)));   // operator [4], x/r [2][8], init [8]
  FEMB_CUDA(h, h->mscal.ensure(MLScal::COUNT));
  FEMB_CUDA(h, h->mflags.ensure(MLFlag::COUNT + 1));
  FEMB_CUDA(h, cudaMemsetAsync(h->mflags.p, 0, sizeof(int32_t) * (MLFlag::COUNT + 1), h->stream));
  constexpr int UT = 256;
  const int grid_xr = occ_grid(h, mpcg_update_xr_linked_kernel<NB, UT>, n, UT);
  const int grid_p = occ_grid(h, mpcg_update_p_linked_kernel<NB, UT>, n, UT);
  const int max_iter = o.max_iter > 0 ? o.max_iter : 200000;
  MLink L;
  L.op_partials = h->mpartials.p; L.xr_partials = h->mpartials.p + (size_t)4 * pstride;
  L.scal = h->mscal.p; L.flags = h->mflags.p;
  L.n_op = ebe_grid(h, NB, h->n_nodes); L.n_xr = grid_xr; L.pstride = pstride; L.it = 0; L.max_iter = max_iter;
  mpcg_init_linked_kernel<NB, kRowThreads><<<gridv, kRowThreads, 0, h->stream>>>(
      d_B, ldb, nb, h->Dinv.p, h->mx.p, h->mr.p, h->mp.p, n, o.rtol, h->mpartials.p + (size_t)20 * pstride, pstride,
      h->mflags.p + MLFlag::COUNT, h->mscal.p, h->mflags.p);
  h->launches++;
  FEMB_CUDA(h, cudaGetLastError());
  struct Peek { int32_t flags[MLFlag::COUNT]; double scal[MLScal::COUNT]; };
  Peek* peek = reinterpret_cast<Peek*>(h->pinned);
  const int check = o.check_every > 0 ? o.check_every : 50;
  PcgLink opl = {};                 // the operator kernel only needs where to publish
  opl.op_partials = L.op_partials; opl.pstride = pstride; opl.flags = h->mflags.p;
  int it = 0, all = 0, spmm = 0;
  while (!all && it < max_iter) {
    const int batch = std::min(check, max_iter - it);
    for (int k = 0; k < batch; ++k, ++it) {
      L.it = it;
      rc = launch_ebe(h, h->mp.p, h->mq.p, NB, true, nullptr, nullptr, nullptr, h->mflags.p + MLFlag::ALLDONE, &opl);
      if (rc) return rc;
      mpcg_update_xr_linked_kernel<NB, UT><<<grid_xr, UT, 0, h->stream>>>(h->Dinv.p, h->mp.p, h->mq.p, h->mx.p, h->mr.p, n, L);
      mpcg_update_p_linked_kernel<NB, UT><<<grid_p, UT, 0, h->stream>>>(h->Dinv.p, h->mr.p, h->mp.p, n, L);
      h->launches += 2;
      ++spmm;
    }
    FEMB_CUDA(h, cudaGetLastError());
    FEMB_CUDA(h, cudaMemcpyAsync(peek->flags, h->mflags.p, sizeof(peek->flags), cudaMemcpyDeviceToHost, h->stream));
    FEMB_CUDA(h, cudaMemcpyAsync(peek->scal, h->mscal.p, sizeof(peek->scal), cudaMemcpyDeviceToHost, h->stream));
    FEMB_CUDA(h, cudaStreamSynchronize(h->stream));
    all = peek->flags[MLFlag::ALLDONE];
  }
  mpcg_extract_kernel<NB><<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->mx.p, d_X, ldx, nb, n);
  h->launches++;
  FEMB_CUDA(h, cudaGetLastError());
  const int fin = peek->flags[MLFlag::ITERS] & 1;   // the done flags the last p kernel wrote
  bool conv = true;
  double worst = 0.0;
  for (int q = 0; q < nb; ++q) {
    conv = conv && peek->flags[MLFlag::DONEQ0 + 4 * fin + q] != 0;
    const double bb = peek->scal[MLScal::BB + q];
    if (bb > 0.0) worst = std::max(worst, std::sqrt(peek->scal[MLScal::RR + q] / bb));
  }
  if (st) {
    st->method_used = FEMB_SOLVER_PCG;
    st->op_used = FEMB_OP_EBE;
    st->iterations = peek->flags[MLFlag::ITERS];
    st->spmv_launches = spmm;
    st->converged = conv ? 1 : 0;
    st->rel_residual = worst;
  }
  if (peek->flags[MLFlag::BAD]) return fail(h, FEMB_ERR_SINGULAR, "multi-RHS PCG breakdown: p^T K p <= 0 (K_ff is not positive definite)");
  if (!conv) return fail(h, FEMB_ERR_NOT_CONVERGED, "multi-RHS PCG did not reach rtol within max_iter");
  return FEMB_OK;
}

template <int NB>
static int pcg_solve_multi_t(femb_handle* h, const femb_solve_opts& o, const double* d_B, int64_t ldb, int nb, double* d_X,
                             int64_t ldx, femb_stats* st) {
  const int64_t n = h->ndof;
  const int gridv = vec_grid(h, n, kRowThreads);
  const int pstride = h->num_sms * 8;
  int rc = setup_precond(h, FEMB_PRECOND_JACOBI);
  if (rc) return rc;
  FEMB_CUDA(h, h->mx.ensure((size_t)n * NB));
  FEMB_CUDA(h, h->mr.ensure((size_t)n * NB));
  FEMB_CUDA(h, h->mp.ensure((size_t)n * NB));
  FEMB_CUDA(h, h->mq.ensure((size_t)n * NB));
  FEMB_CUDA(h, h->mpartials.ensure((size_t)pstride * 3 * 4));
  FEMB_CUDA(h, h->mscal.ensure(MScal::COUNT));
  FEMB_CUDA(h, h->mflags.ensure(MFlag::COUNT));
  FEMB_CUDA(h, cudaMemsetAsync(h->mflags.p, 0, sizeof(int32_t) * MFlag::COUNT, h->stream));
  double* part0 = h->mpartials.p;                          // operator: NB arrays
  double* part1 = h->mpartials.p + (size_t)pstride * 4;    // init / update: 2*NB arrays
  const int max_iter = o.max_iter > 0 ? o.max_iter : 200000;
  mpcg_init_kernel<NB, kRowThreads><<<gridv, kRowThreads, 0, h->stream>>>(d_B, ldb, nb, h->Dinv.p, h->mx.p, h->mr.p, h->mp.p, n,
                                                                           o.rtol, part1, pstride, h->mscal.p, h->mflags.p);
  h->launches++;
  FEMB_CUDA(h, cudaGetLastError());
  struct Peek { int32_t flags[MFlag::COUNT]; double scal[MScal::COUNT]; };
  Peek* peek = reinterpret_cast<Peek*>(h->pinned);
  const int check = o.check_every > 0 ? o.check_every : 50;
  const bool ebe = ebe_selected(h, o.op);
  const int grid_xr = occ_grid(h, mpcg_update_xr_kernel<NB, kRowThreads>, n, kRowThreads);
  const int grid_mm = occ_grid(h, mpcg_spmm_kernel<NB, kRowThreads>, n, kRowThreads);
  int it = 0, all = 0, spmm = 0;
  while (!all && it < max_iter) {
    const int batch = std::min(check, max_iter - it);
    for (int k = 0; k < batch; ++k, ++it) {
      if (ebe) {
        rc = launch_ebe(h, h->mp.p, h->mq.p, NB, true, part0, h->mscal.p + MScal::PQ, h->mflags.p + MFlag::TICKET0,
                        h->mflags.p + MFlag::ALLDONE, nullptr);
        if (rc) return rc;
        h->launches--;   // counted with the two update kernels below
      } else {
        mpcg_spmm_kernel<NB, kRowThreads><<<grid_mm, kRowThreads, 0, h->stream>>>(h->rowptr.p, h->colidx.p, h->Kvals.p, h->free_mask.p,
                                                                                   h->mp.p, h->mq.p, n, part0, pstride, h->mscal.p, h->mflags.p);
      }
      mpcg_update_xr_kernel<NB, kRowThreads><<<grid_xr, kRowThreads, 0, h->stream>>>(h->Dinv.p, h->mp.p, h->mq.p, h->mx.p, h->mr.p, n,
                 max_iter, part1, pstride, h->mscal.p, h->mflags.p);
      mpcg_update_p_kernel<NB, kRowThreads><<<gridv, kRowThreads, 0, h->stream>>>(h->Dinv.p, h->mr.p, h->mp.p, n, h->mscal.p, h->mflags.p);
      h->launches += 3;
      ++spmm;
    }
    FEMB_CUDA(h, cudaGetLastError());
    FEMB_CUDA(h, cudaMemcpyAsync(peek->flags, h->mflags.p, sizeof(peek->flags), cudaMemcpyDeviceToHost, h->stream));
    FEMB_CUDA(h, cudaMemcpyAsync(peek->scal, h->mscal.p, sizeof(peek->scal), cudaMemcpyDeviceToHost, h->stream));
    FEMB_CUDA(h, cudaStreamSynchronize(h->stream));
    all = peek->flags[MFlag::ALLDONE];
  }
  mpcg_extract_kernel<NB><<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->mx.p, d_X, ldx, nb, n);
  h->launches++;
  FEMB_CUDA(h, cudaGetLastError());
  bool conv = true;
  double worst = 0.0;
  for (int q = 0; q < nb; ++q) {
    conv = conv && peek->flags[MFlag::DONEQ + q] != 0;
    const double bb = peek->scal[MScal::BB + q];
    if (bb > 0.0) worst = std::max(worst, std::sqrt(peek->scal[MScal::RR + q] / bb));
  }
  if (st) {
    st->method_used = FEMB_SOLVER_PCG;
    st->op_used = ebe ? FEMB_OP_EBE : FEMB_OP_BSR;
    st->precond_used = (o.precond == FEMB_PRECOND_NONE || o.precond == FEMB_PRECOND_BLOCK_JACOBI) ? o.precond : FEMB_PRECOND_JACOBI;
    st->iterations = peek->flags[MFlag::ITERS];
    st->spmv_launches = spmm;
    st->converged = conv ? 1 : 0;
    st->rel_residual = worst;
  }
  if (peek->flags[MFlag::BAD]) return fail(h, FEMB_ERR_SINGULAR, "multi-RHS PCG breakdown: p^T K p <= 0 (K_ff is not positive definite)");
  if (!conv) return fail(h, FEMB_ERR_NOT_CONVERGED, "multi-RHS PCG did not reach rtol within max_iter");
  return FEMB_OK;
}

int pcg_solve_multi(femb_handle* h, const femb_solve_opts& o, const double* d_B, int64_t ldb, int nb, double* d_X,
                    int64_t ldx, femb_stats* st) {
  if (h->bs != 6 || nb < 1 || nb > 4) return fail(h, FEMB_ERR_ARG, "multi-RHS PCG: frame operator, 1..4 right-hand sides");
  if (ebe_selected(h, o.op) && linked_enabled()) {
    if (nb <= 2) return pcg_solve_multi_linked<2>(h, o, d_B, ldb, nb, d_X, ldx, st);
    return pcg_solve_multi_linked<4>(h, o, d_B, ldb, nb, d_X, ldx, st);
  }
  if (nb <= 2) return pcg_solve_multi_t<2>(h, o, d_B, ldb, nb, d_X, ldx, st);
  return pcg_solve_multi_t<4>(h, o, d_B, ldb, nb, d_X, ldx, st);
}

}  // namespace femb
