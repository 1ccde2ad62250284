// Fused element + assembly kernel (frame and Tet10) and the parity-export element kernels.
//
// One CTA owns a TILE of consecutive block rows (nodes).  The BSR values of a tile are one
// contiguous range of HBM, so the CTA builds them in shared memory and streams them out
// with a single bulk async copy (cp.async.bulk smem -> global, SASS UBLKCP) — or a fully
// coalesced 16-byte store loop when the bulk path is disabled.  Each thread evaluates ONE
// contribution (element e, local block a,b) straight from node coordinates and section
// data in registers; contributions that land in the same block are added in ascending
// list order, one per __syncthreads_or step, so the sum order is fixed by the symbolic
// phase: no float atomics, bit-reproducible run to run.  Element matrices are never
// materialised in HBM: traffic = mesh in + K (+M) out + the 12-byte/contribution map.
#include "common.cuh"
#include "elements.cuh"

namespace femb {

struct PatternDev {
  const int32_t* rowptr;
  const int32_t* contrib_ptr;
  const uint32_t* contrib;
  const int32_t* contrib_blk;
  const int32_t* tile_ptr;
};

struct FrameEl {
  static constexpr int BS = 6;
  static constexpr int BS2 = 36;
  static constexpr bool HAS_MASS = true;
  FrameParams P;
  struct Rec { FrameRec r; };
  __device__ __forceinline__ void kblock(uint32_t code, Rec& rec, double* acc) const {
    frame_record(P, code >> 2, rec.r);
    frame_kblock<true>(rec.r, (code >> 1) & 1, code & 1, acc);
  }
  __device__ __forceinline__ bool has_mass(uint32_t code) const { return ((code >> 1) & 1) == (code & 1); }
  __device__ __forceinline__ void mblock(uint32_t, const Rec& rec, double* acc) const { frame_mblock(rec.r, acc); }
};

struct Tet10El {
  static constexpr int BS = 3;
  static constexpr int BS2 = 9;
  static constexpr bool HAS_MASS = false;
  Tet10Params P;
  struct Rec {};
  __device__ __forceinline__ void kblock(uint32_t code, Rec&, double* acc) const {
    const uint32_t e = code / 100u, ab = code % 100u;
    tet10_kblock(P, e, ab / 10u, ab % 10u, ab == 0u, acc);
  }
  __device__ __forceinline__ bool has_mass(uint32_t) const { return false; }
  __device__ __forceinline__ void mblock(uint32_t, const Rec&, double*) const {}
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

template <class EL, int THREADS, bool BULK>
__global__ void __launch_bounds__(THREADS)
assemble_tiles_kernel(const EL el, const PatternDev pat, double* __restrict__ Kvals,
                      double* __restrict__ Mdiag) {
  constexpr int BS2 = EL::BS2;
  extern __shared__ __align__(128) double s_out[];  // [nblk*BS2] then mass [nnodes*BS2]
  const int tid = threadIdx.x;
  const int n0 = pat.tile_ptr[blockIdx.x], n1 = pat.tile_ptr[blockIdx.x + 1];
  const int b0 = pat.rowptr[n0], b1 = pat.rowptr[n1];
  const int c0 = pat.contrib_ptr[b0], c1 = pat.contrib_ptr[b1];
  const int nblk = b1 - b0;
  double* s_mass = s_out + (size_t)nblk * BS2;

  // blocks nobody contributes to (isolated points) must still be written as zeros
  for (int t = tid; t < nblk; t += THREADS) {
    if (pat.contrib_ptr[b0 + t + 1] == pat.contrib_ptr[b0 + t]) {
#pragma unroll
      for (int k = 0; k < BS2; ++k) s_out[(size_t)t * BS2 + k] = 0.0;
    }
  }
  if (EL::HAS_MASS) {
    for (int t = tid; t < (n1 - n0) * BS2; t += THREADS) s_mass[t] = 0.0;
  }
  __syncthreads();

  for (int base = c0; base < c1; base += THREADS) {  // normally a single pass
    const int c = base + tid;
    int myk = -1, slot = 0, mslot = -1;
    bool first = false;
    double acc[BS2];
    typename EL::Rec rec;
    uint32_t code = 0;
    if (c < c1) {
      code = pat.contrib[c];
      const int blk = pat.contrib_blk[c];
      const int cb = pat.contrib_ptr[blk];
      first = (c == cb);
      myk = c - max(cb, base);
      slot = blk - b0;
      el.kblock(code, rec, acc);
    }
    for (int k = 0; __syncthreads_or(myk >= k); ++k) {
      if (myk == k) {
        double* dst = s_out + (size_t)slot * BS2;
        if (first) {
#pragma unroll
          for (int q = 0; q < BS2; ++q) dst[q] = acc[q];
        } else {
#pragma unroll
          for (int q = 0; q < BS2; ++q) dst[q] += acc[q];
        }
      }
    }
    if (EL::HAS_MASS) {
      // second ordered pass: lumped-mass diagonal blocks (diagonal contributions only);
      // the row node of a diagonal block is found from the block's slot via rowptr.
      if (c < c1 && el.has_mass(code)) {
        el.mblock(code, rec, acc);
        const int blk = slot + b0;
        int lo = n0, hi = n1 - 1;  // largest node with rowptr[node] <= blk
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          if (pat.rowptr[mid] <= blk) lo = mid; else hi = mid - 1;
        }
        mslot = lo - n0;
      } else {
        myk = -1;
      }
      for (int k = 0; __syncthreads_or(myk >= k); ++k) {
        if (myk == k) {
          double* dst = s_mass + (size_t)mslot * BS2;
#pragma unroll
          for (int q = 0; q < BS2; ++q) dst[q] += acc[q];
        }
      }
    }
  }
  // (the last __syncthreads_or above is the barrier that publishes s_out / s_mass)

  double* gK = Kvals + (size_t)b0 * BS2;
  const int nK = nblk * BS2;
  if (BULK) {
    // generic-proxy writes -> visible to the async proxy, then one thread issues the copies
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                   :: "l"(gK), "r"(smem_u32(s_out)), "r"((uint32_t)(nK * 8)) : "memory");
      if (EL::HAS_MASS) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                     :: "l"(Mdiag + (size_t)n0 * BS2), "r"(smem_u32(s_mass)),
                        "r"((uint32_t)((n1 - n0) * BS2 * 8)) : "memory");
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  } else {
    if ((BS2 & 1) == 0) {
      const double2* s2 = reinterpret_cast<const double2*>(s_out);
      double2* g2 = reinterpret_cast<double2*>(gK);
      for (int i = tid; i < nK / 2; i += THREADS) g2[i] = s2[i];
    } else {
      for (int i = tid; i < nK; i += THREADS) gK[i] = s_out[i];
    }
    if (EL::HAS_MASS) {
      double* gM = Mdiag + (size_t)n0 * BS2;
      for (int i = tid; i < (n1 - n0) * BS2; i += THREADS) gM[i] = s_mass[i];
    }
  }
}

// ---- frame fast path: one thread per (node, incident element end) ------------------------
// The thread evaluates the element record ONCE and produces both the off-diagonal block
// (node, other end) — stored straight into the tile's staging area — and its share of the
// node's diagonal block, kept as the 21-value upper triangle (+7 compact mass values) and
// added in element-ascending order, one rank per __syncthreads_or step.  Compared with the
// generic one-thread-per-contribution kernel this halves the record evaluations and cuts
// the shared-memory accumulate traffic from 72 to 28 values per element end.
struct PairDev {
  const int32_t* rowptr;
  const int32_t* diag_blk;
  const int32_t* pair_ptr;
  const uint32_t* pair_code;
  const int32_t* pair_blk;
  const uint8_t* pair_rank;
  const int32_t* tile_ptr;
};

template <int THREADS, bool BULK>
__global__ void __launch_bounds__(THREADS)
frame_assemble_pairs_kernel(const FrameParams P, const PairDev pat, double* __restrict__ Kvals,
                            double* __restrict__ Mdiag) {
  extern __shared__ __align__(128) double s_out[];  // K blocks [nblk*36] | mass [nn*36] | diag acc [nn*28]
  const int tid = threadIdx.x;
  const int n0 = pat.tile_ptr[blockIdx.x], n1 = pat.tile_ptr[blockIdx.x + 1];
  const int nn = n1 - n0;
  const int b0 = pat.rowptr[n0], b1 = pat.rowptr[n1];
  const int p0 = pat.pair_ptr[n0], p1 = pat.pair_ptr[n1];
  const int nblk = b1 - b0;
  double* s_mass = s_out + (size_t)nblk * 36;
  double* s_diag = s_mass + (size_t)nn * 36;
  for (int t = tid; t < nn * 28; t += THREADS) s_diag[t] = 0.0;
  __syncthreads();

  for (int base = p0; base < p1; base += THREADS) {  // normally a single pass
    const int p = base + tid;
    int myk = -1, rank = -1, slot = 0, nslot = 0;
    double acc[36];
    FrameRec R;
    int a = 0;
    if (p < p1) {
      const uint32_t code = pat.pair_code[p];
      a = (int)(code & 1u);
      slot = pat.pair_blk[p] - b0;
      rank = (int)pat.pair_rank[p];
      int lo = n0, hi = n1 - 1;  // node owning pair p: largest node with pair_ptr[node] <= p
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (pat.pair_ptr[mid] <= p) lo = mid; else hi = mid - 1;
      }
      nslot = lo - n0;
      myk = p - max(pat.pair_ptr[lo], base);
      frame_record(P, code >> 1, R);
      frame_kblock<true>(R, a, 1 - a, acc);
      if (rank == 0) {
        double2* dst = reinterpret_cast<double2*>(s_out + (size_t)slot * 36);
#pragma unroll
        for (int q = 0; q < 18; ++q) dst[q] = make_double2(acc[2 * q], acc[2 * q + 1]);
      }
    }
    // duplicate members between the same two nodes (rare): ordered adds after the rank-0 store
    for (int k = 1; __syncthreads_or(rank >= k); ++k) {
      if (rank == k) {
        double* dst = s_out + (size_t)slot * 36;
#pragma unroll
        for (int q = 0; q < 36; ++q) dst[q] += acc[q];
      }
    }
    // diagonal share: 21 symmetric stiffness values + 7 compact mass values, ordered by rank
    if (p < p1) frame_diag_sym(R, a, acc, acc + 21);
    for (int k = 0; __syncthreads_or(myk >= k); ++k) {
      if (myk == k) {
        double* dst = s_diag + (size_t)nslot * 28;
#pragma unroll
        for (int q = 0; q < 28; ++q) dst[q] += acc[q];
      }
    }
  }

  // expand the symmetric accumulators into the diagonal K blocks and the dense mass blocks
  for (int t = tid; t < nn * 36; t += THREADS) {
    const int ns = t / 36, rc = t - ns * 36;
    const int r = rc / 6, c = rc - r * 6;
    const double* d = s_diag + (size_t)ns * 28;
    s_out[(size_t)(pat.diag_blk[n0 + ns] - b0) * 36 + rc] = d[sym6_index(r, c)];
    double mv = 0.0;
    if (r < 3 && c < 3) mv = (r == c) ? d[21] : 0.0;
    else if (r >= 3 && c >= 3) {
      const int rr = r - 3, cc = c - 3;
      const int lo = rr < cc ? rr : cc, hi = rr < cc ? cc : rr;
      mv = d[22 + lo * 3 - (lo * (lo - 1)) / 2 + (hi - lo)];
    }
    s_mass[t] = mv;
  }

  double* gK = Kvals + (size_t)b0 * 36;
  double* gM = Mdiag + (size_t)n0 * 36;
  if (BULK) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                   :: "l"(gK), "r"(smem_u32(s_out)), "r"((uint32_t)(nblk * 288)) : "memory");
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                   :: "l"(gM), "r"(smem_u32(s_mass)), "r"((uint32_t)(nn * 288)) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  } else {
    __syncthreads();
    const double2* s2 = reinterpret_cast<const double2*>(s_out);
    double2* g2 = reinterpret_cast<double2*>(gK);
    for (int i = tid; i < nblk * 18; i += THREADS) g2[i] = s2[i];
    const double2* m2 = reinterpret_cast<const double2*>(s_mass);
    double2* gm2 = reinterpret_cast<double2*>(gM);
    for (int i = tid; i < nn * 18; i += THREADS) gm2[i] = m2[i];
  }
}

// ---- parity export: per-element global-axis matrices ------------------------------------
__global__ void frame_elements_kernel(FrameParams P, int64_t n_elem, double* __restrict__ ke,
                                      double* __restrict__ me) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_elem * 4) return;
  const uint32_t e = (uint32_t)(t >> 2);
  const int a = (int)((t >> 1) & 1), b = (int)(t & 1);
  FrameRec R;
  frame_record(P, e, R);
  double acc[36];
  if (ke) {
    frame_kblock<true>(R, a, b, acc);
    for (int r = 0; r < 6; ++r)
      for (int c = 0; c < 6; ++c) ke[(size_t)e * 144 + (6 * a + r) * 12 + 6 * b + c] = acc[r * 6 + c];
  }
  if (me) {
    if (a == b) frame_mblock(R, acc);
    for (int r = 0; r < 6; ++r)
      for (int c = 0; c < 6; ++c)
        me[(size_t)e * 144 + (6 * a + r) * 12 + 6 * b + c] = (a == b) ? acc[r * 6 + c] : 0.0;
  }
}

__global__ void tet10_elements_kernel(Tet10Params P, int64_t n_elem, double* __restrict__ ke) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_elem * 100) return;
  const uint32_t e = (uint32_t)(t / 100);
  const int a = (int)((t % 100) / 10), b = (int)(t % 10);
  double acc[9];
  Tet10Params Q = P;
  tet10_kblock(Q, e, a, b, false, acc);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) ke[(size_t)e * 900 + (3 * a + r) * 30 + 3 * b + c] = acc[r * 3 + c];
}

static FrameParams frame_params(const femb_handle* h) {
  FrameParams P;
  P.xyz = h->xyz.p; P.conn = h->conn.p; P.elem_sec = h->elem_sec.p; P.sec_props = h->sec_props.p;
  P.E = h->E; P.G = h->G; P.rho = h->rho;
  return P;
}

static Tet10Params tet10_params(const femb_handle* h) {
  Tet10Params P;
  P.xyz = h->xyz.p; P.conn = h->conn.p;
  const double C1 = h->E / ((1.0 + h->nu) * (1.0 - 2.0 * h->nu));  // ReactionSolver.py:89
  const double C2 = (1.0 - 2.0 * h->nu) / 2.0;                      // :90
  P.lam = C1 * h->nu; P.mu2 = C1 * (1.0 - h->nu); P.gsh = C1 * C2;
  P.skipped = h->counters.p;
  return P;
}

int launch_frame_elements(femb_handle* h, double* d_ke, double* d_me) {
  const int64_t n = h->n_elem * 4;
  if (n == 0) return FEMB_OK;
  frame_elements_kernel<<<(unsigned)((n + 127) / 128), 128, 0, h->stream>>>(frame_params(h), h->n_elem, d_ke, d_me);
  h->launches++;
  FEMB_CUDA(h, cudaGetLastError());
  return FEMB_OK;
}

int launch_tet10_elements(femb_handle* h, double* d_ke) {
  const int64_t n = h->n_elem * 100;
  if (n == 0) return FEMB_OK;
  tet10_elements_kernel<<<(unsigned)((n + 127) / 128), 128, 0, h->stream>>>(tet10_params(h), h->n_elem, d_ke);
  h->launches++;
  FEMB_CUDA(h, cudaGetLastError());
  return FEMB_OK;
}

constexpr int kAsmThreads = 128;

static bool bulk_enabled() {
  const char* s = getenv("FEMB_ASM_BULK");
  return !(s && s[0] == '0');
}

template <class EL>
static int launch_assemble_t(femb_handle* h, const EL& el) {
  const femb::Symbolic& S = h->sym;
  const int n_tiles = (int)S.tile_ptr.size() - 1;
  if (n_tiles <= 0) return FEMB_OK;
  // dynamic shared memory: the largest tile
  size_t smem = 0;
  for (int t = 0; t < n_tiles; ++t) {
    const int n0 = S.tile_ptr[t], n1 = S.tile_ptr[t + 1];
    size_t need = (size_t)(S.rowptr[n1] - S.rowptr[n0]) * EL::BS2 * 8;
    if (EL::HAS_MASS) need += (size_t)(n1 - n0) * EL::BS2 * 8;
    smem = need > smem ? need : smem;
  }
  smem = (smem + 127) & ~size_t(127);
  if (smem > 200 * 1024) return fail(h, FEMB_ERR_ARG, "assembly tile exceeds shared memory (node degree too large)");
  PatternDev pat{h->rowptr.p, h->contrib_ptr.p, h->contrib.p, h->contrib_blk.p, h->tile_ptr.p};
  const bool bulk = bulk_enabled() && (EL::BS2 % 2 == 0);
  if (bulk) {
    auto k = assemble_tiles_kernel<EL, kAsmThreads, true>;
    FEMB_CUDA(h, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<n_tiles, kAsmThreads, smem, h->stream>>>(el, pat, h->Kvals.p, h->Mdiag.p);
  } else {
    auto k = assemble_tiles_kernel<EL, kAsmThreads, false>;
    FEMB_CUDA(h, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<n_tiles, kAsmThreads, smem, h->stream>>>(el, pat, h->Kvals.p, h->Mdiag.p);
  }
  h->launches++;
  FEMB_CUDA(h, cudaGetLastError());
  return FEMB_OK;
}

static int launch_assemble_pairs(femb_handle* h) {
  const femb::Symbolic& S = h->sym;
  const int n_tiles = (int)S.pair_tile_ptr.size() - 1;
  if (n_tiles <= 0) return FEMB_OK;
  size_t smem = 0;
  for (int t = 0; t < n_tiles; ++t) {
    const int n0 = S.pair_tile_ptr[t], n1 = S.pair_tile_ptr[t + 1];
    const size_t need = (size_t)(S.rowptr[n1] - S.rowptr[n0]) * 288 + (size_t)(n1 - n0) * (288 + 224);
    smem = need > smem ? need : smem;
  }
  smem = (smem + 127) & ~size_t(127);
  if (smem > 200 * 1024) return fail(h, FEMB_ERR_ARG, "assembly tile exceeds shared memory (node degree too large)");
  PairDev pat{h->rowptr.p, h->diag_blk.p, h->pair_ptr.p, h->pair_code.p, h->pair_blk.p, h->pair_rank.p, h->pair_tile_ptr.p};
  if (bulk_enabled()) {
    auto k = frame_assemble_pairs_kernel<kAsmThreads, true>;
    FEMB_CUDA(h, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<n_tiles, kAsmThreads, smem, h->stream>>>(frame_params(h), pat, h->Kvals.p, h->Mdiag.p);
  } else {
    auto k = frame_assemble_pairs_kernel<kAsmThreads, false>;
    FEMB_CUDA(h, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<n_tiles, kAsmThreads, smem, h->stream>>>(frame_params(h), pat, h->Kvals.p, h->Mdiag.p);
  }
  h->launches++;
  FEMB_CUDA(h, cudaGetLastError());
  return FEMB_OK;
}

static bool generic_forced() {
  const char* s = getenv("FEMB_ASM_GENERIC");
  return s && s[0] == '1';
}

int launch_assemble(femb_handle* h) {
  if (h->kind == Kind::Frame && h->sym.pairs_ok && !generic_forced()) return launch_assemble_pairs(h);
  if (h->kind == Kind::Frame) {
    FrameEl el;
    el.P = frame_params(h);
    return launch_assemble_t(h, el);
  }
  if (h->kind == Kind::Tet10) {
    FEMB_CUDA(h, cudaMemsetAsync(h->counters.p, 0, sizeof(unsigned long long) * h->counters.n, h->stream));
    Tet10El el;
    el.P = tet10_params(h);
    return launch_assemble_t(h, el);
  }
  return fail(h, FEMB_ERR_ARG, "no mesh set");
}

}  // namespace femb
