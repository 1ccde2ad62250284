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
#include <algorithm>

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
    if (P.grad) tet10_kblock_stored(P, e, ab / 10u, ab % 10u, acc);   // two-stage: point records from the pre-pass
    else tet10_kblock(P, e, ab / 10u, ab % 10u, ab == 0u, acc);
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
// Each thread reads ONE 16-byte pair record (own node, other node, target block, section /
// end flag / position in the node's list — everything the symbolic phase can resolve, so the
// only dependent loads left are the two coordinate triples and the section row), evaluates the
// element record once, and
//   * streams the off-diagonal block (node, other end) straight from registers to HBM with
//     nine 32-byte stores (STG.256; a block is 288 contiguous bytes, 32-byte aligned);
//   * parks its share of the node's diagonal block — the 21-value upper triangle plus 7 compact
//     lumped-mass values — in shared memory (odd row stride: conflict-free).
// After one barrier the CTA sums every node's shares in list (= element-ascending) order and
// writes the diagonal block and the 12 non-zero lumped-mass entries.  No float atomics, no rank
// loop: the summation order is fixed by the symbolic phase, so K and M are bit-reproducible.
// History (1M-DOF frame, one B200): generic tile kernel 330 us -> first pair kernel (tile staged
// in shared memory + bulk copy, rank loop) 209 us -> register-streamed stores + share table 127 us
// -> persistent CTAs with software-pipelined inputs 101 us (profiles/r01_ncu_full_assembly_*.txt).
__device__ __forceinline__ void st_global_256(double* p, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

constexpr int kShareStride = 29;  // odd stride (in doubles): conflict-free 64-bit shared accesses

// Persistent form of the pair kernel: one CTA per resident slot walks tiles t, t+G, t+2G, ...
// and software-pipelines the input side — the tile descriptor two tiles ahead, the pair records
// one tile ahead, and (after the first barrier, when the element record's registers are dead)
// the coordinates and section row of the next tile — so a tile's arithmetic never waits on a
// dependent chain of global loads.
struct PairDevP {
  const int4* rec;        // (n_pairs) {node, other, blk, sec | a<<24 | pos<<25}
  const int4* node_rec;   // (n_nodes) {first pair, pair count, diagonal block, 0}
  const int4* tiles;      // (n_tiles) {first node, node count, first pair, pair count}
  int n_tiles;
};

template <int THREADS>
__global__ void __launch_bounds__(THREADS, 4)
frame_assemble_pairs_persistent_kernel(const FrameParams P, const PairDevP pat, double* __restrict__ Kvals,
                                       double* __restrict__ Mdiag) {
  __shared__ double s_c[THREADS * kShareStride];
  __shared__ int4 s_node[THREADS];
  __shared__ uint8_t s_rc[28];
  const int tid = threadIdx.x;
  const int G = gridDim.x;
  int t = blockIdx.x;
  if (t >= pat.n_tiles) return;
  if (tid < 28) {
    int r = 0, c = 0;
    if (tid < 21) {
      int q = tid;
      while (q >= 6 - r) { q -= 6 - r; ++r; }
      c = r + q;
    } else if (tid > 21) {
      int q = tid - 22;
      while (q >= 3 - r) { q -= 3 - r; ++r; }
      c = 3 + r + q; r += 3;
    }
    s_rc[tid] = (uint8_t)(r * 6 + c);
  }
  const int4 zero4 = make_int4(0, 0, 0, 0);
  int4 td = __ldg(pat.tiles + t);
  int4 td1 = (t + G < pat.n_tiles) ? __ldg(pat.tiles + t + G) : zero4;
  int4 rec = (tid < td.w) ? __ldg(pat.rec + td.z + tid) : zero4;
  FrameIn in;
  {
    const int a = (rec.w >> 24) & 1;
    frame_load(P, a ? rec.y : rec.x, a ? rec.x : rec.y, rec.w & 0xFFFFFF, in);  // inactive lanes read node 0 / row 0
  }
  for (; t < pat.n_tiles; t += G) {
    // ---- issue the loads of the tiles ahead
    const int4 td2 = (t + 2 * G < pat.n_tiles) ? __ldg(pat.tiles + t + 2 * G) : zero4;
    const int4 rec1 = (tid < td1.w) ? __ldg(pat.rec + td1.z + tid) : zero4;
    const int4 nrec = (tid < td.y) ? __ldg(pat.node_rec + td.x + tid) : zero4;
    // ---- phase 1: element record, off-diagonal block to HBM, diagonal shares to shared memory
    if (tid < td.w) {
      const int a = (rec.w >> 24) & 1;
      FrameRec R;
      frame_record_from(P, in, R);
      double* g = Kvals + (size_t)rec.z * 36;
      double row[12];
#pragma unroll
      for (int h = 0; h < 3; ++h) {
        frame_offdiag_row(R, a, 2 * h, row);
        frame_offdiag_row(R, a, 2 * h + 1, row + 6);
        st_global_256(g + 12 * h, row[0], row[1], row[2], row[3]);
        st_global_256(g + 12 * h + 4, row[4], row[5], row[6], row[7]);
        st_global_256(g + 12 * h + 8, row[8], row[9], row[10], row[11]);
      }
      double d[28];
      frame_diag_sym(R, a, d, d + 21);
      double* mine = s_c + tid * kShareStride;
#pragma unroll
      for (int q = 0; q < 28; ++q) mine[q] = d[q];
    }
    s_node[tid] = nrec;
    __syncthreads();
    // ---- the next tile's coordinates / section rows travel while phase 2 runs
    {
      const int a = (rec1.w >> 24) & 1;
      frame_load(P, a ? rec1.y : rec1.x, a ? rec1.x : rec1.y, rec1.w & 0xFFFFFF, in);
    }
    // ---- phase 2: ordered per-node sums -> diagonal K block and lumped-mass entries
    for (int o = tid; o < td.y * 28; o += THREADS) {
      const int ns = o / 28, q = o - ns * 28;
      const int4 nr = s_node[ns];
      const int j0 = nr.x - td.z, j1 = j0 + nr.y;
      double v = 0.0;
      for (int j = j0; j < j1; ++j) v += s_c[j * kShareStride + q];
      const int rc = s_rc[q];
      const int r = rc / 6, c = rc - r * 6;
      if (q < 21) {
        double* kd = Kvals + (size_t)nr.z * 36;
        kd[rc] = v;
        if (r != c) kd[c * 6 + r] = v;
      } else {
        double* md = Mdiag + (size_t)(td.x + ns) * 36;
        if (q == 21) { md[0] = v; md[7] = v; md[14] = v; }
        else {
          md[rc] = v;
          if (r != c) md[c * 6 + r] = v;
        }
      }
    }
    __syncthreads();  // s_c / s_node are rewritten by the next tile
    td = td1; td1 = td2; rec = rec1;
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

// pre-pass of the two-stage Tet10 assembly: one thread per (element, Gauss point)
__global__ void tet10_point_records_kernel(Tet10Params P, int64_t n_elem, double* __restrict__ grad) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_elem * 4) return;
  tet10_point_record(P, (uint32_t)(t >> 2), (int)(t & 3), grad + (size_t)t * kTetGradStride);
}

// second stage: one thread per 3x3 block of K walks the block's contribution list (element-ascending,
// fixed by the symbolic phase) and adds the contributions in registers — no barriers, no atomics, one
// 72-byte store per block.  (The generic tile kernel spends its time in the rank loop: a Tet10 block
// collects up to ~24 contributions, one CTA barrier each.)
__global__ void __launch_bounds__(128)
tet10_block_gather_kernel(const Tet10Params P, const int32_t* __restrict__ contrib_ptr, const uint32_t* __restrict__ contrib,
                          int64_t nnzb, double* __restrict__ Kvals) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nnzb) return;
  double acc[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) acc[k] = 0.0;
  const int c0 = __ldg(contrib_ptr + b), c1 = __ldg(contrib_ptr + b + 1);
  for (int c = c0; c < c1; ++c) {
    const uint32_t code = __ldg(contrib + c);
    const uint32_t e = code / 100u, ab = code % 100u;
    double t[9];
    tet10_kblock_stored(P, e, (int)(ab / 10u), (int)(ab % 10u), t);
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] += t[k];
  }
  double* o = Kvals + (size_t)b * 9;
#pragma unroll
  for (int k = 0; k < 9; ++k) o[k] = acc[k];
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
  P.grad = nullptr;
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
  PairDevP pp{reinterpret_cast<const int4*>(h->pair_rec.p), reinterpret_cast<const int4*>(h->pair_node_rec.p),
              reinterpret_cast<const int4*>(h->pair_tiles.p), n_tiles};
  const int grid = std::min(n_tiles, h->num_sms * 4);   // 4 resident CTAs per SM (128 registers, 32 KB shared)
  frame_assemble_pairs_persistent_kernel<kAsmThreads><<<grid, kAsmThreads, 0, h->stream>>>(frame_params(h), pp, h->Kvals.p, h->Mdiag.p);
  h->launches++;
  FEMB_CUDA(h, cudaGetLastError());
  return FEMB_OK;
}

static bool generic_forced() {
  const char* s = getenv("FEMB_ASM_GENERIC");
  return s && s[0] == '1';
}

int launch_assemble(femb_handle* h) {
  if (h->kind == Kind::Frame && h->pairs_dev_ok && !generic_forced()) return launch_assemble_pairs(h);
  if (h->kind == Kind::Frame) {
    FrameEl el;
    el.P = frame_params(h);
    return launch_assemble_t(h, el);
  }
  if (h->kind == Kind::Tet10) {
    FEMB_CUDA(h, cudaMemsetAsync(h->counters.p, 0, sizeof(unsigned long long) * h->counters.n, h->stream));
    Tet10El el;
    el.P = tet10_params(h);
    static int one_stage = -1;   // FEMB_TET10_ONE_STAGE=1: every contribution thread evaluates the Jacobians itself
    if (one_stage < 0) { const char* e = getenv("FEMB_TET10_ONE_STAGE"); one_stage = (e && e[0] == '1') ? 1 : 0; }
    if (!one_stage && h->n_elem > 0) {
      FEMB_CUDA(h, h->tet_grad.ensure((size_t)h->n_elem * 4 * kTetGradStride));
      const int64_t nt = h->n_elem * 4;
      tet10_point_records_kernel<<<(unsigned)((nt + 127) / 128), 128, 0, h->stream>>>(el.P, h->n_elem, h->tet_grad.p);
      h->launches++;
      FEMB_CUDA(h, cudaGetLastError());
      el.P.grad = h->tet_grad.p;
      const int64_t nnzb = h->sym.nnzb;
      tet10_block_gather_kernel<<<(unsigned)((nnzb + 127) / 128), 128, 0, h->stream>>>(el.P, h->contrib_ptr.p, h->contrib.p, nnzb, h->Kvals.p);
      h->launches++;
      FEMB_CUDA(h, cudaGetLastError());
      return FEMB_OK;
    }
    return launch_assemble_t(h, el);
  }
  return fail(h, FEMB_ERR_ARG, "no mesh set");
}

}  // namespace femb
