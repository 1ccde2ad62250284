// BSR-6 SpMV with the matrix streamed by the TMA engine (cp.async.bulk global -> shared with
// mbarrier completion): persistent CTAs walk tiles of consecutive block rows; the values of a
// tile are ONE contiguous range of HBM, so a single elected thread keeps a ring of ST bulk copies
// in flight (ST-1 tiles ahead of the arithmetic) while the CTA's threads — one per scalar row —
// read their 48-byte row pieces from shared memory and gather x through the read-only path.
// The LSU no longer carries the 359 MB matrix stream (no per-thread global loads, no sector
// re-requests); only the x gathers and the small index arrays go through L1.
// Same masking / fused-dot contract as bsr_spmv_kernel in solver.cu.
#include "common.cuh"
#include "pcg_common.cuh"

namespace femb {

namespace {

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
// bounded wait: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_addr(bar);
  for (int spin = 0; spin < (1 << 22); ++spin) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    if (ok) return;
  }
  __trap();
}

}  // namespace

template <int TN, int ST, bool MASKED, bool DOT>
__global__ void __launch_bounds__(TN * 6)
bsr6_spmv_tma_kernel(const int32_t* __restrict__ tile_ptr, int n_tiles, const int32_t* __restrict__ rowptr,
                     const int32_t* __restrict__ colidx, const double* __restrict__ vals,
                     const uint8_t* __restrict__ free_mask, const double* __restrict__ x, double* __restrict__ y,
                     int stage_bytes, double* partials, int pstride, double* scal, int* flags) {
  constexpr int THREADS = TN * 6;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)ST * stage_bytes);
  if (DOT && flags[Flag::DONE]) return;
  const int tid = threadIdx.x;
  const int G = gridDim.x;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < ST; ++s) mbar_init(bars + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int t, int s) {
    const int n0 = __ldg(tile_ptr + t), n1 = __ldg(tile_ptr + t + 1);
    const int b0 = __ldg(rowptr + n0), b1 = __ldg(rowptr + n1);
    const uint32_t bytes = (uint32_t)(b1 - b0) * 288u;
    mbar_expect_tx(bars + s, bytes);
    bulk_g2s(smem + (size_t)s * stage_bytes, vals + (size_t)b0 * 36, bytes, bars + s);
  };
  if (tid == 0) {
#pragma unroll
    for (int k = 0; k < ST; ++k) {
      const int t = blockIdx.x + k * G;
      if (t < n_tiles) issue(t, k);
    }
  }
  const int ln = tid / 6, r = tid - ln * 6;
  double dot = 0.0;
  int i = 0;
  for (int t = blockIdx.x; t < n_tiles; t += G, ++i) {
    const int s = i % ST;
    const uint32_t phase = (uint32_t)((i / ST) & 1);
    const int n0 = __ldg(tile_ptr + t), n1 = __ldg(tile_ptr + t + 1);
    const int tb0 = __ldg(rowptr + n0);
    const int node = n0 + ln;
    const bool active = node < n1;
    int b0 = 0, b1 = 0;
    if (active) { b0 = __ldg(rowptr + node); b1 = __ldg(rowptr + node + 1); }
    mbar_wait(bars + s, phase);
    if (active) {
      const double* sv = reinterpret_cast<const double*>(smem + (size_t)s * stage_bytes) + (size_t)(b0 - tb0) * 36 + r * 6;
      double acc = 0.0;
#pragma unroll 8
      for (int b = b0; b < b1; ++b) {
        const int col = __ldg(colidx + b);
        const double2* a2 = reinterpret_cast<const double2*>(sv + (size_t)(b - b0) * 36);
        const double2* x2 = reinterpret_cast<const double2*>(x + (size_t)col * 6);
        const double2 a0 = a2[0], a1 = a2[1], a2v = a2[2];
        const double2 x0 = __ldg(x2), x1 = __ldg(x2 + 1), x2v = __ldg(x2 + 2);
        acc += a0.x * x0.x; acc += a0.y * x0.y; acc += a1.x * x1.x;
        acc += a1.y * x1.y; acc += a2v.x * x2v.x; acc += a2v.y * x2v.y;
      }
      const int64_t g = (int64_t)node * 6 + r;
      const double xg = x[g];
      if (MASKED && !free_mask[g]) acc = xg;
      y[g] = acc;
      if (DOT) dot += xg * acc;
    }
    __syncthreads();                       // every row of the stage has been consumed
    if (tid == 0) {
      const int tn = t + ST * G;
      if (tn < n_tiles) issue(tn, s);
    }
  }
  if (DOT) {
    double mine[1], tot[1];
    mine[0] = dot;
    if (grid_reduce<THREADS, 1>(mine, partials, pstride, flags + Flag::TICKET0, tot)) {
      if (threadIdx.x == 0) scal[Scal::PQ] = tot[0];
    }
  }
}

// host: tiles of <= TN consecutive nodes whose blocks fit one stage
int build_spmv_tiles(femb_handle* h, int tn, int stage_bytes_cap) {
  const Symbolic& S = h->sym;
  std::vector<int32_t> tp;
  tp.push_back(0);
  int64_t max_bytes = 0;
  int32_t i0 = 0;
  for (int64_t i = 0; i < h->n_nodes; ++i) {
    const int64_t bytes = (int64_t)(S.rowptr[i + 1] - S.rowptr[i0]) * 288;
    if (i + 1 - i0 > tn || bytes > stage_bytes_cap) {
      if (i == i0) return -1;              // a single block row exceeds the stage: caller falls back
      tp.push_back((int32_t)i);
      i0 = (int32_t)i;
    }
    max_bytes = std::max<int64_t>(max_bytes, (int64_t)(S.rowptr[i + 1] - S.rowptr[i0]) * 288);
  }
  tp.push_back((int32_t)h->n_nodes);
  h->spmv_tile_nodes = tn;
  h->spmv_stage_bytes = (int)((max_bytes + 127) & ~int64_t(127));
  h->spmv_n_tiles = (int)tp.size() - 1;
  cudaError_t e = upload(h->spmv_tiles, tp, h->stream);
  if (e != cudaSuccess) return -1;
  cudaStreamSynchronize(h->stream);
  return 0;
}

template <int TN, int ST>
static int launch_t(femb_handle* h, const double* x, double* y, bool masked, double* dot_partials, double* scal_out) {
  const size_t smem = (size_t)ST * h->spmv_stage_bytes + ST * sizeof(uint64_t) + 64;
  const int pstride = h->num_sms * 8;
  int ctas_per_sm = (int)((227 * 1024) / (smem + 1024));
  if (ctas_per_sm < 1) return fail(h, FEMB_ERR_ARG, "TMA SpMV stage does not fit shared memory");
  ctas_per_sm = std::min(ctas_per_sm, 2048 / (TN * 6));
  const int grid = std::min(h->spmv_n_tiles, h->num_sms * ctas_per_sm);
  if (grid > pstride) return fail(h, FEMB_ERR_ARG, "TMA SpMV grid exceeds the reduction scratch");
#define GO(M, D)                                                                                         \
  do {                                                                                                   \
    auto k = bsr6_spmv_tma_kernel<TN, ST, M, D>;                                                         \
    FEMB_CUDA(h, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
    k<<<grid, TN * 6, smem, h->stream>>>(h->spmv_tiles.p, h->spmv_n_tiles, h->rowptr.p, h->colidx.p,     \
                                         h->Kvals.p, h->free_mask.p, x, y, h->spmv_stage_bytes,          \
                                         dot_partials, pstride, scal_out, h->flags.p);                   \
  } while (0)
  const bool dot = dot_partials != nullptr;
  if (masked && dot) GO(true, true);
  else if (masked) GO(true, false);
  else GO(false, false);
#undef GO
  h->launches++;
  FEMB_CUDA(h, cudaGetLastError());
  return FEMB_OK;
}

// variant: 1 = 32-node tiles, 3 stages (1 CTA/SM); 2 = 16-node tiles, 2 stages (3 CTAs/SM);
//          4 = 16-node tiles, 3 stages (2 CTAs/SM)
// Measured at 1M DOF (B200, inside PCG): 133 / 99 / 134 us against 85 us for the plain kernel on the
// same box — with a CTA-wide barrier per tile the x-gather latency of a tile is exposed and only a
// few tiles per SM are in the arithmetic phase at once; the path stays opt-in (FEMB_SPMV_TMA=1|2|4).
int launch_spmv_tma(femb_handle* h, int variant, const double* x, double* y, bool masked, double* dot_partials,
                    double* scal_out) {
  const int tn = variant == 1 ? 32 : 16;
  if (h->spmv_tile_nodes != tn || !h->spmv_tiles.p) {
    if (build_spmv_tiles(h, tn, 72 * 1024) != 0) return fail(h, FEMB_ERR_ARG, "TMA SpMV: block rows too long for a stage");
  }
  switch (variant) {
    case 1: return launch_t<32, 3>(h, x, y, masked, dot_partials, scal_out);
    case 2: return launch_t<16, 2>(h, x, y, masked, dot_partials, scal_out);
    default: return launch_t<16, 3>(h, x, y, masked, dot_partials, scal_out);
  }
}

}  // namespace femb
