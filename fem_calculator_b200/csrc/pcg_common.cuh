// Device helpers shared by the single-GPU PCG (solver.cu) and the row-block distributed PCG
// (dist.cu): scalar / flag slots, ordered reductions, block-Jacobi row application.
#pragma once

#include <cstdlib>
#include <utility>

#include "common.cuh"

namespace femb {

constexpr int kVecThreads = 256;

struct Scal {  // device scalar block (doubles)
  enum { PQ = 0, RZ0 = 1, RZ1 = 2, RR = 3, BB = 4, TOL2 = 5, ALPHA = 6, COUNT = 8 };
};
struct Flag {  // device int block
  enum { DONE = 0, ITERS = 1, TICKET0 = 2, TICKET1 = 3, TICKET2 = 4, COUNT = 8 };
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// CTA-wide ordered sum; result valid in thread 0.
template <int THREADS>
__device__ __forceinline__ double block_sum(double v, double* s_red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) s_red[w] = v;
  __syncthreads();
  double t = 0.0;
  if (w == 0) {
    t = (l < THREADS / 32) ? s_red[l] : 0.0;
    t = warp_sum(t);
  }
  __syncthreads();
  return t;
}

// Ordered grid-wide sum of NV per-thread values (partials[v*stride + cta]) in two stages:
//   1. one CTA-wide pass: warp shuffles, one barrier, thread v adds the warp sums of value v in
//      warp order and publishes the CTA's partial;
//   2. the last CTA to take a ticket sums the partial arrays, warp w handling values w, w+NW, ..
//      (lanes stride the CTAs, then a shuffle tree) — all NV values in parallel, no float atomics.
// The association is fixed by the launch geometry, so results are bit-reproducible run to run.
// Returns true in ALL threads of the last CTA, with the totals in out[] (every thread).
template <int THREADS, int NV>
__device__ __forceinline__ bool grid_reduce(const double (&mine)[NV], double* partials, int stride,
                                            int* ticket, double (&out)[NV]) {
  constexpr int NW = THREADS / 32;
  static_assert(NV <= THREADS, "one thread per value");
  __shared__ double s_part[NV * NW];
  __shared__ double s_tot[NV];
  __shared__ int s_last;
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const double t = warp_sum(mine[v]);
    if (l == 0) s_part[v * NW + w] = t;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < NW; ++k) t += s_part[threadIdx.x * NW + k];
    partials[(size_t)threadIdx.x * stride + blockIdx.x] = t;
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int t = atomicAdd(ticket, 1);
    s_last = (t == (int)gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
  for (int v = w; v < NV; v += NW) {
    double a = 0.0;
    for (int i = l; i < (int)gridDim.x; i += 32) a += __ldcg(partials + (size_t)v * stride + i);
    a = warp_sum(a);
    if (l == 0) s_tot[v] = a;
  }
  __syncthreads();
#pragma unroll
  for (int v = 0; v < NV; ++v) out[v] = s_tot[v];
  if (threadIdx.x == 0) *ticket = 0;
  return true;
}

// CTA-wide ordered sum of NV per-thread values; the result is returned in every thread.
template <int THREADS, int NV>
__device__ __forceinline__ void block_sum_all(double (&v)[NV], double* s_part /* NV * THREADS / 32 */) {
  constexpr int NW = THREADS / 32;
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const double t = warp_sum(v[k]);
    if (l == 0) s_part[k * NW + w] = t;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < NW; ++i) t += s_part[k * NW + i];
    v[k] = t;
  }
  __syncthreads();
}

// ---- linked reductions of the single-vector PCG ---------------------------------------------
// grid_reduce finishes a dot product with a serial tail — ticket atomic, last CTA re-reads all
// partials, writes the scalar, the next kernel reads it — measured at 3.7 us per kernel on a 21 us
// operator launch.  In the linked form a kernel only PUBLISHES its per-CTA partial sums; every CTA
// of the NEXT kernel adds all published partials itself, in the same fixed order (bit-identical in
// all CTAs, still no float atomics), and derives alpha / beta / the convergence decision locally.
//   operator(it): consumes {gamma, rr} of update(it-1) [buffer it & 1] -> convergence / iteration cap,
//                 publishes delta = (z, s)
//   update(it)  : consumes delta and {gamma, rr} [buffer it & 1] -> beta, alpha; publishes the new
//                 {gamma, rr} into buffer (it+1) & 1
// CTA 0 mirrors the scalars into scal / flags for the host; gamma and alpha of the previous iteration
// travel through parity-indexed slots so no kernel reads a slot another CTA of the same launch writes.
struct PcgLink {
  double* upd_partials;   // [2 buffers][2 values][pstride]   {gamma = (r, z), rr = (r, r)}
  double* op_partials;    // [pstride]                         {delta = (z, A z)}
  double* scal;
  int* flags;
  int n_upd, n_op;        // CTAs that publish into upd_partials / op_partials
  int pstride;
  int it, max_iter;
  double rtol;
};

// operator-side prologue: true (in all threads of all CTAs alike) when the solve is over
template <int THREADS>
__device__ __forceinline__ bool pcg_link_decide(const PcgLink& L, double* s_part /* 2 * THREADS / 32 */) {
  const double* pb = L.upd_partials + (size_t)(L.it & 1) * 2 * L.pstride;
  double tot[2] = {0.0, 0.0};
  for (int i = threadIdx.x; i < L.n_upd; i += THREADS) { tot[0] += __ldcg(pb + i); tot[1] += __ldcg(pb + L.pstride + i); }
  block_sum_all<THREADS, 2>(tot, s_part);
  const double rr = tot[1];
  int done = 0;
  double tol2;
  if (L.it == 0) {
    tol2 = L.rtol * L.rtol * rr;             // the init kernel publishes {gamma_0, ||b||^2}
    if (rr == 0.0) done = 1;                 // zero load: u = 0 is the answer
  } else {
    tol2 = L.scal[Scal::TOL2];
    if (rr <= tol2) done = 1;
    else if (L.it >= L.max_iter) done = 3;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    L.scal[Scal::RR] = rr;
    if (L.it == 0) { L.scal[Scal::BB] = rr; L.scal[Scal::TOL2] = tol2; }
    L.flags[Flag::ITERS] = L.it;             // updates completed so far
    if (done) L.flags[Flag::DONE] = done;
  }
  return done != 0;
}

template <int BS>
__device__ __forceinline__ double apply_dinv_row(const double* __restrict__ Dinv, int64_t g, const double* rn) {
  const int64_t node = g / BS;
  const int r = (int)(g - node * BS);
  const double* d = Dinv + (size_t)node * BS * BS + r * BS;
  double z = 0.0;
#pragma unroll
  for (int c = 0; c < BS; ++c) z += __ldg(d + c) * rn[c];
  return z;
}


// ---- peer-memory exchange (dist.cu; hooks in the SpMV and the distributed update kernel) -----
struct MailSlot { double v[4]; long long seq; long long pad[3]; };   // 64 bytes
constexpr int kMaxRanks = 8;
constexpr size_t kP2PZOffset = 8192;   // mailboxes + flags live below, the z vector above (same layout on every rank)

struct P2PDev {
  double* peer_z[kMaxRanks];          // per neighbour k: its z vector (mapped)
  long long peer_ghost_start[kMaxRanks];   // per neighbour k: first local node of MY data in its numbering
  long long* peer_halo_flag[kMaxRanks];    // per neighbour k: its halo flag array (mapped), indexed by source rank
  MailSlot* peer_mail[kMaxRanks];     // per RANK p: its mailbox array [world][2]
  MailSlot* my_mail;
  long long* my_halo_flag;            // [world]
  const long long* base;              // device copy of the per-solve sequence base
  int nbr[kMaxRanks];
  long long send_ptr[kMaxRanks + 1];
  int n_nbr, world, rank;
  // fused mode (no separate exchange kernels): per owned node, destination of its z entries
  // (k << 28 | ghost node index at neighbour k, -1 = interior) and the few nodes with several
  // destinations as (node, slot) pairs handled by the update kernel's last CTA
  const int32_t* send_slot;
  const int32_t* extra;      // (n_extra, 2)
  int n_extra;
  // ---- persistent line-preconditioned PCG (lines.cu): flag-in-data exchanges -------------------------------------
  // Every value travels as ONE 16-byte store {low word, flag, high word, flag} (flag = low 32 bits of the exchange's
  // sequence number): the receiver polls the slot itself until both flags match, so no system-scope fence, no separate
  // flag store and no second round trip sit between "value produced" and "value usable" (the fence + release-flag form
  // of the other kernels costs 10-20 us per exchange inside a 90 us iteration; this form 2-3).
  uint4* peer_ll_scal[kMaxRanks];     // per RANK p: its scalar slots   [source rank][parity][4]
  uint4* my_ll_scal;
  uint4* peer_ll_rb[kMaxRanks];       // per RANK p: its coarse-residual slots [source rank][parity][kLnMaxCoarse]
  uint4* my_ll_rb;
  uint4* peer_ll_halo[kMaxRanks];     // per neighbour k: the slots of MY nodes in its ghost tail (6 per node)
  uint4* my_ll_halo;                  // [ghost node - n_owned][6]
  long long recv_start[kMaxRanks];    // per neighbour k: first local (ghost) node it fills, and how many
  long long recv_count[kMaxRanks];
  long long n_owned;
  // owned nodes with at least one remote destination, ascending (the prolongation handles them first so that their
  // values are on the wire while the interior runs), and their destinations (k << 28 | node offset in peer_ll_halo[k])
  const int32_t* bnd_nodes;
  const int32_t* bnd_dst_ptr;         // (n_bnd + 1)
  const int32_t* bnd_dst;
  int n_bnd;
};
constexpr size_t kLLScalBytes = (size_t)kMaxRanks * 2 * 4 * sizeof(uint4);   // scalar slots of the flag-in-data exchange
constexpr int kLnMaxCoarse = 3072;    // capacity of the coarse-residual mail slots (3 families x 1024 bundles)

__device__ __forceinline__ void st_release_sys(long long* p, long long v) {
  asm volatile("st.release.sys.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ long long ld_acquire_sys(const long long* p) {
  long long v;
  asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
constexpr long long kSpinLimit = 1ll << 24;   // several seconds of polling, then DONE = 4

// flag-in-data slot (see P2PDev): 8-byte halves {data word, flag} are written / read atomically
__device__ __forceinline__ void ll_store(uint4* slot, double v, unsigned flag) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(slot), "r"((unsigned)b), "r"(flag),
               "r"((unsigned)(b >> 32)), "r"(flag) : "memory");
}
// polls until the slot carries `flag`; false after kSpinLimit polls (a lost peer)
__device__ __forceinline__ bool ll_wait(const uint4* slot, unsigned flag, double& v) {
  for (long long spins = 0; spins < kSpinLimit; ++spins) {
    unsigned a, f1, c, f2;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(f1), "=r"(c), "=r"(f2) : "l"(slot) : "memory");
    if (f1 == flag && f2 == flag) {
      v = __longlong_as_double((long long)(((unsigned long long)c << 32) | a));
      return true;
    }
    // tens of thousands of threads poll at once: back off so that the polls do not keep the L2 busy while the peers'
    // stores are trying to get in
    __nanosleep(spins < 64 ? 40 : 200);
  }
  v = 0.0;
  return false;
}

// red[] slots of the distributed solver (doubles): the first three are all-reduced each iteration
struct Red {
  enum { DELTA = 0, GAMMA = 1, RR = 2, DELTA2 = 3, NRED = 4, GPREV = 5, ALPHA = 6, TOL2 = 7, BB = 8, RRFINAL = 9, COUNT = 12 };
};



inline int vec_grid(const femb_handle* h, int64_t n, int threads) {
  int64_t need = (n + threads - 1) / threads;
  int64_t cap = (int64_t)h->num_sms * 8;
  return (int)(need < cap ? (need > 0 ? need : 1) : cap);
}

// Grid of a grid-stride kernel sized to what is actually co-resident: a kernel that needs more
// registers than 65536 / (8 CTAs x threads) would otherwise run its "8 CTAs per SM" in two waves,
// the second one mostly empty (measured: pcg_update_kernel, 56 registers -> 6 resident CTAs of 192
// threads, 46 % warps active with a 1184-CTA grid).  The resident count per kernel is cached.
template <class K>
inline int occ_grid(const femb_handle* h, K kernel, int64_t n, int threads) {
  static std::vector<std::pair<const void*, int>> cache;
  int per_sm = 0;
  for (const auto& e : cache)
    if (e.first == (const void*)kernel) per_sm = e.second;
  if (!per_sm) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;       // partials arrays hold num_sms * 8 entries
    cache.emplace_back((const void*)kernel, per_sm);
  }
  const int64_t need = (n + threads - 1) / threads;
  const int64_t cap = (int64_t)h->num_sms * per_sm;
  return (int)(need < cap ? (need > 0 ? need : 1) : cap);
}

// grid for kernels whose CTAs must own whole nodes: chunk = THREADS rows must be a multiple
// of BS in the grid-stride pattern.  THREADS=256 is not a multiple of 6, so those kernels
// use kRowThreads = 192 (divisible by 6 and 3).
constexpr int kRowThreads = 192;


}  // namespace femb
