// Device helpers shared by the single-GPU PCG (solver.cu) and the row-block distributed PCG
// (dist.cu): scalar / flag slots, ordered reductions, block-Jacobi row application.
#pragma once

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

// Last-CTA-done ordered reduction of NV interleaved partial arrays (partials[v*stride + cta]).
// Returns true in ALL threads of the last CTA, whose thread 0 holds the totals in out[].
template <int THREADS, int NV>
__device__ __forceinline__ bool grid_reduce(const double (&mine)[NV], double* partials, int stride,
                                            int* ticket, double* s_red, double (&out)[NV]) {
  __shared__ int s_last;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int v = 0; v < NV; ++v) partials[(size_t)v * stride + blockIdx.x] = mine[v];
    __threadfence();
    const int t = atomicAdd(ticket, 1);
    s_last = (t == (int)gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return false;
  __threadfence();
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    double a = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += THREADS) a += partials[(size_t)v * stride + i];
    out[v] = block_sum<THREADS>(a, s_red);
  }
  if (threadIdx.x == 0) *ticket = 0;
  return true;
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

inline int vec_grid(const femb_handle* h, int64_t n, int threads) {
  int64_t need = (n + threads - 1) / threads;
  int64_t cap = (int64_t)h->num_sms * 8;
  return (int)(need < cap ? (need > 0 ? need : 1) : cap);
}

// grid for kernels whose CTAs must own whole nodes: chunk = THREADS rows must be a multiple
// of BS in the grid-stride pattern.  THREADS=256 is not a multiple of 6, so those kernels
// use kRowThreads = 192 (divisible by 6 and 3).
constexpr int kRowThreads = 192;


}  // namespace femb
