// Row-block (node-slab) distributed PCG for one large frame / Tet10 mesh across GPUs
// (BASELINE configs 3/5): one process per GPU, each owning a contiguous range of
// the mesh's nodes.  A rank's LOCAL mesh = its owned nodes (numbered first, in global order)
// + the ghost nodes its elements touch (numbered after, sorted by global id, hence grouped by
// owner) + every element touching an owned node ("owner computes": cut elements are duplicated,
// so assembly needs no communication and the owned rows of K are bit-identical to the
// single-GPU matrix).  Only the owned prefix of rows is ever computed.
//
// Per CG iteration (Chronopoulos-Gear form, one reduction):
//   pack boundary entries of z -> ncclSend/ncclRecv with the neighbour ranks (ghost entries land
//   directly in the tail of z) -> s = A z on owned rows with the local (z, s) -> ONE ncclAllReduce
//   of {delta, gamma, ||r||^2} (3 doubles) -> fused vector update with the local partials of the
//   next {gamma, ||r||^2}.
// Everything is enqueued on the handle's stream; the host polls the device flag every
// `check_every` iterations.  NCCL is loaded with dlopen at femb_dist_init, so single-GPU users
// need no NCCL at all.
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "common.cuh"
#include "pcg_common.cuh"

namespace femb {

int bc_build_mask(femb_handle* h, const int64_t* d_fixed, int64_t n_fixed);
int setup_rhs_for_direct(femb_handle* h);
int setup_precond_public(femb_handle* h, int mode);

namespace {

// The few NCCL declarations this file needs, spelled out here so that the library builds without the NCCL
// development header (NCCL is only dlopen'ed, and only by multi-GPU callers); values follow nccl.h's stable ABI.
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclSum = 0 } ncclRedOp_t;
typedef enum { ncclDouble = 8 } ncclDataType_t;
struct ncclConfig_t;

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, ncclConfig_t*) = nullptr;   // optional (NCCL >= 2.18)
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
};

NcclApi g_nccl;

const char* load_nccl() {
  if (g_nccl.lib) return nullptr;
  // libnccl.so.2 resolves to the copy already loaded in the process (e.g. PyTorch's) or the system one
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) return "libnccl.so.2 not found (needed only for multi-GPU row-block solves)";
#define SYM(name)                                                           \
  g_nccl.name = reinterpret_cast<decltype(g_nccl.name)>(dlsym(lib, "nccl" #name)); \
  if (!g_nccl.name) return "libnccl.so.2 lacks nccl" #name
  SYM(GetUniqueId); SYM(CommInitRank); SYM(CommDestroy); SYM(AllReduce); SYM(Send); SYM(Recv);
  SYM(GroupStart); SYM(GroupEnd); SYM(GetErrorString); SYM(GetVersion);
#undef SYM
  // only the opt-in overlapped halo stream needs a second communicator: older NCCLs do without it
  g_nccl.CommSplit = reinterpret_cast<decltype(g_nccl.CommSplit)>(dlsym(lib, "ncclCommSplit"));
  g_nccl.lib = lib;
  return nullptr;
}

#define FEMB_NCCL(h, expr)                                                                   \
  do {                                                                                       \
    ncclResult_t _r = (expr);                                                                \
    if (_r != ncclSuccess)                                                                   \
      return femb::fail((h), FEMB_ERR_CUDA, std::string(#expr) + ": " + g_nccl.GetErrorString(_r)); \
  } while (0)

__global__ void pack_nodes_kernel(const double* __restrict__ v, const int32_t* __restrict__ nodes, int64_t n_send,
                                  int bs, double* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_send * bs) return;
  const int64_t i = t / bs;
  const int c = (int)(t - i * bs);
  out[t] = v[(size_t)nodes[i] * bs + c];
}

// owned rows: x = 0, r = b, z = Dinv r, p = q = 0; local gamma -> red[GAMMA], local ||b||^2 -> red[RR]
template <int BS, int THREADS, bool BLOCKJ>
__global__ void __launch_bounds__(THREADS)
dist_init_kernel(const double* __restrict__ b, const double* __restrict__ Dinv, double* __restrict__ x,
                 double* __restrict__ r, double* __restrict__ z, double* __restrict__ p, double* __restrict__ q,
                 int64_t n, double* partials, int pstride, double* red, int* flags) {
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
      red[Red::GAMMA] = tot[0];
      red[Red::RR] = tot[1];
      red[Red::ALPHA] = 1.0; red[Red::GPREV] = 1.0;
      flags[Flag::ITERS] = 0;
      flags[Flag::DONE] = 0;
    }
  }
}

// red[DELTA..RR] hold GLOBAL sums when this kernel starts (all-reduced on the stream before it)
template <int BS, int THREADS, bool BLOCKJ>
__global__ void __launch_bounds__(THREADS)
dist_update_kernel(const double* __restrict__ Dinv, const double* __restrict__ s, double* __restrict__ p,
                   double* __restrict__ q, double* __restrict__ x, double* __restrict__ r, double* __restrict__ z,
                   int64_t n, int first, double rtol, double* partials, int pstride, double* red, int* flags,
                   const P2PDev* __restrict__ p2p) {
  __shared__ double s_glob[3];
  if (flags[Flag::DONE]) return;
  double delta, gamma, rr;
  if (p2p) {
    // fused peer-memory mode: every rank's SpMV posted {delta, gamma, ||r||^2} in this rank's mailbox;
    // one warp per CTA waits for all of them and adds them in rank order (identical on every rank)
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
      for (int p = 0; p < p2p->world; ++p) {
        a0 += __shfl_sync(0xffffffffu, v0, p); a1 += __shfl_sync(0xffffffffu, v1, p); a2 += __shfl_sync(0xffffffffu, v2, p);
      }
      if (lane == 0) { s_glob[0] = a0; s_glob[1] = a1; s_glob[2] = lost ? -1.0 : a2; if (lost) flags[Flag::DONE] = 4; }
    }
    __syncthreads();
    delta = s_glob[0]; gamma = s_glob[1]; rr = s_glob[2];
    if (rr < 0.0) return;                 // a peer never answered: DONE = 4 is reported by the host
  } else {
    // (z, s) arrives in two parts: rows that read no ghost column, and the boundary rows
    delta = red[Red::DELTA] + red[Red::DELTA2]; gamma = red[Red::GAMMA]; rr = red[Red::RR];
  }
  const double tol2 = first ? rtol * rtol * rr : red[Red::TOL2];
  if (first ? (rr == 0.0) : (rr <= tol2)) {       // uniform across the grid: every CTA leaves
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
  double rz = 0.0, rr_new = 0.0;
  bool pushed = false;
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
      if (p2p) {
        const int sl = __ldg(p2p->send_slot + (int)(g / BS));
        if (sl >= 0) { p2p->peer_z[sl >> 28][(size_t)(sl & 0xFFFFFFF) * BS + rloc] = zg; pushed = true; }
      }
      rz += rg * zg; rr_new += rg * rg;
    }
    __syncthreads();  // q / r are read by the other rows of their node (same CTA, same step)
    for (int64_t g = (int64_t)blockIdx.x * THREADS + threadIdx.x; g < n; g += (int64_t)gridDim.x * THREADS) {
      const double qg = s[g] + beta * q[g];
      q[g] = qg;
      r[g] -= alpha * qg;
    }
  } else {
    for (int64_t g = (int64_t)blockIdx.x * THREADS + threadIdx.x; g < n; g += (int64_t)gridDim.x * THREADS) {
      const double pg = z[g] + beta * p[g];
      const double qg = s[g] + beta * q[g];
      const double rg = r[g] - alpha * qg;
      const double zg = __ldg(Dinv + g) * rg;
      p[g] = pg; q[g] = qg;
      x[g] += alpha * pg;
      r[g] = rg;
      z[g] = zg;
      if (p2p) {
        const int node = (int)(g / BS);
        const int sl = __ldg(p2p->send_slot + node);
        if (sl >= 0) { p2p->peer_z[sl >> 28][(size_t)(sl & 0xFFFFFFF) * BS + (int)(g - (int64_t)node * BS)] = zg; pushed = true; }
      }
      rz += rg * zg; rr_new += rg * rg;
    }
  }
  if (pushed) __threadfence_system();     // this thread's remote stores are ordered before the CTA's ticket
  double mine[2], tot[2];
  mine[0] = rz;
  mine[1] = rr_new;
  if (grid_reduce<THREADS, 2>(mine, partials, pstride, flags + Flag::TICKET1, tot)) {
    if (threadIdx.x == 0) {
      red[Red::GPREV] = gamma;
      red[Red::ALPHA] = alpha;
      if (first) { red[Red::TOL2] = tol2; red[Red::BB] = rr; }
      red[Red::RRFINAL] = rr;
      red[Red::GAMMA] = tot[0];     // local partials: all-reduced before the next update
      red[Red::RR] = tot[1];
      flags[Flag::ITERS] = flags[Flag::ITERS] + 1;
      if (bad) flags[Flag::DONE] = 2;
    }
    if (p2p) {
      // last CTA: every CTA's remote stores are done.  Nodes with several destinations, then the
      // sequence flag of the NEXT SpMV's halo at every neighbour.
      for (int e = threadIdx.x; e < p2p->n_extra * BS; e += THREADS) {
        const int i = e / BS, c = e - i * BS;
        const int node = p2p->extra[2 * i], sl = p2p->extra[2 * i + 1];
        p2p->peer_z[sl >> 28][(size_t)(sl & 0xFFFFFFF) * BS + c] = z[(size_t)node * BS + c];
      }
      __threadfence_system();
      __syncthreads();
      if ((int)threadIdx.x < p2p->n_nbr) {
        const long long seq = p2p->base[0] + flags[Flag::ITERS] + 1;   // ITERS was just advanced
        st_release_sys(p2p->peer_halo_flag[threadIdx.x] + p2p->rank, seq);
      }
    }
  }
}

// ---- peer-memory (NVLink) exchange: no NCCL call inside the iteration ----------------------
// Every rank maps its peers' z vectors and mailboxes through CUDA IPC.  Two small kernels per
// iteration replace the ncclSend/ncclRecv group and the ncclAllReduce:
//   p2p_halo_kernel      stores my boundary entries of z straight into each neighbour's ghost
//                        tail (remote stores over NVLink), then the last CTA releases a
//                        sequence flag at every neighbour and waits for theirs;
//   p2p_allreduce_kernel one warp: lane p stores my four partial sums + sequence number into
//                        peer p's mailbox, waits for peer p's entry in mine, and lane 0 adds the
//                        entries in rank order (deterministic) into red[0..3].
// Sequence numbers come from a per-solve base plus the device-side iteration counter, so the
// kernels have fixed arguments and replay inside a CUDA graph.  Waits are bounded: a lost peer
// sets DONE = 4 instead of hanging the GPU.
__global__ void __launch_bounds__(256)
p2p_halo_kernel(const double* __restrict__ z, const int32_t* __restrict__ send_nodes, const P2PDev pd, int bs,
                int* flags, int* ticket) {
  if (flags[Flag::DONE]) return;
  const long long seq = pd.base[0] + flags[Flag::ITERS] + 1;
  const long long total = pd.send_ptr[pd.n_nbr] * bs;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long i = t / bs;
    const int c = (int)(t - i * bs);
    int k = 0;
    while (k + 1 < pd.n_nbr && i >= pd.send_ptr[k + 1]) ++k;
    double* dst = pd.peer_z[k] + (pd.peer_ghost_start[k] + (i - pd.send_ptr[k])) * bs + c;
    *dst = z[(size_t)send_nodes[i] * bs + c];
  }
  __threadfence_system();
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    const int t = atomicAdd(ticket, 1);
    s_last = (t == (int)gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  if (threadIdx.x == 0) *ticket = 0;
  if ((int)threadIdx.x < pd.n_nbr) {
    const int k = threadIdx.x;
    __threadfence_system();
    st_release_sys(pd.peer_halo_flag[k] + pd.rank, seq);           // my entries of z have landed at neighbour k
    long long spins = 0;
    while (ld_acquire_sys(pd.my_halo_flag + pd.nbr[k]) < seq) {     // neighbour k's entries have landed here
      if (++spins > kSpinLimit) { flags[Flag::DONE] = 4; break; }
    }
  }
}

__global__ void __launch_bounds__(32)
p2p_allreduce_kernel(double* red, const P2PDev pd, int* flags) {
  if (flags[Flag::DONE]) return;
  const long long seq = pd.base[0] + flags[Flag::ITERS] + 1;
  const int par = (int)(seq & 1);
  const int lane = threadIdx.x;
  double v[4] = {0.0, 0.0, 0.0, 0.0};
  bool lost = false;
  if (lane < pd.world) {
    MailSlot* dst = pd.peer_mail[lane] + (pd.rank * 2 + par);
#pragma unroll
    for (int q = 0; q < 4; ++q) dst->v[q] = red[q];
    __threadfence_system();
    st_release_sys(&dst->seq, seq);
    const MailSlot* src = pd.my_mail + (lane * 2 + par);
    long long spins = 0;
    while (ld_acquire_sys(&src->seq) != seq) {
      if (++spins > kSpinLimit) { lost = true; break; }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] = *reinterpret_cast<const volatile double*>(&src->v[q]);
  }
  lost = __any_sync(0xffffffffu, lost);
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  for (int p = 0; p < pd.world; ++p) {        // rank order: identical association on every rank
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[q] += __shfl_sync(0xffffffffu, v[q], p);
  }
  if (lane == 0) {
    if (lost) flags[Flag::DONE] = 4;
#pragma unroll
    for (int q = 0; q < 4; ++q) red[q] = acc[q];
  }
}

__global__ void sub_owned_kernel(double* out, const double* f, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] -= f[i];
}

}  // namespace

int dist_halo_exchange(femb_handle* h, double* v, cudaStream_t stream, void* comm_v) {
  if (h->dist_world <= 1 || h->dist_nbr.empty()) return FEMB_OK;
  const int bs = h->bs;
  const int64_t n_send = h->dist_send_ptr.back();
  if (n_send > 0) {
    const int64_t tot = n_send * bs;
    pack_nodes_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(v, h->dist_send_nodes.p, n_send, bs, h->dist_send_buf.p);
    h->launches++;
    FEMB_CUDA(h, cudaGetLastError());
  }
  ncclComm_t comm = reinterpret_cast<ncclComm_t>(comm_v);
  FEMB_NCCL(h, g_nccl.GroupStart());
  for (size_t k = 0; k < h->dist_nbr.size(); ++k) {
    const int64_t ns = h->dist_send_ptr[k + 1] - h->dist_send_ptr[k];
    if (ns > 0)
      FEMB_NCCL(h, g_nccl.Send(h->dist_send_buf.p + h->dist_send_ptr[k] * bs, (size_t)(ns * bs), ncclDouble, h->dist_nbr[k], comm, stream));
    if (h->dist_recv_count[k] > 0)
      FEMB_NCCL(h, g_nccl.Recv(v + h->dist_recv_start[k] * bs, (size_t)(h->dist_recv_count[k] * bs), ncclDouble, h->dist_nbr[k], comm, stream));
  }
  FEMB_NCCL(h, g_nccl.GroupEnd());
  return FEMB_OK;
}

int run_dist_pcg(femb_handle* h, const femb_solve_opts& o, const double* d_rhs, femb_stats* st) {
  const int64_t n_all = h->ndof;
  const int64_t n = h->n_owned_nodes * h->bs;
  const int gridv = vec_grid(h, n, kRowThreads);
  const int pstride = h->num_sms * 8;
  int rc = FEMB_OK;
  if (!d_rhs) {
    rc = setup_rhs_for_direct(h);             // b = masked(f - K u0) on every local row
    if (rc) return rc;
    d_rhs = h->b.p;
  }
  // peer-memory path (femb_dist_p2p_import done and the exported z vector still the live one); fused: the exchanges
  // live inside the SpMV and update kernels; FEMB_DIST_P2P_KERNELS=1 keeps the two stand-alone exchange kernels
  const bool overlap = h->dist_world > 1 && h->dist_n_bnd > 0 && h->nccl_comm_halo && getenv("FEMB_DIST_OVERLAP");
  const bool p2p = h->dist_world > 1 && h->p2p_dev && h->p2p_z_exported == h->z.p && !overlap && !getenv("FEMB_DIST_NO_P2P");
  const bool fused = p2p && h->p2p_dev_copy.p && !getenv("FEMB_DIST_P2P_KERNELS");
  // line preconditioner on the partition (lines.cu): needs the fused peer-memory exchange; otherwise Jacobi
  bool lines = dist_lines_applicable(h, o, fused);
  rc = setup_precond_public(h, lines ? FEMB_PRECOND_JACOBI : o.precond);
  if (rc) return rc;
  if (lines) {
    rc = dist_lines_setup(h);
    if (rc < 0) return rc;
    if (rc > 0) lines = false;                  // a bundle matrix could not be factored: Jacobi
  }
  FEMB_CUDA(h, h->dist_red.alloc(Red::COUNT));
  FEMB_CUDA(h, cudaMemsetAsync(h->dist_red.p, 0, sizeof(double) * Red::COUNT, h->stream));
  FEMB_CUDA(h, cudaMemsetAsync(h->flags.p, 0, sizeof(int32_t) * Flag::COUNT, h->stream));
  // ghost tails start from zero (z's is overwritten by the first halo exchange)
  for (double* v : {h->x.p, h->r.p, h->z.p, h->p.p, h->q.p, h->s.p})
    FEMB_CUDA(h, cudaMemsetAsync(v, 0, sizeof(double) * n_all, h->stream));
  const bool blockj = !lines && (o.precond == FEMB_PRECOND_BLOCK_JACOBI);
  double* red = h->dist_red.p;
#define INIT(BS, BJ) dist_init_kernel<BS, kRowThreads, BJ><<<gridv, kRowThreads, 0, h->stream>>>(d_rhs, h->Dinv.p, h->x.p, h->r.p, h->z.p, h->p.p, h->q.p, n, h->partials.p + pstride, pstride, red, h->flags.p)
  if (lines) { /* the persistent kernel initialises its own vectors */ }
  else if (h->bs == 6) { if (blockj) INIT(6, true); else INIT(6, false); }
  else { if (blockj) INIT(3, true); else INIT(3, false); }
#undef INIT
  h->launches++;
  FEMB_CUDA(h, cudaGetLastError());

  ncclComm_t comm = reinterpret_cast<ncclComm_t>(h->nccl_comm);
  struct Peek { int32_t flags[Flag::COUNT]; double red[Red::COUNT]; };
  Peek* peek = reinterpret_cast<Peek*>(h->pinned);
  const int check = o.check_every > 0 ? o.check_every : 50;
  int it = 0, done = 0, spmv_launches = 0;
  // Overlapping the halo with the interior rows (FEMB_DIST_OVERLAP) is opt-in: at 1M DOF over 2 GPUs it measured SLOWER
  // (87.8 vs 83.0 us/iteration — the extra launches and the NCCL kernel competing for SMs cost more
  // than the ~12 us exchange they hide), at 8M DOF 3 % faster (profiles/r01_dist_2gpu.log).
  // one CG iteration, enqueued on the handle's stream (and the halo stream when overlapping)
  const P2PDev* pd = reinterpret_cast<const P2PDev*>(h->p2p_dev);
  if (p2p) {
    // line the ranks up once per solve (they arrive from host-side setup of different length): the
    // bounded waits of the peer-memory kernels are sized for iteration-scale skew, NCCL's is unbounded
    FEMB_NCCL(h, g_nccl.AllReduce(red + Red::DELTA2, red + Red::DELTA2, 1, ncclDouble, ncclSum, comm, h->stream));
    long long* hb = reinterpret_cast<long long*>(reinterpret_cast<char*>(h->pinned) + 2048);
    *hb = h->p2p_seq_base;
    FEMB_CUDA(h, cudaMemcpyAsync(h->p2p_base_dev, hb, sizeof(long long), cudaMemcpyHostToDevice, h->stream));
  }
  if (lines) {
    // line-preconditioned PCG: the whole iteration, exchanges included, runs inside one persistent kernel (lines.cu)
    rc = dist_lines_solve(h, o, d_rhs, st);
    FEMB_CUDA(h, cudaMemcpyAsync(peek->flags, h->flags.p, sizeof(peek->flags), cudaMemcpyDeviceToHost, h->stream));
    FEMB_CUDA(h, cudaStreamSynchronize(h->stream));
    h->p2p_seq_base += (long long)peek->flags[Flag::ITERS] + 4;   // identical on every rank
    return rc;
  }
  auto enqueue_iteration = [&](int first) -> int {
    int rc2;
    if (overlap) {
      // halo of z on its own stream / communicator while the rows that read no ghost column run
      FEMB_CUDA(h, cudaEventRecord(h->ev_vec, h->stream));
      FEMB_CUDA(h, cudaStreamWaitEvent(h->halo_stream, h->ev_vec, 0));
      rc2 = dist_halo_exchange(h, h->z.p, h->halo_stream, h->nccl_comm_halo);
      if (rc2) return rc2;
      FEMB_CUDA(h, cudaEventRecord(h->ev_halo, h->halo_stream));
      rc2 = launch_spmv_rows(h, h->z.p, h->s.p, n, true, h->partials.p, red + Red::DELTA, h->dist_bnd_flag.p, nullptr);
      if (rc2) return rc2;
      FEMB_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_halo, 0));
      rc2 = launch_spmv_rows(h, h->z.p, h->s.p, h->dist_n_bnd * h->bs, true, h->partials.p, red + Red::DELTA2, nullptr,
                             h->dist_bnd_nodes.p);
      if (rc2) return rc2;
      ++spmv_launches;
    } else if (p2p && fused && !first) {
      rc2 = launch_spmv_rows(h, h->z.p, h->s.p, n, true, h->partials.p, red + Red::DELTA, nullptr, nullptr, h->p2p_dev_copy.p);
      if (rc2) return rc2;
    } else if (p2p) {
      const long long tot = h->dist_send_ptr.back() * h->bs;
      const int gridh = (int)std::max<long long>(1, std::min<long long>((tot + 255) / 256, h->num_sms));
      p2p_halo_kernel<<<gridh, 256, 0, h->stream>>>(h->z.p, h->dist_send_nodes.p, *pd, h->bs, h->flags.p, h->p2p_ticket);
      h->launches++;
      rc2 = launch_spmv_rows(h, h->z.p, h->s.p, n, true, h->partials.p, red + Red::DELTA, nullptr, nullptr,
                             fused ? h->p2p_dev_copy.p : nullptr);
      if (rc2) return rc2;
    } else {
      rc2 = dist_halo_exchange(h, h->z.p, h->stream, h->nccl_comm);
      if (rc2) return rc2;
      rc2 = launch_spmv_rows(h, h->z.p, h->s.p, n, true, h->partials.p, red + Red::DELTA);   // local (z, s)
      if (rc2) return rc2;
    }
    ++spmv_launches;
    if (p2p && fused) {
      // nothing to launch: the SpMV's last CTA posted the partial sums, the update kernel collects them
    } else if (p2p) {
      p2p_allreduce_kernel<<<1, 32, 0, h->stream>>>(red, *pd, h->flags.p);
      h->launches++;
    } else if (h->dist_world > 1) {
      FEMB_NCCL(h, g_nccl.AllReduce(red, red, Red::NRED, ncclDouble, ncclSum, comm, h->stream));
    }
#define UPD(BS, BJ) dist_update_kernel<BS, kRowThreads, BJ><<<gridv, kRowThreads, 0, h->stream>>>(h->Dinv.p, h->s.p, h->p.p, h->q.p, h->x.p, h->r.p, h->z.p, n, first, o.rtol, h->partials.p + pstride, pstride, red, h->flags.p, (p2p && fused) ? reinterpret_cast<const P2PDev*>(h->p2p_dev_copy.p) : nullptr)
    if (h->bs == 6) { if (blockj) UPD(6, true); else UPD(6, false); }
    else { if (blockj) UPD(3, true); else UPD(3, false); }
#undef UPD
    h->launches++;
    return FEMB_OK;
  };
  auto poll = [&]() -> int {
    FEMB_CUDA(h, cudaGetLastError());
    FEMB_CUDA(h, cudaMemcpyAsync(peek->flags, h->flags.p, sizeof(peek->flags), cudaMemcpyDeviceToHost, h->stream));
    FEMB_CUDA(h, cudaMemcpyAsync(peek->red, red, sizeof(peek->red), cudaMemcpyDeviceToHost, h->stream));
    FEMB_CUDA(h, cudaStreamSynchronize(h->stream));
    done = peek->flags[Flag::DONE];
    return FEMB_OK;
  };
  // iteration 0 runs eagerly (NCCL connects lazily on first use; `first` differs) ...
  rc = enqueue_iteration(1);
  if (rc) return rc;
  it = 1;
  // ... the rest as a CUDA graph of `check` iterations — NCCL calls included — launched once per
  // poll: with 6-8 enqueues per iteration (two of them NCCL) the host, not the GPU, would
  // otherwise set the pace of a 1M-DOF/GPU iteration.
  // (a graph always runs `check` iterations: with a smaller iteration cap the eager path keeps the cap exact)
  const bool use_graph = h->dist_world > 1 && o.max_iter >= check && !getenv("FEMB_DIST_NO_GRAPH");
  (void)comm;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t gexec = nullptr;
  if (use_graph) {
    FEMB_CUDA(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeRelaxed));
    for (int k = 0; k < check && !rc; ++k) rc = enqueue_iteration(0);
    cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
    if (!rc && ce != cudaSuccess) rc = fail(h, FEMB_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce));
    if (!rc && cudaGraphInstantiate(&gexec, graph, 0) != cudaSuccess) rc = fail(h, FEMB_ERR_CUDA, "cudaGraphInstantiate failed");
    spmv_launches = overlap ? 2 : 1;
  }
  while (!rc && !done && it <= o.max_iter) {
    if (use_graph) {
      if (cudaGraphLaunch(gexec, h->stream) != cudaSuccess) { rc = fail(h, FEMB_ERR_CUDA, "cudaGraphLaunch failed"); break; }
      it += check;
      spmv_launches += (overlap ? 2 : 1) * check;
      h->launches += (overlap ? 4 : 3) * check;
    } else {
      const int batch = std::min(check, o.max_iter + 1 - it);
      for (int k = 0; k < batch && !rc; ++k, ++it) rc = enqueue_iteration(0);
      if (rc) break;
    }
    rc = poll();
    if (rc) break;
  }
  // every exit goes through here: graph objects released, and the sequence base advanced by an amount that does not
  // depend on where this rank stopped (its peers' mailbox / halo numbering must stay in step)
  if (gexec) cudaGraphExecDestroy(gexec);
  if (graph) cudaGraphDestroy(graph);
  if (p2p) h->p2p_seq_base += (long long)std::max(it, (int)peek->flags[Flag::ITERS]) + check + 4;
  if (rc) return rc;
  if (done == 4) return fail(h, FEMB_ERR_CUDA, "peer-memory exchange timed out waiting for another rank");
  if (st) {
    st->method_used = FEMB_SOLVER_PCG;
    st->iterations = peek->flags[Flag::ITERS];
    st->converged = (done == 1);
    st->spmv_launches = spmv_launches;
    st->precond_used = blockj ? FEMB_PRECOND_BLOCK_JACOBI : FEMB_PRECOND_JACOBI;
    st->op_used = ebe_available_dist(h) ? FEMB_OP_EBE : FEMB_OP_BSR;
    const double bb = peek->red[Red::BB];
    st->rel_residual = bb > 0.0 ? std::sqrt(peek->red[Red::RRFINAL] / bb) : 0.0;
  }
  if (done == 2) return fail(h, FEMB_ERR_SINGULAR, "distributed PCG breakdown: p^T K p <= 0 (K_ff is not positive definite)");
  if (done != 1) return fail(h, FEMB_ERR_NOT_CONVERGED, "distributed PCG did not reach rtol within max_iter");
  return FEMB_OK;
}

// ---- helpers for the distributed modal solve (modal.cu) ----------------------------------
bool dist_active(const femb_handle* h) { return h->dist_world > 1 && h->n_owned_nodes > 0 && h->nccl_comm; }

// in-place sum over ranks of `count` doubles on the device
int dist_allreduce(femb_handle* h, double* d_buf, int count) {
  if (!dist_active(h)) return FEMB_OK;
  FEMB_NCCL(h, g_nccl.AllReduce(d_buf, d_buf, (size_t)count, ncclDouble, ncclSum,
                                reinterpret_cast<ncclComm_t>(h->nccl_comm), h->stream));
  return FEMB_OK;
}

// y(owned rows) = K_ff x after refreshing the ghost tail of x (x is overwritten there)
int dist_spmv_masked(femb_handle* h, double* x, double* y) {
  int rc = dist_halo_exchange(h, x, h->stream, h->nccl_comm);
  if (rc) return rc;
  return launch_spmv_rows(h, x, y, h->n_owned_nodes * h->bs, true, nullptr, h->scal.p);
}

// K_ff x = b on the owned rows (solution in h->x, owned prefix)
int dist_solve_rhs(femb_handle* h, const femb_solve_opts& o, const double* d_b, femb_stats* st) {
  return run_dist_pcg(h, o, d_b, st);
}

}  // namespace femb

using namespace femb;

extern "C" {

int femb_dist_unique_id(uint8_t* id128) {
  if (!id128) return FEMB_ERR_ARG;
  if (load_nccl()) return FEMB_ERR_CUDA;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) return FEMB_ERR_CUDA;
  std::memcpy(id128, &id, 128);
  return FEMB_OK;
}

int femb_dist_init(femb_handle* h, int rank, int world, const uint8_t* id128) {
  if (!h || world < 1 || rank < 0 || rank >= world || !id128) return fail(h, FEMB_ERR_ARG, "bad femb_dist_init arguments");
  if (const char* e = load_nccl()) return fail(h, FEMB_ERR_CUDA, e);
  FEMB_CUDA(h, cudaSetDevice(h->device));
  femb_dist_finalize(h);
  ncclUniqueId id;
  std::memcpy(&id, id128, 128);
  ncclComm_t comm = nullptr;
  FEMB_NCCL(h, g_nccl.CommInitRank(&comm, world, id, rank));
  h->nccl_comm = comm;
  h->dist_rank = rank;
  h->dist_world = world;
  if (world > 1 && g_nccl.CommSplit) {
    ncclComm_t comm2 = nullptr;
    FEMB_NCCL(h, g_nccl.CommSplit(comm, 0, rank, &comm2, nullptr));
    h->nccl_comm_halo = comm2;
    if (!h->halo_stream) FEMB_CUDA(h, cudaStreamCreateWithFlags(&h->halo_stream, cudaStreamNonBlocking));
    if (!h->ev_vec) FEMB_CUDA(h, cudaEventCreateWithFlags(&h->ev_vec, cudaEventDisableTiming));
    if (!h->ev_halo) FEMB_CUDA(h, cudaEventCreateWithFlags(&h->ev_halo, cudaEventDisableTiming));
  }
  return FEMB_OK;
}

void femb_dist_finalize(femb_handle* h) {
  if (h && h->nccl_comm && g_nccl.CommDestroy) {
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    if (h->halo_stream) cudaStreamSynchronize(h->halo_stream);
    if (h->nccl_comm_halo) g_nccl.CommDestroy(reinterpret_cast<ncclComm_t>(h->nccl_comm_halo));
    g_nccl.CommDestroy(reinterpret_cast<ncclComm_t>(h->nccl_comm));
    h->nccl_comm = h->nccl_comm_halo = nullptr;
  }
  if (h) {
    for (void* p : h->p2p_mapped) cudaIpcCloseMemHandle(p);
    h->p2p_mapped.clear();
    if (h->p2p_dev) { delete reinterpret_cast<P2PDev*>(h->p2p_dev); h->p2p_dev = nullptr; }
  }
  if (h && h->halo_stream) { cudaStreamDestroy(h->halo_stream); h->halo_stream = nullptr; }
  if (h && h->ev_vec) { cudaEventDestroy(h->ev_vec); cudaEventDestroy(h->ev_halo); h->ev_vec = h->ev_halo = nullptr; }
}

int femb_dist_set_halo(femb_handle* h, int64_t n_owned_nodes, int32_t n_nbr, const int32_t* nbr_rank,
                       const int64_t* send_ptr, const int32_t* send_nodes, const int64_t* recv_start,
                       const int64_t* recv_count) {
  if (!h || h->kind == Kind::None) return fail(h, FEMB_ERR_ARG, "set the local mesh first");
  if (n_owned_nodes < 0 || n_owned_nodes > h->n_nodes || n_nbr < 0) return fail(h, FEMB_ERR_ARG, "bad halo arguments");
  if (n_nbr > 0 && (!nbr_rank || !send_ptr || !recv_start || !recv_count)) return fail(h, FEMB_ERR_ARG, "bad halo arguments");
  if (n_nbr > 0 && !h->nccl_comm) return fail(h, FEMB_ERR_ARG, "call femb_dist_init before femb_dist_set_halo");
  FEMB_CUDA(h, cudaSetDevice(h->device));
  h->n_owned_nodes = n_owned_nodes;
  h->dist_nbr.assign(nbr_rank, nbr_rank + n_nbr);
  h->dist_send_ptr.assign(1, 0);
  if (n_nbr > 0) h->dist_send_ptr.assign(send_ptr, send_ptr + n_nbr + 1);
  h->dist_recv_start.assign(recv_start, recv_start + n_nbr);
  h->dist_recv_count.assign(recv_count, recv_count + n_nbr);
  const int64_t n_send = h->dist_send_ptr.back();
  for (int k = 0; k < n_nbr; ++k) {
    if (nbr_rank[k] < 0 || nbr_rank[k] >= h->dist_world || nbr_rank[k] == h->dist_rank) return fail(h, FEMB_ERR_ARG, "bad neighbour rank");
    if (recv_start[k] < n_owned_nodes || recv_start[k] + recv_count[k] > h->n_nodes) return fail(h, FEMB_ERR_ARG, "ghost range outside the local tail");
  }
  for (int64_t i = 0; i < n_send; ++i)
    if (send_nodes[i] < 0 || send_nodes[i] >= n_owned_nodes) return fail(h, FEMB_ERR_ARG, "send node is not owned");
  // owned block rows that read a ghost column wait for the halo; all the others overlap with it
  h->dist_n_bnd = 0;
  if (n_nbr > 0) {
    if (!h->have_symbolic) return fail(h, FEMB_ERR_ARG, "call femb_assemble before femb_dist_set_halo");
    const Symbolic& S = h->sym;
    std::vector<uint8_t> flag((size_t)h->n_nodes, 0);
    std::vector<int32_t> list;
    for (int64_t i = 0; i < n_owned_nodes; ++i)
      for (int32_t b = S.rowptr[i]; b < S.rowptr[i + 1]; ++b)
        if (S.colidx[b] >= n_owned_nodes) { flag[i] = 1; list.push_back((int32_t)i); break; }
    h->dist_n_bnd = (int64_t)list.size();
    FEMB_CUDA(h, upload(h->dist_bnd_flag, flag, h->stream));
    FEMB_CUDA(h, upload(h->dist_bnd_nodes, list, h->stream));
    FEMB_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  FEMB_CUDA(h, upload(h->dist_send_nodes, send_nodes, (size_t)n_send, h->stream));
  FEMB_CUDA(h, h->dist_send_buf.alloc((size_t)std::max<int64_t>(1, n_send * h->bs)));
  FEMB_CUDA(h, cudaStreamSynchronize(h->stream));
  return FEMB_OK;
}

int femb_dist_set_lines(femb_handle* h, int32_t n_coarse, const int32_t* fam_off, const int32_t* node_bundle,
                        const int32_t* node_line, const int32_t* node_pos, const double* node_dir) {
  if (!h || !fam_off || !node_bundle || !node_line || !node_pos || !node_dir) return fail(h, FEMB_ERR_ARG, "bad femb_dist_set_lines arguments");
  FEMB_CUDA(h, cudaSetDevice(h->device));
  return dist_set_lines(h, n_coarse, fam_off, node_bundle, node_line, node_pos, node_dir);
}

int femb_dist_p2p_export(femb_handle* h, uint8_t* handles128) {
  if (!h || !handles128) return FEMB_ERR_ARG;
  if (!h->have_bc || !h->z.p) return fail(h, FEMB_ERR_ARG, "call femb_set_bc before femb_dist_p2p_export");
  if (h->dist_world > kMaxRanks) return fail(h, FEMB_ERR_ARG, "peer-memory path supports up to 8 ranks");
  FEMB_CUDA(h, cudaSetDevice(h->device));
  // ONE dedicated allocation per rank — mailboxes, flags and the z vector itself — of at least
  // 2 MB, so that it is never a sub-allocation of a block shared with other buffers: the IPC
  // handle then maps exactly this memory at offset 0 in every peer.
  const size_t zbytes = ((size_t)h->ndof * sizeof(double) + 255) / 256 * 256;
  // behind z: the flag-in-data slots of the persistent line-preconditioned PCG (lines.cu; 16 bytes per value):
  // scalars [world][2][4] | coarse residuals [world][2][kLnMaxCoarse] | halo [ghost nodes][6]
  const size_t n_ghost = (size_t)(h->n_nodes - h->n_owned_nodes);
  const size_t llbytes = kLLScalBytes + ((size_t)h->dist_world * 2 * kLnMaxCoarse + n_ghost * 6) * sizeof(uint4);
  const size_t bytes = std::max<size_t>(kP2PZOffset + zbytes + llbytes, (size_t)4 << 20);
  h->z.release();                       // z moves into the exported allocation
  FEMB_CUDA(h, h->p2p_comm.alloc(bytes));
  FEMB_CUDA(h, cudaMemset(h->p2p_comm.p, 0, bytes));
  h->z.adopt(reinterpret_cast<double*>(h->p2p_comm.p + kP2PZOffset), (size_t)h->ndof);
  cudaIpcMemHandle_t hc;
  FEMB_CUDA(h, cudaIpcGetMemHandle(&hc, h->p2p_comm.p));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  std::memset(handles128, 0, 128);
  std::memcpy(handles128, &hc, 64);
  const unsigned long long zb = zbytes;            // where this rank's flag-in-data slots start behind its z vector
  std::memcpy(handles128 + 64, &zb, sizeof(zb));
  const long long no = h->n_owned_nodes;           // ... and where its ghost tail starts (halo slots are indexed from there)
  std::memcpy(handles128 + 72, &no, sizeof(no));
  h->p2p_z_exported = h->z.p;
  return FEMB_OK;
}

int femb_dist_p2p_import(femb_handle* h, const uint8_t* all_handles, const int64_t* peer_ghost_start) {
  if (!h || !all_handles) return FEMB_ERR_ARG;
  if (!h->p2p_comm.p || h->p2p_z_exported != h->z.p) return fail(h, FEMB_ERR_ARG, "call femb_dist_p2p_export first");
  if (h->dist_nbr.size() > 0 && !peer_ghost_start) return fail(h, FEMB_ERR_ARG, "peer_ghost_start missing");
  FEMB_CUDA(h, cudaSetDevice(h->device));
  const int world = h->dist_world, rank = h->dist_rank;
  P2PDev* pd = new P2PDev();
  std::memset(pd, 0, sizeof(P2PDev));
  char* cb = reinterpret_cast<char*>(h->p2p_comm.p);
  auto mail_of = [&](char* base) { return reinterpret_cast<MailSlot*>(base); };
  auto flag_of = [&](char* base) { return reinterpret_cast<long long*>(base + sizeof(MailSlot) * world * 2); };
  pd->my_mail = mail_of(cb);
  pd->my_halo_flag = flag_of(cb);
  h->p2p_base_dev = flag_of(cb) + world;
  h->p2p_ticket = reinterpret_cast<int*>(flag_of(cb) + world + 1);
  const size_t zbytes = ((size_t)h->ndof * sizeof(double) + 255) / 256 * 256;
  pd->base = h->p2p_base_dev;
  pd->world = world; pd->rank = rank;
  std::vector<char*> peer_comm(world, nullptr);
  for (int p = 0; p < world; ++p) {
    if (p == rank) { peer_comm[p] = cb; continue; }
    cudaIpcMemHandle_t hc;
    std::memcpy(&hc, all_handles + (size_t)p * 128, 64);
    void* ptr = nullptr;
    FEMB_CUDA(h, cudaIpcOpenMemHandle(&ptr, hc, cudaIpcMemLazyEnablePeerAccess));
    peer_comm[p] = reinterpret_cast<char*>(ptr);
    h->p2p_mapped.push_back(ptr);
  }
  for (int p = 0; p < world; ++p) pd->peer_mail[p] = mail_of(peer_comm[p]);
  // every rank's local vector has its own length, so a peer's mail area sits behind ITS z: the offsets travel with the
  // handles (femb_dist_p2p_export writes the rank's z bytes into its 128-byte blob)
  std::vector<uint4*> peer_halo(world, nullptr);
  std::vector<long long> peer_owned(world, 0);
  for (int p = 0; p < world; ++p) {
    unsigned long long zb = 0;
    std::memcpy(&zb, all_handles + (size_t)p * 128 + 64, sizeof(zb));
    std::memcpy(&peer_owned[p], all_handles + (size_t)p * 128 + 72, sizeof(long long));
    if (p == rank) { zb = zbytes; peer_owned[p] = h->n_owned_nodes; }
    char* ll = peer_comm[p] + kP2PZOffset + zb;
    pd->peer_ll_scal[p] = reinterpret_cast<uint4*>(ll);
    pd->peer_ll_rb[p] = reinterpret_cast<uint4*>(ll + kLLScalBytes);
    peer_halo[p] = pd->peer_ll_rb[p] + (size_t)world * 2 * kLnMaxCoarse;
  }
  pd->my_ll_scal = pd->peer_ll_scal[rank];
  pd->my_ll_rb = pd->peer_ll_rb[rank];
  pd->my_ll_halo = peer_halo[rank];
  pd->n_owned = h->n_owned_nodes;
  pd->n_nbr = (int)h->dist_nbr.size();
  if (pd->n_nbr > kMaxRanks) { delete pd; return fail(h, FEMB_ERR_ARG, "too many neighbour ranks"); }
  for (int k = 0; k < pd->n_nbr; ++k) {
    const int p = h->dist_nbr[k];
    pd->peer_z[k] = reinterpret_cast<double*>(peer_comm[p] + kP2PZOffset);
    pd->peer_ghost_start[k] = peer_ghost_start[k];
    pd->peer_halo_flag[k] = flag_of(peer_comm[p]);
    pd->nbr[k] = p;
    pd->send_ptr[k] = h->dist_send_ptr[k];
    pd->peer_ll_halo[k] = peer_halo[p] + (size_t)(peer_ghost_start[k] - peer_owned[p]) * 6;
    pd->recv_start[k] = h->dist_recv_start[k];
    pd->recv_count[k] = h->dist_recv_count[k];
  }
  pd->send_ptr[pd->n_nbr] = h->dist_send_ptr.back();
  // fused mode tables: destination of every owned node's entries
  {
    std::vector<int32_t> slot((size_t)h->n_owned_nodes, -1), extra;
    std::vector<int32_t> snodes((size_t)h->dist_send_ptr.back());
    FEMB_CUDA(h, cudaMemcpy(snodes.data(), h->dist_send_nodes.p, snodes.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
    bool ok = true;
    for (int k = 0; k < pd->n_nbr && ok; ++k)
      for (int64_t i = h->dist_send_ptr[k]; i < h->dist_send_ptr[k + 1]; ++i) {
        const int64_t idx = peer_ghost_start[k] + (i - h->dist_send_ptr[k]);
        if (idx >= (1 << 28)) { ok = false; break; }
        const int32_t sl = (int32_t)((k << 28) | (int32_t)idx);
        const int32_t node = snodes[i];
        if (slot[node] < 0) slot[node] = sl;
        else { extra.push_back(node); extra.push_back(sl); }
      }
    if (ok) {
      // boundary nodes (ascending) and, per boundary node, all its destinations (k << 28 | offset among MY nodes at k)
      std::vector<std::vector<int32_t>> dst_of((size_t)h->n_owned_nodes);
      for (int k = 0; k < pd->n_nbr; ++k)
        for (int64_t i = h->dist_send_ptr[k]; i < h->dist_send_ptr[k + 1]; ++i)
          dst_of[snodes[i]].push_back((int32_t)((k << 28) | (int32_t)(i - h->dist_send_ptr[k])));
      std::vector<int32_t> bnd, dptr(1, 0), dst;
      for (int64_t i = 0; i < h->n_owned_nodes; ++i)
        if (slot[i] >= 0) {
          bnd.push_back((int32_t)i);
          dst.insert(dst.end(), dst_of[i].begin(), dst_of[i].end());
          dptr.push_back((int32_t)dst.size());
        }
      pd->n_bnd = (int)bnd.size();
      if (bnd.empty()) bnd.push_back(0);
      if (dst.empty()) dst.push_back(0);
      FEMB_CUDA(h, upload(h->p2p_bnd_nodes, bnd, h->stream));
      FEMB_CUDA(h, upload(h->p2p_bnd_dst_ptr, dptr, h->stream));
      FEMB_CUDA(h, upload(h->p2p_bnd_dst, dst, h->stream));
      pd->bnd_nodes = h->p2p_bnd_nodes.p;
      pd->bnd_dst_ptr = h->p2p_bnd_dst_ptr.p;
      pd->bnd_dst = h->p2p_bnd_dst.p;
      FEMB_CUDA(h, upload(h->p2p_send_slot, slot, h->stream));
      if (extra.empty()) { extra.push_back(0); extra.push_back(0); pd->n_extra = 0; }
      else pd->n_extra = (int)(extra.size() / 2);
      FEMB_CUDA(h, upload(h->p2p_extra, extra, h->stream));
      pd->send_slot = h->p2p_send_slot.p;
      pd->extra = h->p2p_extra.p;
      FEMB_CUDA(h, h->p2p_dev_copy.alloc(sizeof(P2PDev)));
      FEMB_CUDA(h, cudaMemcpyAsync(h->p2p_dev_copy.p, pd, sizeof(P2PDev), cudaMemcpyHostToDevice, h->stream));
      FEMB_CUDA(h, cudaStreamSynchronize(h->stream));
    }
  }
  if (h->p2p_dev) delete reinterpret_cast<P2PDev*>(h->p2p_dev);
  h->p2p_dev = pd;
  h->p2p_seq_base = 2;          // never 0: the flags start at 0 and the first coarse-residual exchange of a solve is numbered `base`
  return FEMB_OK;
}

int femb_dist_solve_static(femb_handle* h, const femb_solve_opts* opts, int minus_f, double* u_owned,
                           double* reactions_owned, femb_stats* stats) {
  if (!h) return FEMB_ERR_ARG;
  if (!h->assembled || !h->have_bc) return fail(h, FEMB_ERR_ARG, "call femb_assemble and femb_set_bc first");
  if (h->n_owned_nodes <= 0) return fail(h, FEMB_ERR_ARG, "call femb_dist_set_halo first");
  if (h->dist_world > 1 && !h->nccl_comm) return fail(h, FEMB_ERR_ARG, "call femb_dist_init first");
  if (h->dist_world > 1) { if (const char* e = load_nccl()) return fail(h, FEMB_ERR_CUDA, e); }
  FEMB_CUDA(h, cudaSetDevice(h->device));
  femb_solve_opts o;
  if (opts) o = *opts;
  else { std::memset(&o, 0, sizeof(o)); o.precond = FEMB_PRECOND_JACOBI; o.check_every = 50; }
  if (o.max_iter <= 0) o.max_iter = 200000;
  if (!(o.rtol > 0.0)) o.rtol = 1e-12;
  femb_stats st;
  std::memset(&st, 0, sizeof(st));
  h->launches = 0;
  FEMB_CUDA(h, cudaEventRecord(h->ev0, h->stream));
  int rc = run_dist_pcg(h, o, nullptr, &st);
  const int64_t n = h->n_owned_nodes * h->bs;
  if (rc == FEMB_OK) {
    h->have_solution = true;
    rc = dist_halo_exchange(h, h->x.p, h->stream, h->nccl_comm);   // ghosts of u for K_full u (and stress)
    if (rc == FEMB_OK) rc = launch_spmv_rows(h, h->x.p, h->q.p, n, false, nullptr, h->scal.p);
    if (rc == FEMB_OK && minus_f) {
      sub_owned_kernel<<<vec_grid(h, n, 256), 256, 0, h->stream>>>(h->q.p, h->f.p, n);
      h->launches++;
    }
    if (rc == FEMB_OK && reactions_owned) FEMB_CUDA(h, download(reactions_owned, h->q.p, (size_t)n * 8, h->stream));
    if (rc == FEMB_OK && u_owned) FEMB_CUDA(h, download(u_owned, h->x.p, (size_t)n * 8, h->stream));
  }
  cudaEventRecord(h->ev1, h->stream);
  cudaError_t se = cudaStreamSynchronize(h->stream);
  if (rc == FEMB_OK && se != cudaSuccess) rc = fail(h, FEMB_ERR_CUDA, cudaGetErrorString(se));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, h->ev0, h->ev1);
  st.device_ms = ms;
  st.kernel_launches = (int32_t)h->launches;
  if (stats) *stats = st;
  return rc;
}

}  // extern "C"
