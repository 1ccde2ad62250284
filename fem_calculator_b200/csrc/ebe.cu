// Matrix-free ("element-by-element") form of the frame operator y = K_ff x for the Krylov loops.
//
// The assembled BSR SpMV (solver.cu) streams 8*36 bytes per block — 359 MB per product at 1M DOF,
// three times the L2 — and sits at ~75 % of the measured HBM bandwidth: it cannot get faster
// without moving fewer bytes.  A frame element's four 6x6 blocks, however, are a closed form of
// 19 numbers (the direction-cosine rows t, n1, n2 and ten stiffness magnitudes,
// BeamSolver.py:378-388, 646-660) that follow from two coordinate triples and a section row.  This
// kernel therefore never reads K: for every (node, incident element end) pair — the 16-byte pair
// records of the assembly kernel — it rebuilds the element record in registers, projects the two end
// displacements on (t, n1, n2), applies the ten magnitudes and rotates the 6-vector back:
//     y_node += K_e[a][a] x_node + K_e[a][1-a] x_other            (BeamSolver.py:387, 390-393)
// ~135 FP64 instructions per pair instead of 2 x 288 bytes of matrix; 40 MB of compulsory traffic
// per product at 1M DOF (16 B/pair + node records + coordinates + x + y + mask) instead of 359 MB.
//
// Mapping ("node gather"): LPN = (NBT / NB) * T lanes share one node — lane (qg, part) applies the
// pairs part, part+T, ... of the node's list to the NB right-hand sides [qg*NB, qg*NB+NB) and keeps
// the 6*NB sums in registers; the T partial sums are combined with a fixed xor-shuffle tree, lane
// part 0 masks, stores and accumulates (x, y).  No shared memory, no barriers, no float atomics:
// bit-reproducible run to run.  The node's own coordinates and x entries are loaded once per node.
//
// Measured alternatives at 1M DOF on one B200 (profiles/r01_ebe_operator_variants.log), single vector,
// back to back: one thread per pair with the per-node sum in shared memory (two phases, one barrier
// per 128-pair tile) 29.0 us; this kernel with T = 1 / 2 / 4 lanes per node 22.3 / 20.5 / 36.9 us;
// a three-stage software pipeline over the pair loop (record two ahead, operands one ahead) 21.7 us;
// per-element records stored after assembly (128 B: t, n1, magnitudes) and read back per pair
// 28.5-31 us — the gathers, not the FP64 work, set the time (ncu: FP64 pipe 41 %, DRAM 13 %), which
// is also why trimming the record from ~100 to ~75 FP64 instructions did not move it, nor did padded
// coordinates with one 32-byte gather per node and 32-byte section-row loads (21.7 us).  Four vectors:
// 2 vectors per lane / 2 lanes per node 53 us against 96 us for the shared-memory form and 123 us for
// the assembled SpMM.
//
// The assembled K is still produced (Jacobi diagonal, reactions r = K u - f, CSR export); meshes the
// pair view cannot describe (duplicate members, node degree > 127), Tet10 and the row-block distributed
// solver keep the BSR operator.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "ebe.cuh"
#include "elements.cuh"
#include "pcg_common.cuh"

namespace femb {

template <int NBT, int NB>
__device__ __forceinline__ void load_u(const double* __restrict__ x, int node, int qg, double (&u)[NB][6]) {
  if constexpr (NBT == 1) {
    load6(x, node, u[0]);
  } else if constexpr (NB == 1) {
#pragma unroll
    for (int c = 0; c < 6; ++c) u[0][c] = __ldg(x + ((size_t)node * 6 + c) * NBT + qg);
  } else if constexpr (NB == 2) {
    const double2* p = reinterpret_cast<const double2*>(x + (size_t)node * 6 * NBT + 2 * qg);
#pragma unroll
    for (int c = 0; c < 6; ++c) { const double2 v = __ldg(p + c * (NBT / 2)); u[0][c] = v.x; u[1][c] = v.y; }
  } else {
    static_assert(NB == 4 && NBT == 4, "vector grouping");
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      double4 v;
      asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w)
          : "l"(x + ((size_t)node * 6 + c) * 4));
      u[0][c] = v.x; u[1][c] = v.y; u[2][c] = v.z; u[3][c] = v.w;
    }
  }
}

// L2-coherent gather of a ghost node's entries (single vector): peer GPUs store them while CTAs of this
// kernel may already be resident, so the non-coherent path could serve stale L1 lines (solver.cu)
__device__ __forceinline__ void load6_cg(const double* __restrict__ x, int node, double* u) {
  const double2* p = reinterpret_cast<const double2*>(x + (size_t)node * 6);
  const double2 a = __ldcg(p), b = __ldcg(p + 1), c = __ldcg(p + 2);
  u[0] = a.x; u[1] = a.y; u[2] = b.x; u[3] = b.y; u[4] = c.x; u[5] = c.y;
}

template <int NBT, int NB>
__device__ __forceinline__ void store_u(double* __restrict__ y, int node, int qg, const double (&u)[NB][6]) {
  if constexpr (NBT == 1) {
    double2* p = reinterpret_cast<double2*>(y + (size_t)node * 6);
    p[0] = make_double2(u[0][0], u[0][1]); p[1] = make_double2(u[0][2], u[0][3]); p[2] = make_double2(u[0][4], u[0][5]);
  } else if constexpr (NB == 1) {
#pragma unroll
    for (int c = 0; c < 6; ++c) y[((size_t)node * 6 + c) * NBT + qg] = u[0][c];
  } else if constexpr (NB == 2) {
    double2* p = reinterpret_cast<double2*>(y + (size_t)node * 6 * NBT + 2 * qg);
#pragma unroll
    for (int c = 0; c < 6; ++c) p[c * (NBT / 2)] = make_double2(u[0][c], u[1][c]);
  } else {
#pragma unroll
    for (int c = 0; c < 6; ++c)
      *reinterpret_cast<double4*>(y + ((size_t)node * 6 + c) * 4) = make_double4(u[0][c], u[1][c], u[2][c], u[3][c]);
  }
}

// p2p (DOT, single vector, row-block distributed PCG — dist.cu): the first n_nodes local nodes are
// the rank's owned rows; before gathering, the kernel waits for the neighbours' halo flags, ghost
// entries (node id >= n_nodes) are read L2-coherently, and the last CTA of the (x, y) reduction posts
// the rank's partial sums into every peer's mailbox — the same three hooks as bsr_spmv_kernel.
// LINK: the launch belongs to the linked single-vector PCG (pcg_common.cuh): the kernel first
// finishes the update kernel's reductions (convergence decision) and publishes its own (x, y)
// partial instead of running a last-CTA reduction.  Otherwise DOT adds (x_q, y_q) through the ordered
// grid reduction into scal[0..NBT) and `done` (may be null) is the early-exit flag of the caller.
template <int NBT, int NB, int T, bool MASKED, bool DOT, bool LINK, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
frame_ebe_node_kernel(const FrameParams P, const int4* __restrict__ pair_rec, const int4* __restrict__ node_rec,
                      int n_nodes, const uint8_t* __restrict__ free_mask, const double* __restrict__ x,
                      double* __restrict__ y, double* partials, int pstride, double* scal, int* ticket,
                      const int* done, const PcgLink link, const P2PDev* __restrict__ p2p) {
  constexpr int QS = NBT / NB;
  constexpr int LPN = QS * T;
  constexpr int NPC = THREADS / LPN;   // nodes per CTA and step
  static_assert(32 % LPN == 0 && THREADS % 32 == 0, "a node's lanes share a warp");
  static_assert(!LINK || DOT, "linked reductions publish the (x, y) partials");
  if (LINK && NBT == 1) {
    __shared__ double s_link[2 * THREADS / 32];
    if (link.flags[Flag::DONE]) return;
    if (pcg_link_decide<THREADS>(link, s_link)) return;
  } else if (DOT && done && *done) {   // lockstep PCG (NBT > 1), linked or not: early-exit flag of the caller
    return;
  }
  if (DOT && !LINK && NBT == 1 && p2p) {
    // done points at flags[Flag::DONE] (= flags[0]); the iteration counter sits next to it
    if ((int)threadIdx.x < p2p->n_nbr) {
      const long long seq = p2p->base[0] + done[Flag::ITERS] + 1;
      long long spins = 0;
      while (ld_acquire_sys(p2p->my_halo_flag + p2p->nbr[threadIdx.x]) < seq) {
        if (++spins > kSpinLimit) { const_cast<int*>(done)[Flag::DONE] = 4; break; }
      }
    }
    __syncthreads();
  }
  const int lin = threadIdx.x % LPN;
  const int qg = lin / T, part = lin % T;
  double dot[NB];
#pragma unroll
  for (int q = 0; q < NB; ++q) dot[q] = 0.0;
  // Work split, measured at 1M DOF: one vector is fastest when the CTAs sweep the mesh together
  // (interleaved steps of NPC nodes: the neighbours' x entries are in lines other CTAs just touched,
  // 20.7 vs 22.6 us); four vectors — four times the gather footprint — prefer one contiguous node
  // range per CTA (50.7 vs 53.3 us).
  constexpr bool kContiguous = NBT > 1;
  const int per = kContiguous ? (n_nodes + (int)gridDim.x - 1) / (int)gridDim.x : n_nodes;
  const int lo = kContiguous ? blockIdx.x * per : blockIdx.x * NPC;
  const int hi = kContiguous ? min(n_nodes, lo + per) : n_nodes;
  const int step = kContiguous ? NPC : (int)gridDim.x * NPC;
  for (int base = lo; base < hi; base += step) {
    const int node = base + threadIdx.x / LPN;
    const bool active = node < hi;
    int first = 0, count = 0;
    double px = 0.0, py = 0.0, pz = 0.0;
    double ua[NB][6], acc[NB][6];
#pragma unroll
    for (int q = 0; q < NB; ++q)
#pragma unroll
      for (int c = 0; c < 6; ++c) { ua[q][c] = 0.0; acc[q][c] = 0.0; }
    if (active) {
      const int4 nr = __ldg(node_rec + node);
      first = nr.x; count = nr.y;
      const double* pp = P.xyz + 3 * (size_t)node;
      px = __ldg(pp); py = __ldg(pp + 1); pz = __ldg(pp + 2);
      load_u<NBT, NB>(x, node, qg, ua);
    }
#pragma unroll 2
    for (int j = part; j < count; j += T) {
      const int4 rec = __ldg(pair_rec + first + j);
      const int a = (rec.w >> 24) & 1;
      const double* po = P.xyz + 3 * (size_t)rec.y;
      const double ox = __ldg(po), oy = __ldg(po + 1), oz = __ldg(po + 2);
      const double* sp = P.sec_props + 8 * (size_t)(rec.w & 0xFFFFFF);
      FrameIn in;   // element direction: end 0 -> end 1, as in the assembly kernel
      in.dx = a ? px - ox : ox - px; in.dy = a ? py - oy : oy - py; in.dz = a ? pz - oz : oz - pz;
      in.A = __ldg(sp); in.Ix = __ldg(sp + 1); in.Iy = __ldg(sp + 2); in.J = __ldg(sp + 3);
      in.ky = __ldg(sp + 4); in.kz = __ldg(sp + 5);
      double uo[NB][6];
      if (NBT == 1 && DOT && !LINK && p2p && rec.y >= n_nodes) load6_cg(x, rec.y, uo[0]);
      else load_u<NBT, NB>(x, rec.y, qg, uo);
      KRec k;
      krec_from(P, in, a, k);
#pragma unroll
      for (int q = 0; q < NB; ++q) {
        double o6[6];
        ebe_apply(k, ua[q], uo[q], o6);
#pragma unroll
        for (int c = 0; c < 6; ++c) acc[q][c] += o6[c];
      }
    }
    if (T > 1) {
#pragma unroll
      for (int off = 1; off < T; off <<= 1)
#pragma unroll
        for (int q = 0; q < NB; ++q)
#pragma unroll
          for (int c = 0; c < 6; ++c) acc[q][c] += __shfl_xor_sync(0xffffffffu, acc[q][c], off);
    }
    if (active && part == 0) {
      if (MASKED) {
        const uint8_t* fm = free_mask + (size_t)node * 6;
#pragma unroll
        for (int c = 0; c < 6; ++c)
          if (!fm[c]) {      // identity rows on the fixed DOFs
#pragma unroll
            for (int q = 0; q < NB; ++q) acc[q][c] = ua[q][c];
          }
      }
      store_u<NBT, NB>(y, node, qg, acc);
      if (DOT) {
#pragma unroll
        for (int q = 0; q < NB; ++q)
#pragma unroll
          for (int c = 0; c < 6; ++c) dot[q] += ua[q][c] * acc[q][c];
      }
    }
  }
  if (LINK) {
    // publish this CTA's partial sums, one array per vector: op_partials[q * pstride + cta]
    __shared__ double s_pub[NBT * THREADS / 32];
    double mine[NBT];
#pragma unroll
    for (int q = 0; q < NBT; ++q) mine[q] = 0.0;
#pragma unroll
    for (int q = 0; q < NB; ++q) {
#pragma unroll
      for (int g = 0; g < QS; ++g)
        if (g == qg) mine[g * NB + q] = dot[q];
    }
    block_sum_all<THREADS, NBT>(mine, s_pub);
    if (threadIdx.x == 0) {
#pragma unroll
      for (int q = 0; q < NBT; ++q) link.op_partials[(size_t)q * link.pstride + blockIdx.x] = mine[q];
    }
  } else if (DOT) {
    double mine[NBT], tot[NBT];
#pragma unroll
    for (int q = 0; q < NBT; ++q) mine[q] = 0.0;
#pragma unroll
    for (int q = 0; q < NB; ++q) {
#pragma unroll
      for (int g = 0; g < QS; ++g)
        if (g == qg) mine[g * NB + q] = dot[q];
    }
    if (grid_reduce<THREADS, NBT>(mine, partials, pstride, ticket, tot)) {
      if (threadIdx.x == 0)
        for (int q = 0; q < NBT; ++q) scal[q] = tot[q];
      if (NBT == 1 && p2p && (int)threadIdx.x < p2p->world) {
        // post {delta, gamma, ||r||^2} of this rank in every peer's mailbox (scal = the solver's red[])
        const long long seq = p2p->base[0] + done[Flag::ITERS] + 1;
        MailSlot* dst = p2p->peer_mail[threadIdx.x] + (p2p->rank * 2 + (int)(seq & 1));
        dst->v[0] = tot[0]; dst->v[1] = scal[1]; dst->v[2] = scal[2]; dst->v[3] = 0.0;
        __threadfence_system();
        st_release_sys(&dst->seq, seq);
      }
    }
  }
}

static FrameParams ebe_params(const femb_handle* h) {
  FrameParams P;
  P.xyz = h->xyz.p; P.conn = h->conn.p; P.elem_sec = h->elem_sec.p; P.sec_props = h->sec_props.p;
  P.E = h->E; P.G = h->G; P.rho = h->rho;
  return P;
}

// operator choice of a Krylov solve: opts.op (FEMB_OP_*), overridable with FEMB_OPERATOR=bsr|ebe
bool ebe_available(const femb_handle* h) {
  return h->kind == Kind::Frame && h->pairs_dev_ok && !dist_active(h);
}

// row-block partition: the pair lists of the owned nodes are complete on every rank ("owner computes":
// the local mesh holds every element that touches an owned node), ghost nodes are only gathered from
bool ebe_available_dist(const femb_handle* h) {
  static int off = -1;
  if (off < 0) { const char* e = getenv("FEMB_OPERATOR"); off = (e && (e[0] == 'b' || e[0] == 'B')) ? 1 : 0; }
  return !off && h->kind == Kind::Frame && h->pairs_dev_ok && dist_active(h);
}

bool ebe_selected(const femb_handle* h, int op) {
  static int env = -1;
  if (env < 0) {
    const char* e = getenv("FEMB_OPERATOR");
    env = 0;
    if (e && (e[0] == 'b' || e[0] == 'B')) env = FEMB_OP_BSR;
    if (e && (e[0] == 'e' || e[0] == 'E')) env = FEMB_OP_EBE;
  }
  if (env) op = env;
  if (op == FEMB_OP_BSR) return false;
  return ebe_available(h);
}

constexpr int kEbeThreads = 128;
constexpr int kEbe1CtasPerSm = 4;   // 128 registers: 4 x 128 threads resident per SM
constexpr int kEbe4CtasPerSm = 3;   // 168 registers

// one thread per (node, element end) pair: the four expensive constants of its element record (ebe.cuh: PairAux)
__global__ void ebe_pair_aux_kernel(const FrameParams P, const int4* __restrict__ pair_rec, int64_t n_pairs,
                                    double4* __restrict__ aux) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs) return;
  const int4 rec = __ldg(pair_rec + p);
  const int a = (rec.w >> 24) & 1;
  const double* pn = P.xyz + 3 * (size_t)rec.x;
  const double* po = P.xyz + 3 * (size_t)rec.y;
  const double px = pn[0], py = pn[1], pz = pn[2], ox = po[0], oy = po[1], oz = po[2];
  const double* sp = P.sec_props + 8 * (size_t)(rec.w & 0xFFFFFF);
  FrameIn in;   // element direction: end 0 -> end 1, as in the operator
  in.dx = a ? px - ox : ox - px; in.dy = a ? py - oy : oy - py; in.dz = a ? pz - oz : oz - pz;
  in.A = sp[0]; in.Ix = sp[1]; in.Iy = sp[2]; in.J = sp[3]; in.ky = sp[4]; in.kz = sp[5];
  const PairAux x = pair_aux_from(P, in);
  aux[p] = make_double4(x.iL, x.iD, x.wz, x.wy);
}

int ebe_pair_aux(femb_handle* h) {
  const int64_t n_pairs = (int64_t)h->sym.pair_code.size();
  if (n_pairs == 0 || !h->pair_rec.p) return FEMB_OK;
  FEMB_CUDA(h, h->pair_aux.ensure((size_t)n_pairs * 4));
  ebe_pair_aux_kernel<<<(unsigned)((n_pairs + 255) / 256), 256, 0, h->stream>>>(
      ebe_params(h), reinterpret_cast<const int4*>(h->pair_rec.p), n_pairs, reinterpret_cast<double4*>(h->pair_aux.p));
  h->launches++;
  FEMB_CUDA(h, cudaGetLastError());
  return FEMB_OK;
}

double ebe_bytes(const femb_handle* h, int nb) {
  const Symbolic& S = h->sym;
  // pair + node records, coordinates, x read once, y written, BC mask
  return 16.0 * (double)S.pair_code.size() + 16.0 * h->n_nodes + 24.0 * h->n_nodes + 16.0 * nb * h->ndof +
         1.0 * h->ndof;
}

int ebe_grid(const femb_handle* h, int nb, int64_t n_nodes) {
  const int lanes_per_node = 2;   // nb = 1: T = 2;  nb = 4: two lanes with two vectors each
  const int npc = kEbeThreads / lanes_per_node;
  const int need = ((int)n_nodes + npc - 1) / npc;
  return std::max(1, std::min(need, h->num_sms * (nb == 1 ? kEbe1CtasPerSm : kEbe4CtasPerSm)));   // nb = 2, 4: 168 registers
}

// y = K_ff x (masked) or K x; nb = 1 (plain) or 4 (interleaved by right-hand side, x[g*4 + q]).
// dot_partials != null: (x_q, y_q) -> scal_out[q] through the ordered grid reduction (ticket = its
// ticket slot, done = early-exit flag, may be null).  link != null (nb = 1): linked PCG launch.
// p2p_dev / n_rows_nodes: row-block distributed launch over the first n_rows_nodes (owned) nodes.
int launch_ebe(femb_handle* h, const double* x, double* y, int nb, bool masked, double* dot_partials,
               double* scal_out, int* ticket, const int* done, const PcgLink* link, const void* p2p_dev,
               int64_t n_rows_nodes) {
  if (h->n_nodes <= 0) return FEMB_OK;
  const FrameParams P = ebe_params(h);
  const int pstride = h->num_sms * 8;
  const int4* pr = reinterpret_cast<const int4*>(h->pair_rec.p);
  const int4* nr = reinterpret_cast<const int4*>(h->pair_node_rec.p);
  const int n_nodes = (int)(n_rows_nodes >= 0 ? n_rows_nodes : h->n_nodes);
  if (n_nodes == 0) return FEMB_OK;
  const int grid = ebe_grid(h, nb, n_nodes);
  const PcgLink nolink = {};
  const P2PDev* p2p = reinterpret_cast<const P2PDev*>(p2p_dev);
  if (p2p && !(nb == 1 && masked && dot_partials && done && !link)) return fail(h, FEMB_ERR_ARG, "peer-memory operator launch: masked single vector with reduction");
#define EBE(NBT, NB, T, M, D, LK, MINB)                                                                      \
  frame_ebe_node_kernel<NBT, NB, T, M, D, LK, kEbeThreads, MINB><<<grid, kEbeThreads, 0, h->stream>>>(        \
      P, pr, nr, n_nodes, h->free_mask.p, x, y, dot_partials, pstride, scal_out, ticket, done, LK ? *link : nolink, p2p)
  if (link && !masked) return fail(h, FEMB_ERR_ARG, "linked operator launch is masked");
  if (nb == 1) {
    if (link) {
      EBE(1, 1, 2, true, true, true, kEbe1CtasPerSm);
    } else if (masked && dot_partials) EBE(1, 1, 2, true, true, false, kEbe1CtasPerSm);
    else if (masked) EBE(1, 1, 2, true, false, false, kEbe1CtasPerSm);
    else EBE(1, 1, 2, false, false, false, kEbe1CtasPerSm);
  } else if (nb == 4 && link) {
    EBE(4, 2, 1, true, true, true, kEbe4CtasPerSm);
  } else if (nb == 2 && link) {
    EBE(2, 2, 2, true, true, true, kEbe4CtasPerSm);
  } else if (nb == 4 && !link) {
    if (masked && dot_partials) EBE(4, 2, 1, true, true, false, kEbe4CtasPerSm);
    else if (masked) EBE(4, 2, 1, true, false, false, kEbe4CtasPerSm);
    else EBE(4, 2, 1, false, false, false, kEbe4CtasPerSm);
  } else if (nb == 2 && !link) {      // two interleaved vectors: both on one lane, two lanes split the pairs
    if (masked && dot_partials) EBE(2, 2, 2, true, true, false, kEbe4CtasPerSm);
    else if (masked) EBE(2, 2, 2, true, false, false, kEbe4CtasPerSm);
    else EBE(2, 2, 2, false, false, false, kEbe4CtasPerSm);
  } else {
    return fail(h, FEMB_ERR_ARG, "matrix-free operator: 1, 2 or 4 vectors");
  }
#undef EBE
  h->launches++;
  FEMB_CUDA(h, cudaGetLastError());
  return FEMB_OK;
}

}  // namespace femb
