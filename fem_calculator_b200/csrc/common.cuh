// Shared declarations for libfemb200: handle layout, device buffers, error plumbing.
#pragma once

#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "../../include/femb200.h"

namespace femb {

constexpr int kNumSMsB200 = 148;

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  bool owned = true;   // false: p points into another allocation (adopt)
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p && owned) cudaFree(p);
    p = nullptr;
    n = 0;
    owned = true;
  }
  // view of `count` elements inside memory owned by someone else
  void adopt(T* ptr, size_t count) {
    release();
    p = ptr;
    n = count;
    owned = false;
  }
  cudaError_t alloc(size_t count) {
    if (count == n && p) return cudaSuccess;
    release();
    if (count == 0) return cudaSuccess;
    cudaError_t e = cudaMalloc(&p, count * sizeof(T));
    if (e == cudaSuccess) n = count;
    return e;
  }
  // grow-only variant: keeps a larger existing allocation
  cudaError_t ensure(size_t count) { return (p && count <= n) ? cudaSuccess : alloc(count); }
  size_t bytes() const { return n * sizeof(T); }
};

// Host-side symbolic structure (block-CSR pattern + contribution lists + tiles).
struct Symbolic {
  int32_t bs = 0;              // block size (6 frame, 3 tet10)
  int32_t nper = 0;            // nodes per element
  int64_t n_nodes = 0, n_elem = 0;
  int64_t nnzb = 0;            // number of bs x bs blocks
  int64_t n_contrib = 0;
  std::vector<int32_t> rowptr;       // (n_nodes+1)
  std::vector<int32_t> colidx;       // (nnzb) sorted within each row
  std::vector<int32_t> blk_row;      // (nnzb) row node of each block
  std::vector<int32_t> diag_blk;     // (n_nodes) index of block (i,i)
  std::vector<int32_t> contrib_ptr;  // (nnzb+1)
  std::vector<uint32_t> contrib;     // (n_contrib) e*nper^2 + a*nper + b, ascending per block
  std::vector<int32_t> contrib_blk;  // (n_contrib) owning block
  std::vector<int32_t> tile_ptr;     // (n_tiles+1) node ranges of assembly tiles
  int32_t tile_max_blocks = 0;       // capacity limits the tiles were packed for
  int32_t tile_max_contrib = 0;
  // frame "pair" view: one entry per (node, incident element end), element-ascending per node
  std::vector<int32_t> pair_ptr;     // (n_nodes+1)
  std::vector<uint32_t> pair_code;   // (n_pairs) e*2 + a
  std::vector<int32_t> pair_blk;     // (n_pairs) block (node, other end)
  std::vector<int32_t> pair_tile_ptr;  // (n_pair_tiles+1) node ranges of the pair-kernel tiles
  bool pairs_ok = false;             // false: self-loop / duplicate members / hub node -> generic kernel
  bool is_chain = false;             // path graph(s): block-tridiagonal after chain ordering
  std::vector<int32_t> chain_order;  // (n_nodes) node visited at chain position k
};

// Host-side symbolic part of the two-level preconditioner (coarse.cpp)
struct CoarseSym {
  int32_t n_agg = 0, max_nbr = 0;
  std::vector<int32_t> node_agg;   // (n_nodes) aggregate of each node
  std::vector<int32_t> agg_ptr;    // (n_agg+1)
  std::vector<int32_t> agg_nodes;  // (n_nodes) node lists, ascending node id per aggregate
  std::vector<int32_t> nbr_ptr;    // (n_agg+1)
  std::vector<int32_t> nbr;        // neighbour aggregates (self included), ascending per aggregate
  std::vector<int32_t> blk_slot;   // (nnzb) slot of block (i, j) in the neighbour list of agg(i)
};
void build_aggregates(int64_t n_nodes, const double* xyz, int n_parts, std::vector<int32_t>& agg);
void build_coarse_symbolic(const Symbolic& S, const std::vector<int32_t>& agg, int n_agg, CoarseSym& C);
void build_member_lines(int64_t n_nodes, int64_t n_elem, const int32_t* conn, const double* xyz, double cos_tol,
                        int min_nodes, std::vector<int32_t>& line_ptr, std::vector<int32_t>& line_nodes,
                        std::vector<double>& line_dir, std::vector<int32_t>& line_family);

void build_symbolic(int64_t n_nodes, int64_t n_elem, int nper, int bs, const int32_t* conn,
                    int tile_max_blocks, int tile_max_contrib, Symbolic& out);

// Host-side symbolic part of the line preconditioner (coarse.cpp, lines.cu): member lines sorted by
// (family, bundle), their node entries in order along the line, and the bundles (groups of
// neighbouring lines of one family) whose axial translations span the coarse space.
constexpr int kLnMaxFam = 3;       // families = classes of node-disjoint lines (dominant direction component)
constexpr int kLnMaxLen = 128;     // entries per line (longer member lines are cut): 4 per lane of the line's warp
struct LineSym {
  int32_t n_lines = 0, n_coarse = 0;
  int64_t n_entries = 0;
  double coverage = 0.0;             // fraction of nodes that lie on at least one line
  int32_t fam_off[kLnMaxFam + 1] = {0, 0, 0, 0};   // coarse index range of each family
  std::vector<int32_t> line_ptr;     // (n_lines+1) entry ranges; lines sorted by (family, bundle)
  std::vector<int32_t> line_bundle;  // (n_lines) coarse index (global over the families)
  std::vector<int32_t> bundle_ptr;   // (n_coarse+1) line ranges of the bundles
  std::vector<int32_t> ent_node;     // (n_entries)
  std::vector<int32_t> ent_blk_diag; // (n_entries) block (i, i) of K
  std::vector<int32_t> ent_blk_next; // (n_entries) block (i, next node of the line) or -1 at the line's end
  std::vector<int32_t> node_bundle;  // (kLnMaxFam, n_nodes) coarse index of the node's line in family f, -1: none
  std::vector<int32_t> node_ent;     // (kLnMaxFam, n_nodes) entry of the node in family f, -1: none
  std::vector<int32_t> node_line;    // (kLnMaxFam, n_nodes) line of the node in family f, -1: none
  std::vector<int32_t> bundle_ids;   // (number of bundle_ptr ranges) coarse index of each range; empty = identity.  Filled on a
                                     // row-block partition, where a rank only holds the bundles that cross its slab
  std::vector<double> node_dir;      // (kLnMaxFam, n_nodes, 3) unit line direction per node when the caller provides it
                                     // (row-block partition: the GLOBAL line's direction, identical on all ranks); empty = the
                                     // device computes end-to-end directions from the current coordinates
};
void build_line_symbolic(const Symbolic& S, const int32_t* conn, const double* xyz, int target_per_family, LineSym& out);
// Line tables of one rank of a row-block partition from the GLOBAL symbolic phase restricted to the rank's local
// nodes (owned first, then ghosts): lines are cut where they leave the owned range; bundles keep their global index.
void build_line_symbolic_local(const Symbolic& S, int64_t n_owned, int32_t n_coarse, const int32_t* fam_off,
                               const int32_t* node_bundle, const int32_t* node_line, const int32_t* node_pos,
                               const double* node_dir, LineSym& out);

enum class Kind { None, Frame, Tet10 };

}  // namespace femb

struct femb_handle {
  int device = 0;
  int num_sms = femb::kNumSMsB200;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;  // femb_timer
  std::vector<cudaEvent_t> ev_pool;               // opts.profile timing events (persist across solves)
  std::string err;
  femb::Kind kind = femb::Kind::None;
  int bs = 0;
  int64_t n_nodes = 0, n_elem = 0, ndof = 0;
  int64_t launches = 0;  // kernel launches since the counter was last reset

  // mesh (device)
  femb::DevBuf<double> xyz;        // (n_nodes,3)
  femb::DevBuf<int32_t> conn;      // (n_elem,nper)
  femb::DevBuf<int32_t> elem_sec;  // (n_elem)
  femb::DevBuf<double> sec_props;  // (n_sec,8)
  int32_t n_sec = 0;
  double E = 0, G = 0, rho = 0, nu = 0;
  std::vector<int32_t> h_conn;     // host copy for the symbolic phase
  std::vector<int32_t> h_elem_sec; // host copy (frame): baked into the pair records

  // pattern (device)
  femb::Symbolic sym;
  bool have_symbolic = false, assembled = false;
  femb::DevBuf<int32_t> rowptr, colidx, blk_row, diag_blk, contrib_ptr, contrib_blk, tile_ptr;
  femb::DevBuf<uint32_t> contrib;
  femb::DevBuf<int32_t> pair_rec;   // (n_pairs,4) {node, other node, block, sec | end<<24 | pos<<25}
  femb::DevBuf<double> pair_aux;    // (n_pairs,4) {1/L, 1/sqrt(cx^2+cy^2), w_z, w_y}: persistent PCG kernel (ebe.cuh), per assembled K
  femb::DevBuf<int32_t> pair_node_rec;  // (n_nodes,4) {first pair, pair count, diagonal block, 0}
  femb::DevBuf<int32_t> pair_tiles;     // (n_tiles,4) {first node, node count, first pair, pair count}
  bool pairs_dev_ok = false;        // pair records uploaded (frame fast path usable)
  femb::DevBuf<double> Kvals;      // (nnzb, bs, bs)
  femb::DevBuf<double> Mdiag;      // (n_nodes, bs, bs) frame only
  femb::DevBuf<double> tet_grad;   // (n_elem,4,32) Tet10 Gauss-point records of the two-stage assembly
  femb::DevBuf<unsigned long long> counters;  // device scalars: [0] skipped gauss points
  int64_t neg_detj = 0;

  // BC + vectors (device)
  bool have_bc = false;
  int64_t n_fixed = 0;
  std::vector<int64_t> h_fixed;      // host copy of the fixed DOF list (owned-count for the distributed modal solve)
  femb::DevBuf<uint8_t> free_mask;  // (ndof) 1 = free DOF
  femb::DevBuf<double> f, u0;       // load vector, prescribed values
  femb::DevBuf<double> b, x, r, z, p, q, s;
  femb::DevBuf<double> Dinv;        // block-Jacobi inverse (n_nodes,bs,bs) or Jacobi (ndof)
  femb::DevBuf<double> fpartials;   // published partial sums of the linked reductions (pcg_common.cuh: PcgLink)
  femb::DevBuf<double> mx, mr, mp, mq, mpartials, mscal;   // multi-RHS PCG (interleaved by right-hand side)
  femb::DevBuf<int32_t> mflags;
  femb::DevBuf<double> partials;    // reduction scratch
  femb::DevBuf<double> scal;        // device scalars for PCG
  femb::DevBuf<int32_t> flags;      // [0] done, [1] iterations, [2] ticket counters...
  bool have_solution = false;
  femb::DevBuf<double> stress_u, stress_sigma;   // femb_frame_stress staging

  // persistent direct-solver factors (invalidated by femb_assemble / femb_set_bc)
  bool chain_factored = false, dense_factored = false;
  femb::DevBuf<int32_t> chain_order;
  femb::DevBuf<double> chainG, chainW;   // (n_nodes,36) each, chain positions
  femb::DevBuf<double> denseL;           // (ndof,ndof) Cholesky factor of the masked operator

  // two-level preconditioner (twolevel.cu): aggregates + rigid-body coarse space
  bool coarse_sym_ok = false;     // aggregate tables on the device match the current topology
  bool coarse_num_ok = false;     // coarse_inv matches the current K and BC mask
  bool coarse_failed = false;     // the Galerkin matrix of the current K / BC could not be factored
  int32_t coarse_n_agg = 0, coarse_max_nbr = 0;
  int64_t coarse_n = 0, coarse_n_pad = 0;   // 6 * n_agg and its multiple-of-64 padding
  femb::DevBuf<int32_t> agg_ptr, agg_nodes, agg_nbr_ptr, agg_nbr, blk_slot;
  femb::DevBuf<double> agg_centroid;        // (n_agg,3)
  femb::DevBuf<double> coarse_aug;          // (2 n_pad)^2 work matrix of the inversion
  femb::DevBuf<double> coarse_inv;          // (n_pad, n_pad) inverse of the Galerkin matrix
  femb::DevBuf<double> coarse_r;            // (4 * n_pad) restricted residual(s)
  femb::DevBuf<double> coarse_scratch;      // chunk partial sums of the Galerkin assembly
  int64_t coarse_scratch_per_agg = 0;

  // line preconditioner (lines.cu): member lines, per-line tridiagonal factors, bundle coarse spaces
  bool line_sym_ok = false;       // line tables on the device match the current topology
  bool ln_mask_ok = false;        // row-block partition: the bundles' owner masks are known
  bool line_num_ok = false;       // factors and inverses match the current K and BC mask
  bool line_failed = false;       // a bundle Galerkin matrix of the current K / BC could not be factored
  femb::LineSym line_sym;         // host copy (the per-entry tables are dropped after the upload)
  bool line_dist = false;         // tables of a row-block partition (femb_dist_set_lines): pieces inside the slab, global bundles
  int32_t ln_fam_pad[femb::kLnMaxFam] = {0, 0, 0};
  int64_t ln_inv_off[femb::kLnMaxFam] = {0, 0, 0};
  int32_t ln_range_off[femb::kLnMaxFam + 1] = {0, 0, 0, 0};
  int32_t ln_n_ranges = 0, ln_max_len = 0;
  int32_t ln_grp_count = 0;       // line groups the current ln_grp_ptr / ln_grp_lines were built for
  femb::DevBuf<int32_t> ln_line_ptr, ln_line_bundle, ln_bundle_ptr, ln_ent_node, ln_ent_blk_diag, ln_ent_blk_next, ln_node_bundle,
      ln_bundle_ids, ln_bundle_cnt, ln_rank_mask, ln_ent_of, ln_grp_ptr, ln_grp_lines;
  femb::DevBuf<double> ln_ent_w, ln_node_w, ln_fac, ln_ae, ln_yle, ln_rb, ln_rbt, ln_yb, ln_inv, ln_gal, ln_node_dir, ln_line_sum;
  // line preconditioner setup: one stream, work matrix and Galerkin buffer per family (the three inversions overlap)
  cudaStream_t ln_stream[femb::kLnMaxFam] = {nullptr, nullptr, nullptr};
  cudaEvent_t ln_ev_start = nullptr, ln_ev_done[femb::kLnMaxFam] = {nullptr, nullptr, nullptr};
  femb::DevBuf<double> ln_aug_f[femb::kLnMaxFam], ln_gal_f[femb::kLnMaxFam];
  femb::DevBuf<int> ln_status;
  femb::DevBuf<double> vec_pool;                 // x | r | z | p | q | s (setup_bc_vectors)
  femb::DevBuf<unsigned long long> mega_state;   // persistent PCG kernel: grid barrier words + per-phase clocks

  // row-block distributed solve (dist.cu): this rank owns the first n_owned_nodes local nodes
  void* nccl_comm = nullptr;
  int dist_rank = 0, dist_world = 1;
  int64_t n_owned_nodes = 0;
  std::vector<int32_t> dist_nbr;                       // neighbour ranks
  std::vector<int64_t> dist_send_ptr, dist_recv_start, dist_recv_count;   // per neighbour, in nodes
  femb::DevBuf<int32_t> dist_send_nodes;               // local ids of the owned nodes each neighbour needs
  femb::DevBuf<double> dist_send_buf, dist_red;
  femb::DevBuf<uint8_t> dist_bnd_flag;                 // (n_nodes) 1 = owned block row that reads a ghost column
  femb::DevBuf<int32_t> dist_bnd_nodes;                // those rows, ascending
  int64_t dist_n_bnd = 0;
  void* nccl_comm_halo = nullptr;                      // second communicator: halo traffic on its own stream
  cudaStream_t halo_stream = nullptr;
  cudaEvent_t ev_vec = nullptr, ev_halo = nullptr;

  // peer-memory exchange (CUDA IPC over NVLink): mapped peers, mailboxes, sequence base
  femb::DevBuf<char> p2p_comm;
  void* p2p_dev = nullptr;                 // host copy of the P2PDev kernel argument
  const double* p2p_z_exported = nullptr;  // the z vector whose IPC handle the peers hold
  long long* p2p_base_dev = nullptr;
  int* p2p_ticket = nullptr;
  long long p2p_seq_base = 0;
  std::vector<void*> p2p_mapped;
  femb::DevBuf<char> p2p_dev_copy;         // device-resident P2PDev for the fused kernels
  femb::DevBuf<int32_t> p2p_send_slot, p2p_extra, p2p_bnd_nodes, p2p_bnd_dst_ptr, p2p_bnd_dst;

  // batched chain solve (direct.cu): persistent device buffers and the host ranges page-locked in place
  femb::DevBuf<double> batch_f, batch_u, batch_W, batch_z;
  std::vector<std::pair<void*, size_t>> batch_pinned;

  void* pinned = nullptr;           // small pinned staging area
  size_t pinned_bytes = 0;
};

namespace femb {

inline int fail(femb_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  return code;
}

#define FEMB_CUDA(h, expr)                                                              \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      return femb::fail((h), _e == cudaErrorMemoryAllocation ? FEMB_ERR_NOMEM : FEMB_ERR_CUDA, \
                        std::string(#expr) + ": " + cudaGetErrorString(_e));            \
    }                                                                                   \
  } while (0)

// process-wide host<->device traffic counters (bench.py reports them per step)
extern long long g_h2d_bytes, g_d2h_bytes;

template <typename T>
inline cudaError_t upload(DevBuf<T>& d, const T* src, size_t count, cudaStream_t s) {
  cudaError_t e = d.alloc(count);
  if (e != cudaSuccess || count == 0) return e;
  g_h2d_bytes += (long long)(count * sizeof(T));
  return cudaMemcpyAsync(d.p, src, count * sizeof(T), cudaMemcpyHostToDevice, s);
}

inline cudaError_t download(void* dst, const void* src, size_t bytes, cudaStream_t s) {
  g_d2h_bytes += (long long)bytes;
  return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s);
}

template <typename T>
inline cudaError_t upload(DevBuf<T>& d, const std::vector<T>& v, cudaStream_t s) {
  return upload(d, v.data(), v.size(), s);
}

// ---- kernel launch wrappers implemented in the .cu files ---------------------------------
int launch_frame_elements(femb_handle* h, double* d_ke, double* d_me);
int launch_tet10_elements(femb_handle* h, double* d_ke);
int launch_assemble(femb_handle* h);
int launch_expand_csr(femb_handle* h, int which, int32_t* d_indptr, int32_t* d_indices, double* d_vals);
int run_pcg(femb_handle* h, const femb_solve_opts& o, femb_stats* st);
int launch_spmv(femb_handle* h, const double* x, double* y, bool masked, double* dot_partials);
int launch_spmv_rows(femb_handle* h, const double* x, double* y, int64_t n, bool masked, double* dot_partials,
                     double* scal_out, const uint8_t* skip_node = nullptr, const int32_t* node_list = nullptr,
                     const void* p2p_dev = nullptr);
// matrix-free frame operator (ebe.cu)
bool ebe_available(const femb_handle* h);
bool ebe_selected(const femb_handle* h, int op);
double ebe_bytes(const femb_handle* h, int nb);
struct PcgLink;
int launch_ebe(femb_handle* h, const double* x, double* y, int nb, bool masked, double* dot_partials,
               double* scal_out, int* ticket, const int* done, const PcgLink* link, const void* p2p_dev = nullptr,
               int64_t n_rows_nodes = -1);
int ebe_grid(const femb_handle* h, int nb, int64_t n_nodes);
bool ebe_available_dist(const femb_handle* h);
int setup_precond_public(femb_handle* h, int mode);
int launch_reactions(femb_handle* h, bool minus_f, double* d_out);
int setup_bc_vectors(femb_handle* h);
int launch_frame_stress(femb_handle* h, const double* d_u, double* d_sigma);
int run_chain_solve(femb_handle* h, femb_stats* st);
int run_dense_solve(femb_handle* h, femb_stats* st);
int chain_factor(femb_handle* h);
int chain_apply(femb_handle* h, const double* d_b, double* d_x, int nrhs, int64_t ld);
int dense_factor(femb_handle* h);
int dense_apply(femb_handle* h, const double* d_b, double* d_x, int nrhs, int64_t ld);
int pcg_solve_rhs(femb_handle* h, const femb_solve_opts& o, const double* d_b, femb_stats* st);
// two-level preconditioner (twolevel.cu, direct.cu)
bool twolevel_applicable(const femb_handle* h, const femb_solve_opts& o);
int pcg_twolevel(femb_handle* h, const femb_solve_opts& o, const double* d_b, femb_stats* st);
int coarse_invert(femb_handle* h, double* aug, int64_t n_pad, double* inv, bool* ok);
// line preconditioner (lines.cu)
bool lines_applicable(femb_handle* h, const femb_solve_opts& o);
int pcg_lines(femb_handle* h, const femb_solve_opts& o, const double* d_b, femb_stats* st);
int dist_set_lines(femb_handle* h, int32_t n_coarse, const int32_t* fam_off, const int32_t* node_bundle,
                   const int32_t* node_line, const int32_t* node_pos, const double* node_dir);
bool dist_lines_applicable(femb_handle* h, const femb_solve_opts& o, bool fused_p2p);
int ebe_pair_aux(femb_handle* h);
int coarse_invert_async(femb_handle* h, cudaStream_t stream, double* aug, int64_t n_pad, double* inv, int* status_dev);
int dist_lines_setup(femb_handle* h);
int dist_lines_solve(femb_handle* h, const femb_solve_opts& o, const double* d_b, femb_stats* st);
double lines_iteration_bytes(const femb_handle* h);
int pcg_solve_multi(femb_handle* h, const femb_solve_opts& o, const double* d_B, int64_t ldb, int nb, double* d_X,
                    int64_t ldx, femb_stats* st);
int run_modal(femb_handle* h, const femb_eig_opts& o, double* lambda, double* phi, int32_t* n_found,
              femb_stats* st);
int run_batch_chain(femb_handle* h, int64_t n_models, int64_t n_elem, const double* xyz,
                    const double* sec_props, double E, double G, const uint8_t* fixed_mask,
                    const double* f, double* u, femb_stats* st);
int time_kernel(femb_handle* h, int which, int warm, int reps, double* ms, double* bytes);
bool dist_active(const femb_handle* h);
int dist_allreduce(femb_handle* h, double* d_buf, int count);
int dist_spmv_masked(femb_handle* h, double* x, double* y);
int dist_solve_rhs(femb_handle* h, const femb_solve_opts& o, const double* d_b, femb_stats* st);

}  // namespace femb
