/* femb200 — C ABI of the B200-native FEM hot path (libfemb200.so).
 *
 * Drop-in boundary for the numerical core of euler8511/FEM-calculator.  The reference has
 * no FFI of its own: its "interface" for this path is a set of Python methods on two
 * classes.  Each entry point below names the reference code it replaces (file:line into
 * /root/reference); fem_calculator_b200/compat.py binds them with ctypes and re-creates
 * the reference's Python signatures on top.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; femb_last_error(h) gives the text;
 *   - all array arguments are caller-owned, C-contiguous HOST pointers (the library copies
 *     to / from HBM on the handle's stream); a NULL output pointer means "leave it on the
 *     device, do not copy back";
 *   - indices: node / element / DOF numbers are 0-based; DOF = bs*node + c with bs = 6
 *     (frame: ux,uy,uz,rx,ry,rz — BeamSolver.py:354,360,390-393) or bs = 3 (Tet10 —
 *     ReactionSolver.py:148);
 *   - a handle owns one CUDA device + one stream and is not thread-safe; different
 *     handles are independent; there is no hidden global state and NO CPU fallback: every
 *     compute entry point fails with FEMB_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef FEMB200_H
#define FEMB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct femb_handle femb_handle;

enum {
  FEMB_OK = 0,
  FEMB_ERR_ARG = -1,       /* bad argument / call order */
  FEMB_ERR_CUDA = -2,      /* CUDA runtime error or no usable device */
  FEMB_ERR_NOT_CONVERGED = -3,
  FEMB_ERR_SINGULAR = -4,  /* non-positive pivot / breakdown: K_ff not SPD */
  FEMB_ERR_NOMEM = -5
};

/* which matrix femb_get_csr exports */
enum { FEMB_MAT_K = 0, FEMB_MAT_M = 1 };

/* femb_solve_opts.method */
enum {
  FEMB_SOLVER_AUTO = 0,    /* chain -> block-tridiagonal; tiny -> dense Cholesky; else PCG */
  FEMB_SOLVER_PCG = 1,     /* preconditioned conjugate gradients on the BSR operator */
  FEMB_SOLVER_CHAIN = 2,   /* direct block-tridiagonal Cholesky (path graphs only) */
  FEMB_SOLVER_DENSE = 3    /* dense FP64 blocked Cholesky of K_ff, trailing update on DMMA (<= 16384 DOF;
                              AUTO picks it up to 2048 DOF) */
};

/* femb_solve_opts.precond */
enum { FEMB_PRECOND_NONE = 0, FEMB_PRECOND_JACOBI = 1, FEMB_PRECOND_BLOCK_JACOBI = 2,
       FEMB_PRECOND_TWO_LEVEL = 3 /* Jacobi + coarse-space correction: the nodes are grouped into aggregates
                                     (recursive coordinate bisection) and the six rigid-body modes of every
                                     aggregate span a coarse space whose Galerkin matrix is inverted explicitly
                                     (DMMA Cholesky kernels); M^-1 = D^-1 + P (P^T K_ff P)^-1 P^T.  Frames on one
                                     GPU with the matrix-free operator (csrc/twolevel.cu): 5x fewer iterations
                                     than Jacobi at 1M DOF.  Where it does not apply (Tet10, BSR operator, the
                                     row-block distributed solve) the solve runs with JACOBI;
                                     femb_stats.coarse_dim tells which one ran.                              */,
       FEMB_PRECOND_AUTO = 4      /* frames with >= 50,000 nodes and the matrix-free operator: LINES when at least half
                                     of the nodes lie on member lines, else TWO_LEVEL; everything else JACOBI (below
                                     50,000 nodes the coarse setup costs more than it saves)                     */,
       FEMB_PRECOND_LINES = 5     /* Jacobi + exact solves along member lines + bundle coarse space (csrc/lines.cu): the
                                     axial DOFs of every maximal chain of collinear members are solved exactly (one
                                     tridiagonal system per line, factored per assembled K), and one axial translation
                                     per bundle of neighbouring parallel lines spans a coarse space inverted explicitly
                                     (DMMA Cholesky kernels):
                                     M^-1 = D^-1 + sum_lines Q (Q^T K_ff Q)^-1 Q^T + sum_families P (P^T K_ff P)^-1 P^T.
                                     Frames on one GPU with the matrix-free operator: 199 iterations instead of 1,324
                                     (TWO_LEVEL) / 6,931 (JACOBI) at 1M DOF.  Falls back like TWO_LEVEL does.          */ };

/* femb_solve_opts.op / femb_eig_opts.op: how the Krylov loops apply K_ff.
 *   BSR  the assembled block-CSR matrix (HBM-bound: 8 B per stored value and product);
 *   EBE  matrix-free: every (node, element end) pair rebuilds its element record from the node
 *        coordinates and the section row and applies the closed form of R^T k R
 *        (BeamSolver.py:386-387) to the two end displacement vectors — frames whose members are
 *        unique (no duplicate members, node degree <= 127), single GPU;
 *   AUTO EBE where it applies, else BSR.
 * K is assembled either way (Jacobi diagonal, reactions K u - f, CSR export).               */
enum { FEMB_OP_AUTO = 0, FEMB_OP_BSR = 1, FEMB_OP_EBE = 2 };

typedef struct {
  int32_t method;        /* FEMB_SOLVER_*                         default AUTO            */
  int32_t precond;       /* FEMB_PRECOND_*                        default AUTO            */
  int32_t max_iter;      /* PCG iteration cap                     default 200000          */
  int32_t check_every;   /* host polls convergence every N its    default 50              */
  double rtol;           /* stop at ||r||_2 <= rtol*||b||_2       default 1e-12           */
  int32_t profile;       /* P>0: time every P-th SpMV launch with CUDA events              */
  int32_t op;            /* FEMB_OP_*                             default AUTO            */
} femb_solve_opts;

typedef struct {
  int32_t k;             /* number of lowest modes wanted                                  */
  int32_t block;         /* Krylov block size 1..4; 0 = 4 behind a factorisation (chain / dense), 2 behind
                            PCG (every solve costs thousands of iterations); use >= the multiplicity
                            of the wanted eigenvalues                                              */
  int32_t max_iter;      /* default 5000                                                   */
  int32_t op;            /* FEMB_OP_* of the inner shift-invert solves   default AUTO             */
  double rtol;           /* ||K phi - lambda M phi|| <= rtol*||K phi||   default 1e-8      */
  double lambda_min;     /* keep eigenvalues > lambda_min (BeamSolver.py:448)  default 1e-6 */
  int32_t precond;       /* preconditioner of the inner PCG solves: FEMB_PRECOND_TWO_LEVEL (one right-hand side
                            at a time), FEMB_PRECOND_AUTO (as for the static solve), anything else = JACOBI
                            (4-/2-vector lockstep PCG); FEMB_PRECOND_LINES as TWO_LEVEL; ignored behind a factorisation                   */
  int32_t reserved;
  double accept_rtol;    /* a solve that stagnates above rtol (ill-conditioned chains: cond(K) eps > rtol) is still returned
                            with FEMB_OK when its worst residual is <= accept_rtol (femb_stats.converged stays 0 and
                            rel_residual tells what was reached); 0 = strict: FEMB_ERR_NOT_CONVERGED above rtol   */
} femb_eig_opts;

typedef struct {
  int32_t method_used;
  int32_t iterations;
  int32_t converged;
  int32_t spmv_launches;      /* SpMV kernel launches in this call                         */
  int32_t kernel_launches;    /* all kernel launches of this library in this call          */
  int32_t spmv_timed;         /* SpMV launches bracketed by events (opts.profile); modal: restarts */
  int32_t op_used;            /* FEMB_OP_BSR or FEMB_OP_EBE for the Krylov loops of this call (0: none) */
  int32_t coarse_dim;         /* dimension of the coarse space of FEMB_PRECOND_TWO_LEVEL / LINES (0: plain Jacobi ran) */
  int32_t precond_used;       /* FEMB_PRECOND_* the PCG of this call ran with (0 when no PCG ran)                 */
  int32_t reserved;
  double rel_residual;        /* final ||r||/||b||                                         */
  double device_ms;           /* CUDA-event time of the whole call on the handle's stream  */
  double spmv_ms;             /* summed device time of the timed SpMV launches              */
  double update_ms;           /* summed device time of the vector-update launches that followed them */
} femb_stats;

int femb_version(void);
/* number of visible CUDA devices (0 when there is none; never fails) */
int femb_device_count(void);

int femb_create(int device, femb_handle** out);
void femb_destroy(femb_handle* h);
const char* femb_last_error(const femb_handle* h);

/* ---- frame (3-D Timoshenko / Euler-Bernoulli beam) path ------------------------------
 * Replaces the per-element loop of BeamAnalysisWindow.run_simulation, BeamSolver.py:364-393,
 * i.e. get_timoshenko_stiffness_matrix (:646-660), get_lumped_mass_matrix (:662-675),
 * the direction-cosine matrix (:378-384), R^T k R / R^T m R (:386-388) and the scatter
 * (:390-393).
 *   xyz       (n_nodes,3) float64  — mesh.points                       (:372)
 *   conn      (n_elem,2)  int64    — mesh.cells_dict['line']           (:364)
 *   elem_sec  (n_elem)    int32    — index into sec_props per element  (:365-371)
 *   sec_props (n_sec,8)   float64  — rows (A, I_x, I_y, J, kappa_y, kappa_z, c_y_max,
 *                                    c_z_max) = calculate_section_properties' record (:79)
 *   E, G, rho                      — :350-352 and the literal 7850 at :376            */
int femb_frame_set_mesh(femb_handle* h, int64_t n_nodes, int64_t n_elem, const double* xyz,
                        const int64_t* conn, const int32_t* elem_sec, int32_t n_sec,
                        const double* sec_props, double E, double G, double rho);

/* Parity export of the per-element global-axis matrices R^T k R and R^T m R
 * (BeamSolver.py:387-388): ke, me are (n_elem,12,12) float64, either may be NULL. */
int femb_frame_elements(femb_handle* h, double* ke, double* me);

/* ---- Tet10 path ----------------------------------------------------------------------
 * Replaces ForceAnalysis._create_material_matrix / _shape_funcs_tet10 /
 * assemble_stiffness_matrix, ReactionSolver.py:87-152.
 *   conn10 (n_elem,10) int64 in meshio/VTK node order (:102-110).
 * Parity constants kept: Gauss literals 0.58541020 / 0.13819660 (:120-123), w = 1/4
 * (:124), Gauss points with detJ <= 1e-12 skipped and counted (:133-135).              */
int femb_tet10_set_mesh(femb_handle* h, int64_t n_nodes, int64_t n_elem, const double* xyz,
                        const int64_t* conn10, double E, double nu);
/* ke: (n_elem,30,30) float64 parity export of Ke (ReactionSolver.py:128-146). */
int femb_tet10_elements(femb_handle* h, double* ke);
/* ForceAnalysis.negative_detJ_count (ReactionSolver.py:49,134) after femb_assemble. */
int64_t femb_tet10_negative_detj(const femb_handle* h);

/* ---- shared: assembly, BC, solve, modal ----------------------------------------------
 * femb_assemble: symbolic pattern on first call (block-CSR over node adjacency, columns
 * sorted, one diagonal block per node), then the fused element+assembly kernel:
 * deterministic, atomic-free, bit-reproducible run to run.  Replaces the scatter at
 * BeamSolver.py:390-393 (dense K, M) and ReactionSolver.py:148-151 (lil -> csr).        */
int femb_assemble(femb_handle* h);

/* CSR export (two calls: sizes, then data).  Scalar CSR expanded from the stored BSR:
 * indices sorted within rows, no duplicates, structural zeros kept.  which = FEMB_MAT_K
 * or FEMB_MAT_M (frame only; M is block-diagonal).                                      */
int femb_get_csr_size(femb_handle* h, int which, int64_t* n_rows, int64_t* nnz);
int femb_get_csr(femb_handle* h, int which, int32_t* indptr, int32_t* indices, double* vals);

/* Boundary conditions + load vector: replaces BeamSolver.py:395-416 /
 * ReactionSolver.py:174,194,199-200 (the group -> node -> DOF bookkeeping stays in the
 * Python shim).  fixed_dofs: sorted unique global DOFs (Up_node / fixed_dofs); f: (ndof)
 * nodal loads; u_prescribed: (ndof) or NULL (the reference always prescribes 0,
 * BeamSolver.py:413).  Elimination is applied on the device as a DOF mask on the
 * operator, which is algebraically K_ff with u[fixed] = prescribed.                    */
int femb_set_bc(femb_handle* h, int64_t n_fixed, const int64_t* fixed_dofs, const double* f,
                const double* u_prescribed);

/* Static solve K u = F + reaction recovery: replaces np.linalg.solve(k_ff, f_f)
 * (BeamSolver.py:417-418) and spsolve + K @ u (ReactionSolver.py:201-205).
 *   u          (ndof) out, fixed DOFs carry the prescribed value
 *   reactions  (ndof) out, K_full @ u - f   (minus_f != 0)   [frame convention]
 *                       or K_full @ u       (minus_f == 0)   [ReactionSolver.py:205]    */
int femb_solve_static(femb_handle* h, const femb_solve_opts* opts, int minus_f, double* u,
                      double* reactions, femb_stats* stats);

/* y = K x (masked == 0: the un-eliminated matrix, the product behind K @ u at ReactionSolver.py:205)
 * or y = K_ff x with identity rows on the fixed DOFs (masked != 0; x must be zero there), applied by
 * the operator the Krylov loops would use for `op` (FEMB_OP_*).  x, y: (ndof) host arrays.
 * Returns the operator actually used in *op_used (may be NULL).                                 */
int femb_apply_k(femb_handle* h, int op, int masked, const double* x, double* y, int32_t* op_used);

/* Lowest-k modes of K_ff phi = lambda M_ff phi: replaces inv(m_ff) @ k_ff + qr_algorithm
 * (BeamSolver.py:440-455,467-481).  lambda: (k) ascending, eigenvalues <= lambda_min
 * dropped (:448); phi: (ndof,k) column-major (phi[j*ndof + i]), M-normalised, zeros on
 * fixed DOFs (:453-455).  n_found receives the number of modes returned.  On a handle set up
 * with femb_dist_init / femb_dist_set_halo (row-block partition) every rank calls it, lambda is
 * identical on all ranks and phi holds the rank's OWNED rows only: (bs*n_owned_nodes, k).      */
int femb_modal(femb_handle* h, const femb_eig_opts* opts, double* lambda, double* phi,
               int32_t* n_found, femb_stats* stats);

/* Node-averaged axial + bending stress: replaces BeamSolver.py:420-438.  u: (ndof) or
 * NULL to use the last solution held on the device; sigma_node: (n_nodes).             */
int femb_frame_stress(femb_handle* h, const double* u, double* sigma_node);

/* ---- batched independent chain models (BASELINE config 4) ----------------------------
 * n_models cantilever/chain models of n_elem elements each (nodes 0..n_elem in chain
 * order along xyz), one section record per model, solved by a batched block-tridiagonal
 * factorisation (six lanes per model, one row of every 6x6 block each; five models per warp).  Equivalent to
 * n_models calls of run_simulation's static part (BeamSolver.py:360-418).  f and u are hundreds of MB at BASELINE
 * config 4: page-lock them with femb_host_register for PCIe-speed copies (optional; same results without).
 *   xyz        (n_nodes,3) shared node coordinates, n_nodes = n_elem+1
 *   sec_props  (n_models,8)
 *   fixed_mask (ndof) uint8, 1 = fixed DOF, shared by all models
 *   f          (n_models, ndof) loads;  u (n_models, ndof) out                          */
int femb_frame_batch_solve(femb_handle* h, int64_t n_models, int64_t n_elem, const double* xyz,
                           const double* sec_props, double E, double G,
                           const uint8_t* fixed_mask, const double* f, double* u,
                           femb_stats* stats);

/* Page-lock a caller-owned host buffer in place (cudaHostRegister) so that the library's copies to / from it run
 * asynchronously at PCIe speed.  The registration belongs to the CALLER: the buffer must stay allocated until
 * femb_host_unregister (or femb_destroy) — the library never keeps a registration of memory it was not told to keep,
 * because a freed and re-allocated address would silently alias the old pages.                                  */
int femb_host_register(femb_handle* h, void* ptr, int64_t bytes);
int femb_host_unregister(femb_handle* h, void* ptr);

/* ---- one large mesh across GPUs: row-block partition ------------------------------------------
 * One process (and one handle) per GPU.  Each rank passes femb_*_set_mesh its LOCAL mesh: the
 * owned nodes first (ascending global id: a coordinate-bisection box, a contiguous slab of the
 * node order, or any other set), then the ghost nodes its elements touch (grouped by owner rank,
 * ascending global id inside a group), and every element that
 * touches an owned node; femb_assemble / femb_set_bc work on that local mesh unchanged ("owner
 * computes": no assembly communication; owned rows equal the single-GPU rows bit for bit).
 * fem_calculator_b200/partition.py builds these inputs from the global arrays the reference
 * holds.  The solve exchanges the halo of the CG direction with ncclSend/ncclRecv and
 * all-reduces the three CG scalars once per iteration (NCCL, loaded with dlopen on first use).
 * Scales the static solve of BeamSolver.py:417 / ReactionSolver.py:201 beyond one GPU.
 *
 * femb_dist_unique_id: rank 0 creates the 128-byte NCCL id; the host broadcasts it by any
 *   means (torch.distributed, MPI, a file) and every rank calls femb_dist_init.
 * femb_dist_set_halo: n_owned_nodes = size of the owned prefix; for neighbour k (rank
 *   nbr_rank[k]) send_nodes[send_ptr[k]..send_ptr[k+1]) are the LOCAL ids of owned nodes it
 *   needs (ascending global id), and [recv_start[k], recv_start[k]+recv_count[k]) is the range
 *   of local ghost nodes it owns.
 * femb_dist_solve_static: u_owned / reactions_owned are (bs*n_owned_nodes) or NULL.          */
int femb_dist_unique_id(uint8_t* id128);
int femb_dist_init(femb_handle* h, int rank, int world, const uint8_t* id128);
void femb_dist_finalize(femb_handle* h);
int femb_dist_set_halo(femb_handle* h, int64_t n_owned_nodes, int32_t n_nbr, const int32_t* nbr_rank,
                       const int64_t* send_ptr, const int32_t* send_nodes, const int64_t* recv_start,
                       const int64_t* recv_count);
int femb_dist_solve_static(femb_handle* h, const femb_solve_opts* opts, int minus_f, double* u_owned,
                           double* reactions_owned, femb_stats* stats);
/* Optional peer-memory (NVLink) exchange for the iteration: after femb_set_bc + femb_dist_set_halo
 * every rank exports one CUDA-IPC handle (128-byte blob: the handle of ONE allocation holding its mailboxes, its CG
 * direction vector and its flag-in-data slots, plus two sizes the peers need to address it), the
 * host gathers the world*128 bytes and every rank imports them together with, per neighbour k, the
 * first local node index of THIS rank's nodes in neighbour k's ghost numbering.  The distributed
 * PCG then stores halo entries, reduction scalars and coarse residuals straight into the peers' memory from inside
 * its kernels (csrc/dist.cu, csrc/lines.cu) instead of calling ncclSend/ncclRecv + ncclAllReduce; without the import
 * it uses NCCL (Jacobi only).  Up to 8 ranks.                                                                    */
int femb_dist_p2p_export(femb_handle* h, uint8_t* handles128);
/* Line preconditioner on the partition (FEMB_PRECOND_LINES / AUTO in femb_dist_solve_static; needs the peer-memory
 * exchange): every rank runs femb_symbolic_line_bundles on the GLOBAL mesh and passes the rows of its LOCAL nodes
 * (owned, then ghosts): node_bundle / node_line / node_pos (3, n_local) and node_dir (3, n_local, 3), plus the global
 * n_coarse and fam_off (4).  Lines are cut where they leave the rank's owned set; a bundle's residual is the sum of the
 * partial residuals of the ranks that own a piece of it (stored by their line groups into every rank's flag-in-data
 * slots, added in rank order) and the bundle Galerkin matrices are summed over the ranks once per assembled K.
 * Call after femb_dist_set_halo.                                                                                  */
int femb_dist_set_lines(femb_handle* h, int32_t n_coarse, const int32_t* fam_off, const int32_t* node_bundle,
                        const int32_t* node_line, const int32_t* node_pos, const double* node_dir);
int femb_dist_p2p_import(femb_handle* h, const uint8_t* all_handles, const int64_t* peer_ghost_start);

/* ---- measurement hooks (bench.py) ------------------------------------------------------
 * Time `reps` back-to-back launches of one kernel with CUDA events on the handle's stream
 * (after `warm` untimed launches); *ms receives the mean per launch, *bytes the
 * algorithmic bytes of one launch.  which: 0 = BSR SpMV (masked
 * K_ff operator), 1 = fused element+assembly, 3 = matrix-free (EBE) operator, 4 = its 4-vector
 * form, 5 = EBE with the fused (x, y) reduction, 7 = dense blocked Cholesky (fill + factor; *bytes
 * then receives the FLOP count n^3/3), 9 = plain 16-byte read of the K values (streaming ceiling), 10 = FP64 FMA issue
 * ceiling (8 independent chains per thread; *bytes receives the FMA count of one launch), 11 = no launch: *bytes receives
 * the algorithmic bytes of one iteration of the persistent line-preconditioned PCG kernel. */
int femb_time_kernel(femb_handle* h, int which, int warm, int reps, double* ms, double* bytes);

/* CUDA-event stopwatch on the handle's stream: stop = 0 records the start (after draining the
 * stream), stop = 1 records the end and returns the elapsed device time in *ms. */
int femb_timer(femb_handle* h, int stop, double* ms);

/* Bytes this process has copied host->device / device->host through the library since the
 * last reset (bench.py's e2e accounting). */
void femb_io_bytes(int64_t* h2d, int64_t* d2h, int reset);

/* Host-only symbolic analysis (no GPU needed; exercised by the CPU test-suite):
 * block-CSR pattern of an element mesh.  Two calls: pass NULL outputs to get sizes.     */
int femb_symbolic_pattern(int64_t n_nodes, int64_t n_elem, int32_t nodes_per_elem,
                          const int64_t* conn, int64_t* n_blocks, int32_t* rowptr,
                          int32_t* colidx);

/* Host-only: the node aggregates of FEMB_PRECOND_TWO_LEVEL — proportional recursive coordinate
 * bisection of the n_nodes points into n_parts groups whose sizes differ by at most one.
 * agg_of_node: (n_nodes) out, values in [0, n_parts).                                      */
int femb_symbolic_aggregates(int64_t n_nodes, const double* xyz, int32_t n_parts, int32_t* agg_of_node);

/* Host-only: the whole symbolic phase of the two-level preconditioner for a frame mesh (conn: (n_elem,2)) —
 * aggregates, aggregate adjacency induced by the block pattern (nbr_ptr (n_parts+1), nbr: ascending, self
 * included) and, per stored block b = (i, j) of femb_symbolic_pattern, blk_slot[b] with
 * nbr[nbr_ptr[agg(i)] + blk_slot[b]] == agg(j).  Two calls: nbr == blk_slot == NULL returns *n_nbr.      */
int femb_symbolic_coarse(int64_t n_nodes, int64_t n_elem, const int64_t* conn, const double* xyz,
                         int32_t n_parts, int32_t* agg_of_node, int32_t* nbr_ptr, int64_t* n_nbr,
                         int32_t* nbr, int32_t* blk_slot);

/* Host-only: member lines of a frame mesh — maximal chains of members whose consecutive directions differ by
 * less than acos(cos_tol) (groundwork for the line coarse space).  Every element belongs
 * to exactly one chain; chains with fewer than min_nodes nodes are dropped.  line_nodes: node ids ordered
 * along each line (line_ptr: (n_lines+1)); line_dir: (n_lines,3) unit end-to-end direction; line_family:
 * index of its dominant component.  Two calls: line_nodes == NULL returns *n_lines and *n_line_nodes.     */
int femb_symbolic_lines(int64_t n_nodes, int64_t n_elem, const int64_t* conn, const double* xyz,
                        double cos_tol, int32_t min_nodes, int64_t* n_lines, int64_t* n_line_nodes,
                        int32_t* line_ptr, int32_t* line_nodes, double* line_dir, int32_t* line_family);

/* Host-only: symbolic phase of FEMB_PRECOND_LINES (csrc/lines.cu) — member lines (femb_symbolic_lines, cut to 128
 * nodes) are split into three families of node-disjoint lines (dominant direction component) and the lines of
 * a family are grouped into at most target_per_family bundles (recursive coordinate bisection of the line
 * midpoints).  node_bundle: (3, n_nodes) coarse index of the node's line per family, -1 = none; node_pos:
 * (3, n_nodes) index of the node in the sorted entry list (consecutive along a line, lines of a bundle
 * consecutive); fam_off: (4) coarse index range per family; coverage: fraction of nodes on at least one line;
 * node_line: (3, n_nodes) line index per family; node_dir: (3, n_nodes, 3) unit end-to-end direction of that line.
 * Any output pointer may be NULL.                                                                          */
int femb_symbolic_line_bundles(int64_t n_nodes, int64_t n_elem, const int64_t* conn, const double* xyz,
                               int32_t target_per_family, int32_t* node_bundle, int32_t* node_pos,
                               int32_t* fam_off, int64_t* n_lines, int64_t* n_entries, double* coverage,
                               int32_t* node_line, double* node_dir);

#ifdef __cplusplus
}
#endif
#endif /* FEMB200_H */
