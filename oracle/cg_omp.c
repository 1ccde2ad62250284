/* TEST / BENCH INFRASTRUCTURE ONLY — never linked into or called by the product path.
 *
 * Multi-threaded C restatement of the CPU solve the oracle runs at scale: the Jacobi-preconditioned
 * conjugate gradients of scipy.sparse.linalg.cg on the CSR matrix K_ff that oracle/ref_sparse.py
 * assembles (the sparse stand-in for np.linalg.solve(k_ff, f_f), /root/reference/BeamSolver.py:417, and
 * spsolve(K[act, act], f[act]), /root/reference/ReactionSolver.py:199-201, which cannot run at 1M DOF).
 * scipy's own CG is single-threaded; this one uses every host core (OpenMP) so that
 *   - bench.py --impl reference can time the WHOLE 1M-DOF CPU solve instead of extrapolating a sample;
 *   - oracle/make_golden_large.py can produce full-size golden vectors in minutes.
 * Same recurrence as scipy's cg (x0 = 0; z = r / diag; alpha = (r,z)/(p,Ap); stop at ||r|| <= rtol ||b||).
 * Pinned against scipy.sparse.linalg.cg in tests/test_oracle.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int femb_oracle_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers: the reference arm of bench.py (rank 0 under torchrun at N > 1)
 * asks for the host's cores explicitly */
void femb_oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* y = A x, CSR, rows split across threads */
static void spmv(int64_t n, const int32_t* indptr, const int32_t* indices, const double* data, const double* x, double* y) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    double s = 0.0;
    for (int32_t k = indptr[i]; k < indptr[i + 1]; ++k) s += data[k] * x[indices[k]];
    y[i] = s;
  }
}

/* returns 0 when converged, 1 when max_iter was reached, -1 on breakdown / allocation failure */
int femb_oracle_pcg_jacobi(int64_t n, const int32_t* indptr, const int32_t* indices, const double* data,
                           const double* b, double* x, double rtol, int32_t max_iter, int32_t* iters, double* rel_res) {
  double* r = (double*)malloc(sizeof(double) * (size_t)n);
  double* z = (double*)malloc(sizeof(double) * (size_t)n);
  double* p = (double*)malloc(sizeof(double) * (size_t)n);
  double* q = (double*)malloc(sizeof(double) * (size_t)n);
  double* dinv = (double*)malloc(sizeof(double) * (size_t)n);
  if (!r || !z || !p || !q || !dinv) { free(r); free(z); free(p); free(q); free(dinv); return -1; }
  double bb = 0.0, rz = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : bb, rz)
  for (int64_t i = 0; i < n; ++i) {
    double d = 0.0;
    for (int32_t k = indptr[i]; k < indptr[i + 1]; ++k)
      if (indices[k] == i) d = data[k];
    dinv[i] = d != 0.0 ? 1.0 / d : 1.0;
    x[i] = 0.0;
    r[i] = b[i];
    z[i] = dinv[i] * r[i];
    p[i] = z[i];
    bb += b[i] * b[i];
    rz += r[i] * z[i];
  }
  int rc = 1;
  int32_t it = 0;
  double rr = bb;
  if (bb == 0.0) rc = 0;
  while (rc == 1 && it < max_iter) {
    spmv(n, indptr, indices, data, p, q);
    double pq = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : pq)
    for (int64_t i = 0; i < n; ++i) pq += p[i] * q[i];
    if (!(pq > 0.0)) { rc = -1; break; }
    const double alpha = rz / pq;
    double rz_new = 0.0;
    rr = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : rr, rz_new)
    for (int64_t i = 0; i < n; ++i) {
      x[i] += alpha * p[i];
      r[i] -= alpha * q[i];
      z[i] = dinv[i] * r[i];
      rr += r[i] * r[i];
      rz_new += r[i] * z[i];
    }
    ++it;
    if (sqrt(rr) <= rtol * sqrt(bb)) { rc = 0; break; }
    const double beta = rz_new / rz;
    rz = rz_new;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) p[i] = z[i] + beta * p[i];
  }
  if (iters) *iters = it;
  if (rel_res) *rel_res = bb > 0.0 ? sqrt(rr / bb) : 0.0;
  free(r); free(z); free(p); free(q); free(dinv);
  return rc;
}
