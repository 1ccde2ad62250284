// Stress recovery: node-averaged axial + bending stress, BeamSolver.py:420-438.
// Gather form: one thread per node walks the node's incident elements in ascending element
// order (the diagonal block's contribution list), so the per-node sum has the same order as
// the reference's element loop and needs no atomics.
#include "common.cuh"
#include "elements.cuh"

namespace femb {

// f_local = k_ (R u_e) restricted to the rows the reference uses: 6 (axial force at end 2),
// 4,5 (moments at end 1), 10,11 (moments at end 2) — BeamSolver.py:431-434.
__device__ __forceinline__ double frame_end_stress(const FrameParams& P, uint32_t e, int end,
                                                   const double* __restrict__ u) {
  FrameRec R;
  frame_record(P, e, R);
  const int32_t na = P.conn[2 * e], nb = P.conn[2 * e + 1];
  double ul[12];  // R u_e : local displacements / rotations of both ends
#pragma unroll
  for (int blk = 0; blk < 4; ++blk) {
    const double* src = u + 6 * (size_t)(blk < 2 ? na : nb) + 3 * (blk & 1);
    const double g0 = src[0], g1 = src[1], g2 = src[2];
    ul[3 * blk + 0] = R.t[0] * g0 + R.t[1] * g1 + R.t[2] * g2;
    ul[3 * blk + 1] = R.n1[0] * g0 + R.n1[1] * g1 + R.n1[2] * g2;
    ul[3 * blk + 2] = R.n2[0] * g0 + R.n2[1] * g1 + R.n2[2] * g2;
  }
  // rows of the local stiffness (BeamSolver.py:655-660)
  const double f6 = -R.ax * ul[0] + R.ax * ul[6];
  const double f4 = -R.k12y * ul[2] + R.k22y * ul[4] + R.k12y * ul[8] + R.k23y * ul[10];
  const double f5 = R.k12z * ul[1] + R.k22z * ul[5] - R.k12z * ul[7] + R.k23z * ul[11];
  const double f10 = -R.k12y * ul[2] + R.k23y * ul[4] + R.k12y * ul[8] + R.k22y * ul[10];
  const double f11 = R.k12z * ul[1] + R.k23z * ul[5] - R.k12z * ul[7] + R.k22z * ul[11];
  const double* sp = P.sec_props + 8 * (size_t)P.elem_sec[e];
  const double A = sp[0], Ix = sp[1], Iy = sp[2], cy = sp[6], cz = sp[7];
  const double sa = A > 0.0 ? f6 / A : 0.0;
  const double my = (end == 0) ? f4 : f10, mz = (end == 0) ? f5 : f11;
  const double sb = fabs(Ix > 0.0 ? my * cz / Ix : 0.0) + fabs(Iy > 0.0 ? mz * cy / Iy : 0.0);
  return sa + sb;
}

__global__ void frame_stress_kernel(FrameParams P, const int32_t* __restrict__ diag_blk,
                                    const int32_t* __restrict__ contrib_ptr,
                                    const uint32_t* __restrict__ contrib, const double* __restrict__ u,
                                    double* __restrict__ sigma, int64_t n_nodes) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_nodes) return;
  const int b = diag_blk[i];
  const int c0 = contrib_ptr[b], c1 = contrib_ptr[b + 1];
  double s = 0.0;
  int cnt = 0;
  for (int c = c0; c < c1; ++c) {
    const uint32_t code = contrib[c];
    const int a = (code >> 1) & 1, bb = code & 1;
    if (a != bb) continue;  // degenerate self-loop element lists (a,b) too; count ends only
    s += frame_end_stress(P, code >> 2, a, u);
    ++cnt;
  }
  sigma[i] = cnt ? s / (double)cnt : 0.0;
}

int launch_frame_stress(femb_handle* h, const double* d_u, double* d_sigma) {
  FrameParams P;
  P.xyz = h->xyz.p; P.conn = h->conn.p; P.elem_sec = h->elem_sec.p; P.sec_props = h->sec_props.p;
  P.E = h->E; P.G = h->G; P.rho = h->rho;
  const int grid = (int)((h->n_nodes + 127) / 128);
  frame_stress_kernel<<<grid, 128, 0, h->stream>>>(P, h->diag_blk.p, h->contrib_ptr.p, h->contrib.p, d_u, d_sigma, h->n_nodes);
  h->launches++;
  FEMB_CUDA(h, cudaGetLastError());
  return FEMB_OK;
}

}  // namespace femb
