// Per-element device math shared by the parity-export kernels, the fused assembly kernel
// and the stress kernel.  Nothing here touches global memory except the explicit loads in
// the *_load functions, so every consumer computes bit-identical element values.
#pragma once

#include <cstdint>

namespace femb {

// ------------------------------------------------------------------------------ frame
// Closed form of R^T k R and R^T m R (BeamSolver.py:386-388) in terms of the rows
// t, n1, n2 of the direction-cosine matrix lambda (BeamSolver.py:378-384) and the ten
// distinct stiffness magnitudes of get_timoshenko_stiffness_matrix (BeamSolver.py:646-660):
//   R^T k R [a][b] = | s_ab*(ax tt' + k11z n1n1' + k11y n2n2')     s_a*(k12z n1n2' - k12y n2n1') |
//                    | s_b*(k12z n2n1' - k12y n1n2')     (+-tor) tt' + kq_y n1n1' + kq_z n2n2'    |
// with s_ab = +1 (a==b) / -1, s_a = +1 for end 0 / -1 for end 1, kq = k22 (a==b) / k23.
struct FrameParams {
  const double* xyz;        // (n_nodes,3)
  const int32_t* conn;      // (n_elem,2)
  const int32_t* elem_sec;  // (n_elem)
  const double* sec_props;  // (n_sec,8): A, I_x, I_y, J, kappa_y, kappa_z, c_y_max, c_z_max
  double E, G, rho;
};

struct FrameRec {
  double t[3], n1[3], n2[3];
  double ax, tor, k11z, k12z, k22z, k23z, k11y, k12y, k22y, k23y;
  double mt, mrx, mry, mrz;
  double L;
};

// raw inputs of one element record: end-to-end vector and the section row
struct FrameIn {
  double dx, dy, dz;
  double A, Ix, Iy, J, ky, kz;
};

__device__ __forceinline__ void frame_load(const FrameParams& P, int32_t na, int32_t nb, int32_t sec, FrameIn& in) {
  const double* pa = P.xyz + 3 * (size_t)na;
  const double* pb = P.xyz + 3 * (size_t)nb;
  in.dx = __ldg(pb) - __ldg(pa); in.dy = __ldg(pb + 1) - __ldg(pa + 1); in.dz = __ldg(pb + 2) - __ldg(pa + 2);
  const double* sp = P.sec_props + 8 * (size_t)sec;
  in.A = __ldg(sp); in.Ix = __ldg(sp + 1); in.Iy = __ldg(sp + 2); in.J = __ldg(sp + 3);
  in.ky = __ldg(sp + 4); in.kz = __ldg(sp + 5);
}

__device__ __forceinline__ void frame_record_from(const FrameParams& P, const FrameIn& in, FrameRec& r);

// element record from its two end nodes (na -> nb) and section row
__device__ __forceinline__ void frame_record_nodes(const FrameParams& P, int32_t na, int32_t nb, int32_t sec,
                                                   FrameRec& r) {
  FrameIn in;
  frame_load(P, na, nb, sec, in);
  frame_record_from(P, in, r);
}

__device__ __forceinline__ void frame_record(const FrameParams& P, uint32_t e, FrameRec& r) {
  frame_record_nodes(P, P.conn[2 * e], P.conn[2 * e + 1], P.elem_sec[e], r);
}

__device__ __forceinline__ void frame_record_from(const FrameParams& P, const FrameIn& in, FrameRec& r) {
  const double dx = in.dx, dy = in.dy, dz = in.dz;
  const double A = in.A, Ix = in.Ix, Iy = in.Iy, J = in.J, ky = in.ky, kz = in.kz;
  const double L2 = dx * dx + dy * dy + dz * dz;
  // 1/L through rsqrt (one MUFU + Newton) instead of sqrt followed by a division; the few-ulp
  // difference to numpy's norm/divide is far inside the 1e-14 parity band.
  const double iL = rsqrt(L2);
  const double L = L2 * iL;                        // BeamSolver.py:373
  const double cx = dx * iL, cy = dy * iL, cz = dz * iL;  // :378-379
  const double h2 = cx * cx + cy * cy;
  if (h2 < 1e-12) {                                // vertical member, :380-381 (eps = 1e-6)
    const double s = cz > 0.0 ? 1.0 : -1.0;
    r.t[0] = 0.0; r.t[1] = 0.0; r.t[2] = s;
    r.n1[0] = 0.0; r.n1[1] = 1.0; r.n1[2] = 0.0;
    r.n2[0] = -s; r.n2[1] = 0.0; r.n2[2] = 0.0;
  } else {                                         // :383-384 (a zero-length element lands here
                                                   // with NaN direction cosines, as in numpy)
    const double iD = rsqrt(h2);
    const double D = h2 * iD;
    r.t[0] = cx; r.t[1] = cy; r.t[2] = cz;
    r.n1[0] = -cy * iD; r.n1[1] = cx * iD; r.n1[2] = 0.0;
    r.n2[0] = -cx * cz * iD; r.n2[1] = -cy * cz * iD; r.n2[2] = D;
  }
  const double E = P.E, G = P.G;
  const bool ok = L > 0.0;                         // every term is guarded by L > 0 (:649-653)
  const double iL1 = ok ? iL : 0.0;
  const double iL2 = iL1 * iL1, iL3 = iL2 * iL1;
  // Timoshenko factors with ONE reciprocal per bending plane: with q = 12 E I and
  // den = G kappa A L^2 (:647-648), phi = q/den and
  //   1/(1+phi) = den/(den+q), (4+phi)/(1+phi) = (4 den+q)/(den+q), (2-phi)/(1+phi) = (2 den-q)/(den+q);
  // den <= 0 selects the Euler-Bernoulli fallback phi = 0.
  const double EIz = E * Iy, EIy = E * Ix;         // local x-y bending uses I_y, x-z uses I_x
  const double qz = 12.0 * EIz, qy = 12.0 * EIy;
  const double den_z = G * ky * A * L2;
  const double den_y = G * kz * A * L2;
  double oz = 1.0, f22z = 4.0, f23z = 2.0, oy = 1.0, f22y = 4.0, f23y = 2.0;
  if (den_z > 0.0) {
    const double rq = 1.0 / (den_z + qz);
    oz = den_z * rq; f22z = (4.0 * den_z + qz) * rq; f23z = (2.0 * den_z - qz) * rq;
  }
  if (den_y > 0.0) {
    const double rq = 1.0 / (den_y + qy);
    oy = den_y * rq; f22y = (4.0 * den_y + qy) * rq; f23y = (2.0 * den_y - qy) * rq;
  }
  r.k11z = qz * iL3 * oz;
  r.k12z = 0.5 * qz * iL2 * oz;
  r.k22z = f22z * EIz * iL1;
  r.k23z = f23z * EIz * iL1;
  r.k11y = qy * iL3 * oy;
  r.k12y = 0.5 * qy * iL2 * oy;
  r.k22y = f22y * EIy * iL1;
  r.k23y = f23y * EIy * iL1;
  r.tor = G * J * iL1;                             // :653
  r.ax = A * E * iL1;                              // :655
  const double hl = 0.5 * P.rho * L;               // BeamSolver.py:667-670
  r.mt = hl * A; r.mrx = hl * J; r.mry = hl * Ix; r.mrz = hl * Iy;
  r.L = L;
}

// acc (6x6 row-major) (+)= block [a][b] of R^T k R.  ASSIGN: overwrite instead of add.
template <bool ASSIGN>
__device__ __forceinline__ void frame_kblock(const FrameRec& R, int a, int b, double* acc) {
  const bool same = (a == b);
  const double suu = same ? 1.0 : -1.0;
  const double d0 = suu * R.ax, d1 = suu * R.k11z, d2 = suu * R.k11y;
  const double sa = (a == 0) ? 1.0 : -1.0;
  const double sb = (b == 0) ? 1.0 : -1.0;
  const double az = sa * R.k12z, ay = sa * R.k12y;
  const double bz = sb * R.k12z, by = sb * R.k12y;
  const double e0 = same ? R.tor : -R.tor;
  const double e1 = same ? R.k22y : R.k23y;
  const double e2 = same ? R.k22z : R.k23z;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double tt = R.t[r] * R.t[c];
      const double n11 = R.n1[r] * R.n1[c];
      const double n22 = R.n2[r] * R.n2[c];
      const double n12 = R.n1[r] * R.n2[c];
      const double n21 = R.n2[r] * R.n1[c];
      const double uu = d0 * tt + d1 * n11 + d2 * n22;
      const double ut = az * n12 - ay * n21;
      const double tu = bz * n21 - by * n12;
      const double th = e0 * tt + e1 * n11 + e2 * n22;
      if (ASSIGN) {
        acc[r * 6 + c] = uu; acc[r * 6 + 3 + c] = ut;
        acc[(3 + r) * 6 + c] = tu; acc[(3 + r) * 6 + 3 + c] = th;
      } else {
        acc[r * 6 + c] += uu; acc[r * 6 + 3 + c] += ut;
        acc[(3 + r) * 6 + c] += tu; acc[(3 + r) * 6 + 3 + c] += th;
      }
    }
  }
}

// row6 = row `row` (0..5) of the off-diagonal block [a][1-a] of R^T k R — the same expressions
// as frame_kblock, one row at a time so the pair kernel can stream rows out of registers.
__device__ __forceinline__ void frame_offdiag_row(const FrameRec& R, int a, int row, double* row6) {
  const double sa = (a == 0) ? 1.0 : -1.0;
  const int r = row < 3 ? row : row - 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double tt = R.t[r] * R.t[c];
    const double n11 = R.n1[r] * R.n1[c];
    const double n22 = R.n2[r] * R.n2[c];
    const double n12 = R.n1[r] * R.n2[c];
    const double n21 = R.n2[r] * R.n1[c];
    if (row < 3) {
      row6[c] = (-R.ax) * tt + (-R.k11z) * n11 + (-R.k11y) * n22;
      row6[3 + c] = (sa * R.k12z) * n12 - (sa * R.k12y) * n21;
    } else {
      row6[c] = (-sa * R.k12z) * n21 - (-sa * R.k12y) * n12;
      row6[3 + c] = (-R.tor) * tt + R.k23y * n11 + R.k23z * n22;
    }
  }
}

// Upper triangle (21 values, row-major r<=c) of the diagonal block [a][a] of R^T k R, which is
// exactly symmetric in floating point (every entry is the same sum of commuting products),
// plus the compact lumped mass of that end: m[0] = rho A L/2 (translational, times I) and
// m[1..6] = upper triangle of the rotational 3x3 block (BeamSolver.py:662-675, :388).
__device__ __forceinline__ void frame_diag_sym(const FrameRec& R, int a, double* d, double* m) {
  const double sa = (a == 0) ? 1.0 : -1.0;
  const double az = sa * R.k12z, ay = sa * R.k12y;
  int q = 0, qm = 1;
  m[0] = R.mt;
#pragma unroll
  for (int r = 0; r < 3; ++r) {  // rows 0..2: uu(r, c>=r) then u-theta(r, 0..2)
#pragma unroll
    for (int c = r; c < 3; ++c)
      d[q++] = R.ax * (R.t[r] * R.t[c]) + R.k11z * (R.n1[r] * R.n1[c]) + R.k11y * (R.n2[r] * R.n2[c]);
#pragma unroll
    for (int c = 0; c < 3; ++c) d[q++] = az * (R.n1[r] * R.n2[c]) - ay * (R.n2[r] * R.n1[c]);
  }
#pragma unroll
  for (int r = 0; r < 3; ++r) {  // rows 3..5: theta-theta(r, c>=r)
#pragma unroll
    for (int c = r; c < 3; ++c) {
      const double tt = R.t[r] * R.t[c], n11 = R.n1[r] * R.n1[c], n22 = R.n2[r] * R.n2[c];
      d[q++] = R.tor * tt + R.k22y * n11 + R.k22z * n22;
      m[qm++] = R.mrx * tt + R.mry * n11 + R.mrz * n22;
    }
  }
}

// index into the 21-value upper triangle of a symmetric 6x6
__device__ __forceinline__ int sym6_index(int r, int c) {
  const int lo = r < c ? r : c, hi = r < c ? c : r;
  return lo * 6 - (lo * (lo - 1)) / 2 + (hi - lo);
}

// acc (6x6) = diagonal block [a][a] of R^T m R (lumped mass, BeamSolver.py:662-675, :388).
__device__ __forceinline__ void frame_mblock(const FrameRec& R, double* acc) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double tt = R.t[r] * R.t[c];
      const double n11 = R.n1[r] * R.n1[c];
      const double n22 = R.n2[r] * R.n2[c];
      acc[r * 6 + c] = (r == c) ? R.mt : 0.0;  // lambda^T (m_t I) lambda = m_t I
      acc[r * 6 + 3 + c] = 0.0;
      acc[(3 + r) * 6 + c] = 0.0;
      acc[(3 + r) * 6 + 3 + c] = R.mrx * tt + R.mry * n11 + R.mrz * n22;
    }
  }
}

// ------------------------------------------------------------------------------ Tet10
struct Tet10Params {
  const double* xyz;     // (n_nodes,3)
  const int32_t* conn;   // (n_elem,10) meshio node order
  double lam, mu2, gsh;  // C1*nu, C1*(1-nu), C1*(1-2nu)/2   (ReactionSolver.py:89-98)
  unsigned long long* skipped;  // counter of Gauss points with detJ <= 1e-12
  const double* grad;    // (n_elem, 4, 32): per Gauss point dN_global (10 x 3), detJ*w, pad — or nullptr
};

constexpr int kTetGradStride = 32;   // doubles per (element, Gauss point) record

// natural-coordinate derivative dN_i/d(xi,eta,zeta), ReactionSolver.py:100-113.
__device__ __forceinline__ void tet10_dn(int i, double L1, double L2, double L3, double L4, double* d) {
  double g1 = 0.0, g2 = 0.0, g3 = 0.0, g4 = 0.0;  // dN_i/dL1..dL4
  switch (i) {
    case 0: g1 = 4.0 * L1 - 1.0; break;
    case 1: g2 = 4.0 * L2 - 1.0; break;
    case 2: g3 = 4.0 * L3 - 1.0; break;
    case 3: g4 = 4.0 * L4 - 1.0; break;
    case 4: g1 = 4.0 * L2; g2 = 4.0 * L1; break;
    case 5: g2 = 4.0 * L3; g3 = 4.0 * L2; break;
    case 6: g1 = 4.0 * L3; g3 = 4.0 * L1; break;
    case 7: g1 = 4.0 * L4; g4 = 4.0 * L1; break;
    case 8: g2 = 4.0 * L4; g4 = 4.0 * L2; break;
    default: g3 = 4.0 * L4; g4 = 4.0 * L3; break;
  }
  d[0] = g2 - g1; d[1] = g3 - g1; d[2] = g4 - g1;  // dL^T @ dN_L
}

// acc (3x3 row-major) = block [a][b] of Ke (ReactionSolver.py:126-146).  count_skips: this
// caller is the one contribution per element that reports skipped Gauss points.
__device__ __forceinline__ void tet10_kblock(const Tet10Params& P, uint32_t e, int a, int b,
                                             bool count_skips, double* acc) {
  double X[10][3];
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const double* p = P.xyz + 3 * (size_t)P.conn[10 * (size_t)e + i];
    X[i][0] = __ldg(p); X[i][1] = __ldg(p + 1); X[i][2] = __ldg(p + 2);
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) acc[k] = 0.0;
  const double GA = 0.58541020, GB = 0.13819660;   // 8-digit literals, ReactionSolver.py:120-123
  unsigned skipped = 0;
#pragma unroll 1
  for (int g = 0; g < 4; ++g) {
    const double xi = (g == 0) ? GA : GB, eta = (g == 1) ? GA : GB, zeta = (g == 2) ? GA : GB;
    const double L1 = 1.0 - xi - eta - zeta, L2 = xi, L3 = eta, L4 = zeta;
    double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      double d[3];
      tet10_dn(i, L1, L2, L3, L4, d);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        J[k][0] += d[k] * X[i][0]; J[k][1] += d[k] * X[i][1]; J[k][2] += d[k] * X[i][2];
      }
    }
    const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
    const double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
    const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    const double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
    if (det <= 1e-12) { ++skipped; continue; }     // ReactionSolver.py:133-135
    const double id = 1.0 / det;
    double iJ[3][3];
    iJ[0][0] = c00 * id; iJ[1][0] = c01 * id; iJ[2][0] = c02 * id;
    iJ[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id;
    iJ[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id;
    iJ[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id;
    iJ[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
    iJ[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
    iJ[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
    double da[3], db[3], ga[3], gb[3];
    tet10_dn(a, L1, L2, L3, L4, da);
    tet10_dn(b, L1, L2, L3, L4, db);
#pragma unroll
    for (int k = 0; k < 3; ++k) {                  // dN_global = inv(J) @ dN_natural, :137
      ga[k] = iJ[k][0] * da[0] + iJ[k][1] * da[1] + iJ[k][2] * da[2];
      gb[k] = iJ[k][0] * db[0] + iJ[k][1] * db[1] + iJ[k][2] * db[2];
    }
    const double w = det * 0.25;                   // detJ * w, w = 1/4 (:124,146)
    const double xx = ga[0] * gb[0], yy = ga[1] * gb[1], zz = ga[2] * gb[2];
    acc[0] += (P.mu2 * xx + P.gsh * (yy + zz)) * w;
    acc[1] += (P.lam * ga[0] * gb[1] + P.gsh * ga[1] * gb[0]) * w;
    acc[2] += (P.lam * ga[0] * gb[2] + P.gsh * ga[2] * gb[0]) * w;
    acc[3] += (P.lam * ga[1] * gb[0] + P.gsh * ga[0] * gb[1]) * w;
    acc[4] += (P.mu2 * yy + P.gsh * (xx + zz)) * w;
    acc[5] += (P.lam * ga[1] * gb[2] + P.gsh * ga[2] * gb[1]) * w;
    acc[6] += (P.lam * ga[2] * gb[0] + P.gsh * ga[0] * gb[2]) * w;
    acc[7] += (P.lam * ga[2] * gb[1] + P.gsh * ga[1] * gb[2]) * w;
    acc[8] += (P.mu2 * zz + P.gsh * (yy + xx)) * w;
  }
  if (count_skips && skipped) atomicAdd(P.skipped, (unsigned long long)skipped);
}

// ---- two-stage Tet10: the geometry of a Gauss point is evaluated ONCE per (element, point) into a
// 256-byte record {dN_global[10][3], detJ*w, 0} (w = 0 for a skipped point), and each of the 100
// block contributions of the element reads the two gradients it needs.  The one-stage tet10_kblock
// above evaluates all four Jacobians in every one of the 100 contribution threads — measured 56 M
// tets/s at 4 % of the HBM roofline, i.e. compute-bound on redundant work.  Same expressions, same
// order of operations as the one-stage path (values agree to rounding; both are tested at 1e-13).
__device__ __forceinline__ void tet10_point_record(const Tet10Params& P, uint32_t e, int g, double* __restrict__ rec) {
  double X[10][3];
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const double* p = P.xyz + 3 * (size_t)P.conn[10 * (size_t)e + i];
    X[i][0] = __ldg(p); X[i][1] = __ldg(p + 1); X[i][2] = __ldg(p + 2);
  }
  const double GA = 0.58541020, GB = 0.13819660;   // 8-digit literals, ReactionSolver.py:120-123
  const double xi = (g == 0) ? GA : GB, eta = (g == 1) ? GA : GB, zeta = (g == 2) ? GA : GB;
  const double L1 = 1.0 - xi - eta - zeta, L2 = xi, L3 = eta, L4 = zeta;
  double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    double d[3];
    tet10_dn(i, L1, L2, L3, L4, d);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      J[k][0] += d[k] * X[i][0]; J[k][1] += d[k] * X[i][1]; J[k][2] += d[k] * X[i][2];
    }
  }
  const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
  const double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
  const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
  const double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
  if (det <= 1e-12) {                              // ReactionSolver.py:133-135: skipped and counted
#pragma unroll
    for (int k = 0; k < kTetGradStride; ++k) rec[k] = 0.0;
    atomicAdd(P.skipped, 1ull);
    return;
  }
  const double id = 1.0 / det;
  double iJ[3][3];
  iJ[0][0] = c00 * id; iJ[1][0] = c01 * id; iJ[2][0] = c02 * id;
  iJ[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id;
  iJ[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id;
  iJ[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id;
  iJ[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
  iJ[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
  iJ[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    double d[3];
    tet10_dn(i, L1, L2, L3, L4, d);
#pragma unroll
    for (int k = 0; k < 3; ++k) rec[3 * i + k] = iJ[k][0] * d[0] + iJ[k][1] * d[1] + iJ[k][2] * d[2];   // :137
  }
  rec[30] = det * 0.25;                            // detJ * w, w = 1/4 (:124,146)
  rec[31] = 0.0;
}

// acc (3x3 row-major) = block [a][b] of Ke from the stored point records
__device__ __forceinline__ void tet10_kblock_stored(const Tet10Params& P, uint32_t e, int a, int b, double* acc) {
#pragma unroll
  for (int k = 0; k < 9; ++k) acc[k] = 0.0;
  const double* base = P.grad + (size_t)e * 4 * kTetGradStride;
#pragma unroll 1
  for (int g = 0; g < 4; ++g) {
    const double* rec = base + g * kTetGradStride;
    const double w = __ldg(rec + 30);
    if (w == 0.0) continue;                        // skipped point (detJ <= 1e-12)
    double ga[3], gb[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { ga[k] = __ldg(rec + 3 * a + k); gb[k] = __ldg(rec + 3 * b + k); }
    const double xx = ga[0] * gb[0], yy = ga[1] * gb[1], zz = ga[2] * gb[2];
    acc[0] += (P.mu2 * xx + P.gsh * (yy + zz)) * w;
    acc[1] += (P.lam * ga[0] * gb[1] + P.gsh * ga[1] * gb[0]) * w;
    acc[2] += (P.lam * ga[0] * gb[2] + P.gsh * ga[2] * gb[0]) * w;
    acc[3] += (P.lam * ga[1] * gb[0] + P.gsh * ga[0] * gb[1]) * w;
    acc[4] += (P.mu2 * yy + P.gsh * (xx + zz)) * w;
    acc[5] += (P.lam * ga[1] * gb[2] + P.gsh * ga[2] * gb[1]) * w;
    acc[6] += (P.lam * ga[2] * gb[0] + P.gsh * ga[0] * gb[2]) * w;
    acc[7] += (P.lam * ga[2] * gb[1] + P.gsh * ga[1] * gb[2]) * w;
    acc[8] += (P.mu2 * zz + P.gsh * (yy + xx)) * w;
  }
}

}  // namespace femb
