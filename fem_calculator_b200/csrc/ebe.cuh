// Device pieces of the matrix-free frame operator (ebe.cu): the stiffness part of the element record and the
// closed form of K_e[a][a] ua + K_e[a][1-a] uo (BeamSolver.py:386-387, 646-660).
#pragma once

#include "common.cuh"
#include "elements.cuh"

namespace femb {

// stiffness part of the element record (the lumped-mass terms of FrameRec are dead code here)
struct KRec {
  double t[3], n1[3], n2[3];
  double ax, tor, k11z, k11y, c12z, c12y, k23z, k23y, d22z, d22y;   // c12 = s_a*k12, d22 = k22 - k23
};

// Lean form of frame_record_from (elements.cuh) for the operator: only the stiffness magnitudes
// that ebe_apply needs, with the Timoshenko algebra reduced per bending plane to
//   psi = 1/(1+phi) = den/(den + 12 EI)   (den = G kappa A L^2; den <= 0 -> psi = 1, BeamSolver.py:647-648)
//   e = EI/L,  w = psi e:   k23 = (2-phi)/(1+phi) e = 3w - e,   k22 - k23 = 2e,
//   k12 = 6 w / L,  k11 = 2 k12 / L                                         (BeamSolver.py:649-652)
// and n2 = t x n1 (n1.z = 0 in both branches of BeamSolver.py:380-384).  ~75 FP64 instructions
// instead of ~100; values agree with the assembled K to a few ulp (tests/test_gpu_operator.py).
__device__ __forceinline__ void krec_from(const FrameParams& P, const FrameIn& in, int a, KRec& k) {
  const double dx = in.dx, dy = in.dy, dz = in.dz;
  const double L2 = dx * dx + dy * dy + dz * dz;
  const double iL = rsqrt(L2);
  const double cx = dx * iL, cy = dy * iL, cz = dz * iL;
  const double h2 = cx * cx + cy * cy;
  if (h2 < 1e-12) {                                  // vertical member (eps = 1e-6)
    const double s = cz > 0.0 ? 1.0 : -1.0;
    k.t[0] = 0.0; k.t[1] = 0.0; k.t[2] = s;
    k.n1[0] = 0.0; k.n1[1] = 1.0; k.n1[2] = 0.0;
    k.n2[0] = -s; k.n2[1] = 0.0; k.n2[2] = 0.0;
  } else {
    const double iD = rsqrt(h2);
    k.t[0] = cx; k.t[1] = cy; k.t[2] = cz;
    k.n1[0] = -cy * iD; k.n1[1] = cx * iD; k.n1[2] = 0.0;
    k.n2[0] = -cz * k.n1[1]; k.n2[1] = cz * k.n1[0]; k.n2[2] = cx * k.n1[1] - cy * k.n1[0];
  }
  const double iL1 = (L2 > 0.0) ? iL : 0.0;          // every term is guarded by L > 0 (BeamSolver.py:649-655)
  const double sa = a ? -1.0 : 1.0;
  const double E = P.E, G = P.G;
  k.ax = in.A * E * iL1;
  k.tor = G * in.J * iL1;
  {
    const double EI = E * in.Iy;                     // local x-y bending uses I_y
    const double den = G * in.ky * in.A * L2;
    const double e = EI * iL1;
    double w = e;
    if (den > 0.0) w = e * (den / (den + 12.0 * EI));
    const double k12 = 6.0 * w * iL1;
    k.k23z = 3.0 * w - e; k.d22z = 2.0 * e; k.c12z = sa * k12; k.k11z = 2.0 * k12 * iL1;
  }
  {
    const double EI = E * in.Ix;                     // local x-z bending uses I_x
    const double den = G * in.kz * in.A * L2;
    const double e = EI * iL1;
    double w = e;
    if (den > 0.0) w = e * (den / (den + 12.0 * EI));
    const double k12 = 6.0 * w * iL1;
    k.k23y = 3.0 * w - e; k.d22y = 2.0 * e; k.c12y = sa * k12; k.k11y = 2.0 * k12 * iL1;
  }
}

// The four numbers of the element record that cost a square root or a division — 1/L, 1/sqrt(cx^2 + cy^2) and the
// two Timoshenko factors w = psi EI / L — depend on the geometry and the section only: ebe_pair_aux_kernel (ebe.cu)
// stores them per (node, element end) pair once per assembled K, computed by exactly the expressions above, and the
// persistent PCG kernel rebuilds the record from them with ~35 multiply-adds instead of ~75 FP64 instructions in four
// long dependent chains (ncu, profiles/r02_ncu_full_persistent_pcg.txt: the two divisions and the two rsqrt were
// 77 of the ~235 instructions of a pair and most of its fixed-latency stalls).  Same values, bit for bit.
struct PairAux { double iL, iD, wz, wy; };

__device__ __forceinline__ PairAux pair_aux_from(const FrameParams& P, const FrameIn& in) {
  PairAux x;
  const double L2 = in.dx * in.dx + in.dy * in.dy + in.dz * in.dz;
  const double iL = rsqrt(L2);
  const double cx = in.dx * iL, cy = in.dy * iL;
  const double h2 = cx * cx + cy * cy;
  x.iL = iL;
  x.iD = (h2 < 1e-12) ? 0.0 : rsqrt(h2);
  const double iL1 = (L2 > 0.0) ? iL : 0.0;
  const double E = P.E, G = P.G;
  {
    const double EI = E * in.Iy;
    const double den = G * in.ky * in.A * L2;
    const double e = EI * iL1;
    double w = e;
    if (den > 0.0) w = e * (den / (den + 12.0 * EI));
    x.wz = w;
  }
  {
    const double EI = E * in.Ix;
    const double den = G * in.kz * in.A * L2;
    const double e = EI * iL1;
    double w = e;
    if (den > 0.0) w = e * (den / (den + 12.0 * EI));
    x.wy = w;
  }
  return x;
}

__device__ __forceinline__ void krec_from_aux(const FrameParams& P, const FrameIn& in, int a, const PairAux& x, KRec& k) {
  const double dx = in.dx, dy = in.dy, dz = in.dz;
  const double L2 = dx * dx + dy * dy + dz * dz;
  const double iL = x.iL;
  const double cx = dx * iL, cy = dy * iL, cz = dz * iL;
  const double h2 = cx * cx + cy * cy;
  if (h2 < 1e-12) {                                  // vertical member (eps = 1e-6)
    const double s = cz > 0.0 ? 1.0 : -1.0;
    k.t[0] = 0.0; k.t[1] = 0.0; k.t[2] = s;
    k.n1[0] = 0.0; k.n1[1] = 1.0; k.n1[2] = 0.0;
    k.n2[0] = -s; k.n2[1] = 0.0; k.n2[2] = 0.0;
  } else {
    const double iD = x.iD;
    k.t[0] = cx; k.t[1] = cy; k.t[2] = cz;
    k.n1[0] = -cy * iD; k.n1[1] = cx * iD; k.n1[2] = 0.0;
    k.n2[0] = -cz * k.n1[1]; k.n2[1] = cz * k.n1[0]; k.n2[2] = cx * k.n1[1] - cy * k.n1[0];
  }
  const double iL1 = (L2 > 0.0) ? iL : 0.0;
  const double sa = a ? -1.0 : 1.0;
  const double E = P.E, G = P.G;
  k.ax = in.A * E * iL1;
  k.tor = G * in.J * iL1;
  {
    const double e = E * in.Iy * iL1;
    const double w = x.wz;
    const double k12 = 6.0 * w * iL1;
    k.k23z = 3.0 * w - e; k.d22z = 2.0 * e; k.c12z = sa * k12; k.k11z = 2.0 * k12 * iL1;
  }
  {
    const double e = E * in.Ix * iL1;
    const double w = x.wy;
    const double k12 = 6.0 * w * iL1;
    k.k23y = 3.0 * w - e; k.d22y = 2.0 * e; k.c12y = sa * k12; k.k11y = 2.0 * k12 * iL1;
  }
}

__device__ __forceinline__ double dot3(const double* a, double x, double y, double z) {
  return a[0] * x + a[1] * y + a[2] * z;
}

// out[0..5] = K_e[a][a] ua + K_e[a][1-a] uo in global axes.  With d = translations, th = rotations
// and the projections on (t, n1, n2) the two block rows of elements.cuh collapse to
//   force : ax (t.dd) t + [k11z (n1.dd) + c12z (n2.ts)] n1 + [k11y (n2.dd) - c12y (n1.ts)] n2
//   moment: tor (t.td) t + [-c12y (n2.dd) + k23y (n1.ts) + d22y (n1.ta)] n1
//                        + [ c12z (n1.dd) + k23z (n2.ts) + d22z (n2.ta)] n2
// where dd = d_a - d_o, td = th_a - th_o, ts = th_a + th_o, ta = th_a.
__device__ __forceinline__ void ebe_apply(const KRec& k, const double* ua, const double* uo, double* out) {
  const double ddx = ua[0] - uo[0], ddy = ua[1] - uo[1], ddz = ua[2] - uo[2];
  const double tdx = ua[3] - uo[3], tdy = ua[4] - uo[4], tdz = ua[5] - uo[5];
  const double tsx = ua[3] + uo[3], tsy = ua[4] + uo[4], tsz = ua[5] + uo[5];
  // n1.z == 0 in both branches of the direction-cosine matrix: its products are left out
  const double dt = dot3(k.t, ddx, ddy, ddz), d1 = k.n1[0] * ddx + k.n1[1] * ddy, d2 = dot3(k.n2, ddx, ddy, ddz);
  const double tt = dot3(k.t, tdx, tdy, tdz);
  const double s1 = k.n1[0] * tsx + k.n1[1] * tsy, s2 = dot3(k.n2, tsx, tsy, tsz);
  const double a1 = k.n1[0] * ua[3] + k.n1[1] * ua[4], a2 = dot3(k.n2, ua[3], ua[4], ua[5]);
  const double ft = k.ax * dt;
  const double f1 = k.k11z * d1 + k.c12z * s2;
  const double f2 = k.k11y * d2 - k.c12y * s1;
  const double mt = k.tor * tt;
  const double m1 = k.k23y * s1 + k.d22y * a1 - k.c12y * d2;
  const double m2 = k.k23z * s2 + k.d22z * a2 + k.c12z * d1;
  out[0] = ft * k.t[0] + f1 * k.n1[0] + f2 * k.n2[0];
  out[1] = ft * k.t[1] + f1 * k.n1[1] + f2 * k.n2[1];
  out[2] = ft * k.t[2] + f2 * k.n2[2];
  out[3] = mt * k.t[0] + m1 * k.n1[0] + m2 * k.n2[0];
  out[4] = mt * k.t[1] + m1 * k.n1[1] + m2 * k.n2[1];
  out[5] = mt * k.t[2] + m2 * k.n2[2];
}

__device__ __forceinline__ void load6(const double* __restrict__ x, int node, double* u) {
  const double2* p = reinterpret_cast<const double2*>(x + (size_t)node * 6);
  const double2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
  u[0] = a.x; u[1] = a.y; u[2] = b.x; u[3] = b.y; u[4] = c.x; u[5] = c.y;
}

// y = K_ff x over all nodes by the CTAs [cta, cta + ncta, ..) of a PERSISTENT kernel (lines.cu): the node-gather
// mapping of frame_ebe_node_kernel<1, 1, 2> (two lanes per node, interleaved steps of THREADS / 2 nodes) as a device
// function.  x was written by other CTAs of the same launch: it is read with plain coherent loads (the grid barrier
// before the phase carries acquire semantics), never through the non-coherent read-only path.  Returns the thread's
// share of (x, y).
template <int THREADS>
__device__ __forceinline__ double ebe_nodes_phase(const FrameParams& P, const int4* __restrict__ pair_rec,
                                                  const int4* __restrict__ node_rec, int n_nodes,
                                                  const uint8_t* __restrict__ free_mask, const double* x, double* y,
                                                  int cta, int ncta, const double4* __restrict__ pair_aux) {
  constexpr int NPC = THREADS / 2;
  const int part = threadIdx.x & 1;
  double dot = 0.0;
  for (int base = cta * NPC; base < n_nodes; base += ncta * NPC) {
    const int node = base + (threadIdx.x >> 1);
    const bool active = node < n_nodes;
    int first = 0, count = 0;
    double px = 0.0, py = 0.0, pz = 0.0;
    double ua[6], acc[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) { ua[c] = 0.0; acc[c] = 0.0; }
    if (active) {
      const int4 nr = __ldg(node_rec + node);
      first = nr.x; count = nr.y;
      const double* pp = P.xyz + 3 * (size_t)node;
      px = __ldg(pp); py = __ldg(pp + 1); pz = __ldg(pp + 2);
      const double2* xp = reinterpret_cast<const double2*>(x + (size_t)node * 6);
      const double2 a = xp[0], b = xp[1], c2 = xp[2];
      ua[0] = a.x; ua[1] = a.y; ua[2] = b.x; ua[3] = b.y; ua[4] = c2.x; ua[5] = c2.y;
    }
#pragma unroll 2
    for (int j = part; j < count; j += 2) {
      const int4 rec = __ldg(pair_rec + first + j);
      double4 av;      // one 32-byte read-only load
      asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(av.x), "=d"(av.y), "=d"(av.z), "=d"(av.w) : "l"(pair_aux + first + j));
      const int a = (rec.w >> 24) & 1;
      const double* po = P.xyz + 3 * (size_t)rec.y;
      const double ox = __ldg(po), oy = __ldg(po + 1), oz = __ldg(po + 2);
      const double* sp = P.sec_props + 8 * (size_t)(rec.w & 0xFFFFFF);
      FrameIn in;   // element direction: end 0 -> end 1, as in the assembly kernel
      in.dx = a ? px - ox : ox - px; in.dy = a ? py - oy : oy - py; in.dz = a ? pz - oz : oz - pz;
      in.A = __ldg(sp); in.Ix = __ldg(sp + 1); in.Iy = __ldg(sp + 2); in.J = __ldg(sp + 3);
      in.ky = 0.0; in.kz = 0.0;
      double uo[6];
      {
        const double2* xo = reinterpret_cast<const double2*>(x + (size_t)rec.y * 6);
        const double2 u0 = xo[0], u1 = xo[1], u2 = xo[2];
        uo[0] = u0.x; uo[1] = u0.y; uo[2] = u1.x; uo[3] = u1.y; uo[4] = u2.x; uo[5] = u2.y;
      }
      KRec k;
      const PairAux pa = {av.x, av.y, av.z, av.w};
      krec_from_aux(P, in, a, pa, k);
      double o6[6];
      ebe_apply(k, ua, uo, o6);
#pragma unroll
      for (int c = 0; c < 6; ++c) acc[c] += o6[c];
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 1);
    if (active && part == 0) {
      const uint8_t* fm = free_mask + (size_t)node * 6;
#pragma unroll
      for (int c = 0; c < 6; ++c)
        if (!fm[c]) acc[c] = ua[c];      // identity rows on the fixed DOFs
      double2* yp = reinterpret_cast<double2*>(y + (size_t)node * 6);
      yp[0] = make_double2(acc[0], acc[1]); yp[1] = make_double2(acc[2], acc[3]); yp[2] = make_double2(acc[4], acc[5]);
#pragma unroll
      for (int c = 0; c < 6; ++c) dot += ua[c] * acc[c];
    }
  }
  return dot;
}

}  // namespace femb
