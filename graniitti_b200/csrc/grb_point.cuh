// grb_point.cuh -- one phase-space point of the <F> class on the device: sampling map, 2->3 skeleton,
// central decay chain, Lorentz scalars, cuts, Born amplitude, and the per-node work of the screening loop.
// See grb_dev.cuh for the arithmetic conventions (-fmad=false, explicit fma()).
#pragma once
#include "grb_dev.cuh"
#include "grb_fastmath.cuh"

namespace grb {

extern __constant__ DevProc c_proc;  // defined in grb_consts.cu (relocatable device code)
extern __constant__ DevLoop c_loop;

}  // namespace grb

#include "grb_res.cuh"

namespace grb {

// ---- Philox4x32-10 counter-based RNG (Salmon et al., SC'11): key = seed, counter = (point, iter, block) --
struct Philox {
  uint32_t k0, k1;
  __device__ __forceinline__ void block(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]) const {
    uint32_t ka = k0, kb = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
      const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
      const uint32_t n0 = hi1 ^ c1 ^ ka, n1 = lo1, n2 = hi0 ^ c3 ^ kb, n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      ka += 0x9E3779B9u;
      kb += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
  }
};
__device__ __forceinline__ double u01_from_bits(uint32_t hi, uint32_t lo) {
  // 53-bit uniform in [0,1), like std::uniform_real_distribution<double> over a 64-bit engine
  const uint64_t x = ((uint64_t)hi << 32) | lo;
  return (double)(x >> 11) * 0x1.0p-53;
}

// Source of the VEGAS-visible coordinates u[] and of the internal draws ("tape") of one point.
// (a) caller supplied SoA arrays (parity mode, grb_eval); (b) Philox (grb_vegas_iterate).
struct DrawsSoA {
  const double *u, *tape;
  int64_t       n, idx;
  int           dtape, pos;
  bool          overrun;
  // rng: the internal draws come from the point's Philox stream instead of a caller tape (grb_eval_rng: seam A as the
  // reference calls it -- the caller supplies randvec only, EventWeight draws from its own generator).  Draw d0 + pos, the
  // same stream positions the VEGAS mode uses after its d0 = dim coordinates.
  bool     rng;
  Philox   g;
  uint32_t p_lo, p_hi, iter;
  int      d0;
  __device__ __forceinline__ double coord(int j) const { return u[(int64_t)j * n + idx]; }
  __device__ __forceinline__ double next() {
    double v = 0.5;
    if (rng) {
      const int d = d0 + pos;
      uint32_t  o[4];
      g.block(p_lo, p_hi, iter, (uint32_t)(d >> 1), o);
      v = (d & 1) ? u01_from_bits(o[3], o[2]) : u01_from_bits(o[1], o[0]);
    } else if (pos < dtape) v = tape[(int64_t)pos * n + idx];
    else overrun = true;
    ++pos;
    return v;
  }
};
struct DrawsPhilox {
  Philox   g;
  uint32_t p_lo, p_hi, iter;
  int      pos;       // next draw index (0..: coordinates first, then tape)
  double   cached;
  bool     overrun;
  __device__ __forceinline__ double draw(int d) {
    uint32_t o[4];
    g.block(p_lo, p_hi, iter, (uint32_t)(d >> 1), o);
    return (d & 1) ? u01_from_bits(o[3], o[2]) : u01_from_bits(o[1], o[0]);
  }
  __device__ __forceinline__ double next() {
    double v;
    if (pos & 1) {
      v = cached;
    } else {
      uint32_t o[4];
      g.block(p_lo, p_hi, iter, (uint32_t)(pos >> 1), o);
      v      = u01_from_bits(o[1], o[0]);
      cached = u01_from_bits(o[3], o[2]);
    }
    ++pos;
    return v;
  }
};

// ---- per-point state ------------------------------------------------------------------------------------
template <int NN>
struct PointState {
  v4     p1, p2, pX;       // lts.pfinal[1], [2], [0]
  v4     np4[NN];          // decay tree node momenta; nodes [0,n_legs) are lts.pfinal[3..]
  double moff[NN];         // off-shell masses
  double M_MIN, M_MAX, DW;
  double t1, t2, s1, s2, t_hat, u_hat, m2, Y, Pt;
  double ss13, ss14, ss23, ss24;
  double qx, qy;           // <C> class: intermediate difference momentum q0 (MContinuum::pkt_[0])
  int    exc;              // bit 0 / 1: forward leg 1 / 2 is an excited N* system (lts.excite1/2)
  double qprod;            // <C> class: prod_i |q_i| of the intermediate difference momenta (BNPhaseSpaceWeight)
};

// reference src/MRandom.cc:82-107 RelativisticBWRandom; node constants are pre-digested on the host with
// the reference's libm for the M_MIN = 0 case (final-state legs); mothers with a lower bound take atan2 here
template <class Draws>
__device__ __forceinline__ double bw_mass(const DevNode &nd, double M_MIN_arg, bool has_min, Draws &dr) {
  if (!nd.bw_active) return nd.mass;
  double A, BA;
  if (!has_min) {
    A  = nd.bwA0;
    BA = nd.bwBA0;
  } else {
    const double m2min = fmax(fmax(0.0, pow2(M_MIN_arg)), nd.m2lo);
    A  = atan2(m2min - nd.m2, nd.m0G);
    BA = atan2(nd.m2max - nd.m2, nd.m0G) - A;
  }
  double m2val = 0.0;
  for (int it = 0; it < 64; ++it) {
    const double R = dr.next();
    m2val = nd.m2 + nd.m0G * tan(A + R * BA);
    if (m2val >= 0) break;
  }
  return msqrt(m2val);
}

// reference src/MProcess.cc:753-812 GetOffShellMass
template <class Draws>
__device__ __forceinline__ double offshell_mass(int node, Draws &dr) {
  const DevNode &b = c_proc.nodes[node];
  if (b.width < 1e-40) return b.mass;
  double dm = 0.0;
  if (b.n_children != 0) {
    for (int i = 0; i < b.n_children; ++i)
      dm += fmax(0.0, bw_mass(c_proc.nodes[b.first_child + i], 0.0, false, dr));
    dm += 1e-5;
    return fmax(dm, bw_mass(b, dm, true, dr));
  }
  return fmax(dm, bw_mass(b, 0.0, false, dr));
}

// Lorentz scalars needed downstream (reference src/MProcess.cc:1499-1584 computes ~60; the ss[i][j] >= 0
// tests on the pairs not kept here cannot fire: ss >= (m_i+m_j)^2 exceeds the rounding noise by >1e3)
template <int NN>
__device__ __forceinline__ bool lorentz_scalars(PointState<NN> &S) {
  const v4 &p3 = S.np4[0], &p4 = S.np4[1];
  S.s1   = m2_of(S.p1.x + S.pX.x, S.p1.y + S.pX.y, S.p1.z + S.pX.z, S.p1.e + S.pX.e);  // ss[0][1]
  S.s2   = m2_of(S.p2.x + S.pX.x, S.p2.y + S.pX.y, S.p2.z + S.pX.z, S.p2.e + S.pX.e);  // ss[0][2]
  const double s12 = m2_of(S.p1.x + S.p2.x, S.p1.y + S.p2.y, S.p1.z + S.p2.z, S.p1.e + S.p2.e);
  S.ss13 = m2_of(S.p1.x + p3.x, S.p1.y + p3.y, S.p1.z + p3.z, S.p1.e + p3.e);
  S.ss14 = m2_of(S.p1.x + p4.x, S.p1.y + p4.y, S.p1.z + p4.z, S.p1.e + p4.e);
  S.ss23 = m2_of(S.p2.x + p3.x, S.p2.y + p3.y, S.p2.z + p3.z, S.p2.e + p3.e);
  S.ss24 = m2_of(S.p2.x + p4.x, S.p2.y + p4.y, S.p2.z + p4.z, S.p2.e + p4.e);
  const double s03 = m2_of(S.pX.x + p3.x, S.pX.y + p3.y, S.pX.z + p3.z, S.pX.e + p3.e);
  const double s04 = m2_of(S.pX.x + p4.x, S.pX.y + p4.y, S.pX.z + p4.z, S.pX.e + p4.e);
  const double s34 = m2_of(p3.x + p4.x, p3.y + p4.y, p3.z + p4.z, p3.e + p4.e);
  if (S.s1 < 0 || S.s2 < 0 || s12 < 0 || S.ss13 < 0 || S.ss14 < 0 || S.ss23 < 0 || S.ss24 < 0 || s03 < 0 ||
      s04 < 0 || s34 < 0)
    return false;
  const double q1x = 0.0 - S.p1.x, q1y = 0.0 - S.p1.y, q1z = c_proc.pb1z - S.p1.z, q1e = c_proc.pb1e - S.p1.e;
  const double q2x = 0.0 - S.p2.x, q2y = 0.0 - S.p2.y, q2z = c_proc.pb2z - S.p2.z, q2e = c_proc.pb2e - S.p2.e;
  S.t1 = m2_of(q1x, q1y, q1z, q1e);
  S.t2 = m2_of(q2x, q2y, q2z, q2e);
  if (S.t1 > 0 || S.t2 > 0) return false;
  S.t_hat = m2_of(q1x - p3.x, q1y - p3.y, q1z - p3.z, q1e - p3.e);
  S.u_hat = m2_of(q1x - p4.x, q1y - p4.y, q1z - p4.z, q1e - p4.e);
  S.m2 = m2_of(S.pX);
  S.Y  = rap_of(S.pX);
  S.Pt = msqrt(pt2_of(S.pX));
  return true;
}

// reference src/MFactorized.cc:245-412 B51RandomKin + B51BuildKin, src/MProcess.cc:1231-1284
template <int NN, class Draws>
__device__ __forceinline__ bool cascade_decays(Draws &dr, PointState<NN> &S);

// forward N* system masses (SampleForwardMasses, MProcess.cc:1317-1365): log-flat M^2 from the coordinates ua, ub; single
// excitation picks the side with one internal uniform.  Returns lts.excite1/2 as bits 0, 1.
template <class Draws>
__device__ __forceinline__ int sample_forward_masses(double ua, double ub, Draws &dr, double &m1, double &m2) {
  const DevProc &P = c_proc;
  m1 = P.mb1; m2 = P.mb2;
  if (P.excitation == 1) {
    const double r = exp(P.log_M2f_min + P.dlog_M2f * ua);
    if (dr.next() < 0.5) { m1 = msqrt(r); return 1; }
    m2 = msqrt(r);
    return 2;
  }
  if (P.excitation == 2) {
    const double r1 = exp(P.log_M2f_min + P.dlog_M2f * ua), r2 = exp(P.log_M2f_min + P.dlog_M2f * ub);
    m1 = msqrt(r1); m2 = msqrt(r2);
    return 3;
  }
  return 0;
}

template <int NN, class Draws>
__device__ __forceinline__ bool point_kinematics(const double *u, Draws &dr, PointState<NN> &S) {
  const DevProc &P = c_proc;
  const double   pt1  = exp(P.log_pt0 + P.dlog_pt * u[0]);
  const double   pt2  = exp(P.log_pt0 + P.dlog_pt * u[1]);
  const double   phi1 = 2.0 * GRB_PI * u[2];
  const double   phi2 = 2.0 * GRB_PI * u[3];
  const double   yX   = P.Y_min + P.dY * u[4];
  // off-shell masses of the first-level legs; the M_MIN < M_MAX retry of :264-289 is bounded
  for (int trial = 0;; ++trial) {
    double M_sum = 0.0;
#pragma unroll
    for (int i = 0; i < NN; ++i)
      if (i < P.n_legs) {
        S.moff[i] = offshell_mass(i, dr);
        M_sum += S.moff[i];
      }
    S.M_MIN = fmax(M_sum + 1e-4, P.gM_min);
    S.M_MAX = fmin(P.M_MAX0, P.gM_max);
    if (S.M_MIN < S.M_MAX) break;
    if (trial >= 32) return false;
  }
  const double m2X = pow2(S.M_MIN) + (pow2(S.M_MAX) - pow2(S.M_MIN)) * u[5];
  // forward N* system masses (SampleForwardMasses, MProcess.cc:1317-1365): log-flat M^2; single excitation picks the
  // side with one internal uniform
  double m1 = P.mb1, m2 = P.mb2;
  S.exc = 0;
  if (P.excitation == 1) {
    const double r = exp(P.log_M2f_min + P.dlog_M2f * u[6]);
    if (dr.next() < 0.5) { m1 = msqrt(r); S.exc = 1; }
    else { m2 = msqrt(r); S.exc = 2; }
  } else if (P.excitation == 2) {
    const double r1 = exp(P.log_M2f_min + P.dlog_M2f * u[6]), r2 = exp(P.log_M2f_min + P.dlog_M2f * u[7]);
    m1 = msqrt(r1); m2 = msqrt(r2); S.exc = 3;
  }

  double s1, c1, s2, c2;
  sincos(phi1, &s1, &c1);
  sincos(phi2, &s2, &c2);
  S.p1.x = pt1 * c1; S.p1.y = pt1 * s1;
  S.p2.x = pt2 * c2; S.p2.y = pt2 * s2;
  S.pX.x = -(S.p1.x + S.p2.x);
  S.pX.y = -(S.p1.y + S.p2.y);
  const double mtX = msqrt(m2X + pt2_of(S.pX));
  S.pX.z = mtX * sinh(yX);
  S.pX.e = mtX * cosh(yX);
  if (S.pX.e > (P.sqrt_s - (m1 + m2))) return false;
  PzConst pc;
  pc.s = P.s; pc.t7 = P.sqrt_s; pc.t9 = P.s_pow15; pc.t14 = P.s * P.s;
  const double p1z = solve_pz(m1, m2, pt1, pt2, S.pX.z, S.pX.e, pc);
  const double p2z = -(S.pX.z + p1z);
  if (p1z < 0 || p2z > 0 || isnan(p1z)) return false;
  S.p1.z = p1z; S.p1.e = msqrt(pow2(m1) + pow2(pt1) + pow2(p1z));
  S.p2.z = p2z; S.p2.e = msqrt(pow2(m2) + pow2(pt2) + pow2(p2z));
  // CheckEMC, MMath.h:191-198 (beamsum has px = py = 0)
  {
    const double dx = 0.0 - ((S.p1.x + S.p2.x) + S.pX.x), dy = 0.0 - ((S.p1.y + S.p2.y) + S.pX.y);
    const double dz = P.bsum_z - ((S.p1.z + S.p2.z) + S.pX.z), de = P.bsum_e - ((S.p1.e + S.p2.e) + S.pX.e);
    if (fabs(dx) > 1e-6 || fabs(dy) > 1e-6 || fabs(dz) > 1e-6 || de > 1e-6) return false;
  }
  // central two-body decay (the device path covers 2 first-level legs)
  {
    const double rc = dr.next();
    const double rp = dr.next();
    S.DW = two_body(S.pX, msqrt(m2X), S.moff[0], S.moff[1], rc, rp, S.np4[0], S.np4[1]);
    if (S.DW < 0) return false;
  }
  if (!cascade_decays<NN>(dr, S)) return false;
  return lorentz_scalars(S);
}

// cascade decays in the reference's depth-first order (MProcess::ConstructDecayKinematics)
template <int NN, class Draws>
__device__ __forceinline__ bool cascade_decays(Draws &dr, PointState<NN> &S) {
  const DevProc &P = c_proc;
  if (NN > 2) {
    for (int k = 0; k < P.n_decays; ++k) {
      const int      node = P.decay_order[k];
      const DevNode &b    = P.nodes[node];
      const int      c0 = b.first_child, c1n = b.first_child + 1;
      const double   Mm = m_of(S.np4[node % NN]);
      double         m0v = 0, m1v = 0;
      for (int it = 0;; ++it) {
        m0v = offshell_mass(c0, dr);
        m1v = offshell_mass(c1n, dr);
        if (Mm < (0.0 + m1v) + m0v) {
          if (it >= 64) return false;
          continue;
        }
        break;
      }
      S.moff[c0 % NN]  = m0v;
      S.moff[c1n % NN] = m1v;
      const double rc = dr.next();
      const double rp = dr.next();
      const double w  = two_body(S.np4[node % NN], Mm, m0v, m1v, rc, rp, S.np4[c0 % NN], S.np4[c1n % NN]);
      if (w < 0) return false;
    }
  }
  return true;
}

// ---- <C> phase-space class, two central legs: reference src/MContinuum.cc:276-435 BNRandomKin + BNBuildKin ---------
// u = [pt1, pt2, phi1, phi2, kt, phi_k, y3, y4]; BLinearSystem for Kf = 2 is p3T = (q0 - w)/2, p4T = (-q0 - w)/2 with
// w = p1T + p2T (the reference's matrix loop also adds b_j * 0.0, a signed zero).
template <int NN, class Draws>
__device__ __forceinline__ bool point_kinematics_C(const double *u, Draws &dr, PointState<NN> &S) {
  const DevProc &P = c_proc;
  const double   pt1  = exp(P.log_pt0 + P.dlog_pt * u[0]);
  const double   pt2  = exp(P.log_pt0 + P.dlog_pt * u[1]);
  const double   phi1 = 2.0 * GRB_PI * u[2];
  const double   phi2 = 2.0 * GRB_PI * u[3];
#pragma unroll
  for (int i = 0; i < NN; ++i)
    if (i < P.n_legs) S.moff[i] = offshell_mass(i, dr);
  const double kt   = P.kt_min + P.dkt * u[4];
  const double phik = 2.0 * GRB_PI * u[5];
  const double y3 = P.rap_min + P.drap * u[6], y4 = P.rap_min + P.drap * u[7];
  double m1, m2;   // forward N* masses from the coordinates after the rapidities (MContinuum.cc:322-328)
  S.exc = sample_forward_masses(u[8], u[9], dr, m1, m2);
  S.M_MIN = 0.0; S.M_MAX = 0.0; S.DW = 0.0;
  double s1, c1, s2, c2, sk, ck;
  sincos(phi1, &s1, &c1);
  sincos(phi2, &s2, &c2);
  sincos(phik, &sk, &ck);
  S.p1.x = pt1 * c1; S.p1.y = pt1 * s1;
  S.p2.x = pt2 * c2; S.p2.y = pt2 * s2;
  S.qx = kt * ck; S.qy = kt * sk;
  S.qprod = 1.0 * msqrt(S.qx * S.qx + S.qy * S.qy);
  const double wx = S.p1.x + S.p2.x, wy = S.p1.y + S.p2.y;
  v4 &p3 = S.np4[0], &p4 = S.np4[1];
  p3.x = (S.qx - wx) * (1.0 / 2.0); p3.y = (S.qy - wy) * (1.0 / 2.0);
  p4.x = (S.qx * (-1.0) - wx) * (1.0 / 2.0); p4.y = (S.qy * (-1.0) - wy) * (1.0 / 2.0);
  const double mt3 = msqrt(pow2(S.moff[0]) + pt2_of(p3)), mt4 = msqrt(pow2(S.moff[1]) + pt2_of(p4));
  p3.z = mt3 * sinh(y3); p3.e = mt3 * cosh(y3);
  p4.z = mt4 * sinh(y4); p4.e = mt4 * cosh(y4);
  S.pX.x = p3.x + p4.x; S.pX.y = p3.y + p4.y; S.pX.z = p3.z + p4.z; S.pX.e = p3.e + p4.e;
  if (S.pX.e > P.sqrt_s) return false;
  PzConst pc;
  pc.s = P.s; pc.t7 = P.sqrt_s; pc.t9 = P.s_pow15; pc.t14 = P.s * P.s;
  const double p1z = solve_pz(m1, m2, pt1, pt2, S.pX.z, S.pX.e, pc);
  const double p2z = -(S.pX.z + p1z);
  if (p1z < 0 || p2z > 0 || isnan(p1z)) return false;
  S.p1.z = p1z; S.p1.e = msqrt(pow2(m1) + pow2(pt1) + pow2(p1z));
  S.p2.z = p2z; S.p2.e = msqrt(pow2(m2) + pow2(pt2) + pow2(p2z));
  if (m_of(S.pX) < m_of(p3) + m_of(p4)) return false;  // mass threshold, :415-416 (sumM = 0 + M3 + M4)
  {
    const double dx = 0.0 - ((S.p1.x + S.p2.x) + S.pX.x), dy = 0.0 - ((S.p1.y + S.p2.y) + S.pX.y);
    const double dz = P.bsum_z - ((S.p1.z + S.p2.z) + S.pX.z), de = P.bsum_e - ((S.p1.e + S.p2.e) + S.pX.e);
    if (fabs(dx) > 1e-6 || fabs(dy) > 1e-6 || fabs(dz) > 1e-6 || de > 1e-6) return false;
  }
  if (!cascade_decays<NN>(dr, S)) return false;
  return lorentz_scalars(S);
}

// ---- <C> class, K = 4 or 6 central legs (2 -> 6 / 2 -> 8 direct, 14 / 20 coordinates) + the Regge ladders MRegge::ME6, ME8 --
// u = [pt1, pt2, phi1, phi2, kt_1..K-1, phi_1..K-1, y_1..K]; BLinearSystem with the Kf = K inverse (MContinuum.cc:485-497).
// The ladders need s_1a, s_2a (K + K), s_ab (pairs), t-type scalars tt_1[a], tt_2[a] (K + K) and tt_xy[a][b] (pairs).
#define GRB_BLIN4 {{7.0 / 8.0, 1.0 / 8.0, -1.0 / 2.0, -1.0 / 4.0}, {3.0 / 8.0, 5.0 / 8.0, -1.0 / 2.0, -1.0 / 4.0}, \
                   {-1.0 / 8.0, 1.0 / 8.0, 1.0 / 2.0, -1.0 / 4.0}, {-5.0 / 8.0, -3.0 / 8.0, 1.0 / 2.0, 3.0 / 4.0}}
#define GRB_BLIN6 {{4.0 / 3.0, 1.0 / 2.0, -2.0 / 3.0, -1.0 / 2.0, -1.0 / 3.0, -1.0 / 6.0}, {5.0 / 6.0, 1.0 / 1.0, -2.0 / 3.0, -1.0 / 2.0, -1.0 / 3.0, -1.0 / 6.0}, \
                   {1.0 / 3.0, 1.0 / 2.0, 1.0 / 3.0, -1.0 / 2.0, -1.0 / 3.0, -1.0 / 6.0}, {-1.0 / 6.0, 0.0, 1.0 / 3.0, 1.0 / 2.0, -1.0 / 3.0, -1.0 / 6.0}, \
                   {-2.0 / 3.0, -1.0 / 2.0, 1.0 / 3.0, 1.0 / 2.0, 2.0 / 3.0, -1.0 / 6.0}, {-7.0 / 6.0, -1.0 / 1.0, 1.0 / 3.0, 1.0 / 2.0, 2.0 / 3.0, 5.0 / 6.0}}
template <int K>
struct Ladder {
  static constexpr int NP = K * (K - 1) / 2;
  double s1a[K], s2a[K], sab[NP], tt1[K], tt2[K], txy[NP];  // pair index of (a < b): a*(2K-1-a)/2 + b - a - 1 (a, b = 0..K-1)
  double qx[K - 1], qy[K - 1];                              // intermediate difference momenta (kept for the screening loop)
};
using Ladder4 = Ladder<4>;
template <int K>
__device__ __forceinline__ int pairK(int a, int b) { if (a > b) { const int t = a; a = b; b = t; } return a * (2 * K - 1 - a) / 2 + b - a - 1; }
__device__ __forceinline__ int pair4(int a, int b) { return pairK<4>(a, b); }
// p = A b for one leg: the reference's loop adds b[j] * A[i][j] for every j, zero entries included (signed zeros)
template <int K>
__device__ __forceinline__ double blin_row(int i, const double *b) {
  double acc = 0.0;
  if (K == 4) {
    const double A[4][4] = GRB_BLIN4;
#pragma unroll
    for (int j = 0; j < 4; ++j) acc = acc + b[j] * A[i][j];
  } else {
    const double A[6][6] = GRB_BLIN6;
#pragma unroll
    for (int j = 0; j < 6; ++j) acc = acc + b[j] * A[i][j];
  }
  return acc;
}

// Lorentz scalars of the ladder (MProcess.cc:1499-1584): every pair mass^2 must be >= 0, t1, t2 <= 0
template <int K>
__device__ __forceinline__ bool ladder_scalars(const v4 &p1, const v4 &p2, const v4 &pX, const v4 *leg, double &t1, double &t2,
                                               double &s1, double &s2, Ladder<K> &X) {
  const DevProc &P = c_proc;
  s1 = m2_of(p1.x + pX.x, p1.y + pX.y, p1.z + pX.z, p1.e + pX.e);
  s2 = m2_of(p2.x + pX.x, p2.y + pX.y, p2.z + pX.z, p2.e + pX.e);
  const double s12 = m2_of(p1.x + p2.x, p1.y + p2.y, p1.z + p2.z, p1.e + p2.e);
  bool ok = !(s1 < 0 || s2 < 0 || s12 < 0);
  const double q1x = 0.0 - p1.x, q1y = 0.0 - p1.y, q1z = P.pb1z - p1.z, q1e = P.pb1e - p1.e;
  const double q2x = 0.0 - p2.x, q2y = 0.0 - p2.y, q2z = P.pb2z - p2.z, q2e = P.pb2e - p2.e;
#pragma unroll
  for (int a = 0; a < K; ++a) {
    const v4 &pa = leg[a];
    X.s1a[a] = m2_of(p1.x + pa.x, p1.y + pa.y, p1.z + pa.z, p1.e + pa.e);
    X.s2a[a] = m2_of(p2.x + pa.x, p2.y + pa.y, p2.z + pa.z, p2.e + pa.e);
    const double s0a = m2_of(pX.x + pa.x, pX.y + pa.y, pX.z + pa.z, pX.e + pa.e);
    ok = ok && !(X.s1a[a] < 0 || X.s2a[a] < 0 || s0a < 0);
    X.tt1[a] = m2_of(q1x - pa.x, q1y - pa.y, q1z - pa.z, q1e - pa.e);
    X.tt2[a] = m2_of(q2x - pa.x, q2y - pa.y, q2z - pa.z, q2e - pa.e);
#pragma unroll
    for (int b = a + 1; b < K; ++b) {
      const v4 &pb = leg[b];
      const int k  = a * (2 * K - 1 - a) / 2 + b - a - 1;
      X.sab[k] = m2_of(pa.x + pb.x, pa.y + pb.y, pa.z + pb.z, pa.e + pb.e);
      ok = ok && !(X.sab[k] < 0);
      X.txy[k] = m2_of((q1x - pa.x) - pb.x, (q1y - pa.y) - pb.y, (q1z - pa.z) - pb.z, (q1e - pa.e) - pb.e);
    }
  }
  if (!ok) return false;
  t1 = m2_of(q1x, q1y, q1z, q1e);
  t2 = m2_of(q2x, q2y, q2z, q2e);
  return !(t1 > 0 || t2 > 0);
}

template <int NN, int K, class Draws>
__device__ __forceinline__ bool point_kinematics_CK(const double *u, Draws &dr, PointState<NN> &S, Ladder<K> &X) {
  const DevProc &P = c_proc;
  const double   pt1  = exp(P.log_pt0 + P.dlog_pt * u[0]);
  const double   pt2  = exp(P.log_pt0 + P.dlog_pt * u[1]);
  const double   phi1 = 2.0 * GRB_PI * u[2];
  const double   phi2 = 2.0 * GRB_PI * u[3];
#pragma unroll
  for (int i = 0; i < K; ++i) S.moff[i % NN] = offshell_mass(i, dr);
  double m1, m2;
  S.exc = sample_forward_masses(u[3 * K + 2], u[3 * K + 3], dr, m1, m2);
  S.M_MIN = 0.0; S.M_MAX = 0.0; S.DW = 0.0; S.qx = 0.0; S.qy = 0.0;
  double s1, c1, s2, c2;
  sincos(phi1, &s1, &c1);
  sincos(phi2, &s2, &c2);
  S.p1.x = pt1 * c1; S.p1.y = pt1 * s1;
  S.p2.x = pt2 * c2; S.p2.y = pt2 * s2;
  const double wx = S.p1.x + S.p2.x, wy = S.p1.y + S.p2.y;
  double bx[K], by[K];
  S.qprod = 1.0;
#pragma unroll
  for (int i = 0; i < K - 1; ++i) {
    const double kt = P.kt_min + P.dkt * u[4 + i];
    double       sk, ck;
    sincos(2.0 * GRB_PI * u[4 + (K - 1) + i], &sk, &ck);
    const double qx = kt * ck, qy = kt * sk;
    X.qx[i] = qx; X.qy[i] = qy;
    S.qprod *= msqrt(qx * qx + qy * qy);
    if (i == 0) { bx[0] = qx - wx; by[0] = qy - wy; }
    bx[i + 1] = qx * (-1.0) - wx; by[i + 1] = qy * (-1.0) - wy;
  }
  S.pX.x = 0.0; S.pX.y = 0.0; S.pX.z = 0.0; S.pX.e = 0.0;
  double sumM = 0.0;
#pragma unroll
  for (int i = 0; i < K; ++i) {
    v4 &p = S.np4[i % NN];
    p.x = blin_row<K>(i, bx);
    p.y = blin_row<K>(i, by);
    const double y  = P.rap_min + P.drap * u[4 + 2 * (K - 1) + i];
    const double mt = msqrt(pow2(S.moff[i % NN]) + pt2_of(p));
    p.z = mt * sinh(y); p.e = mt * cosh(y);
    S.pX.x += p.x; S.pX.y += p.y; S.pX.z += p.z; S.pX.e += p.e;
  }
  if (S.pX.e > P.sqrt_s) return false;
  PzConst pc;
  pc.s = P.s; pc.t7 = P.sqrt_s; pc.t9 = P.s_pow15; pc.t14 = P.s * P.s;
  const double p1z = solve_pz(m1, m2, pt1, pt2, S.pX.z, S.pX.e, pc);
  const double p2z = -(S.pX.z + p1z);
  if (p1z < 0 || p2z > 0 || isnan(p1z)) return false;
  S.p1.z = p1z; S.p1.e = msqrt(pow2(m1) + pow2(pt1) + pow2(p1z));
  S.p2.z = p2z; S.p2.e = msqrt(pow2(m2) + pow2(pt2) + pow2(p2z));
#pragma unroll
  for (int i = 0; i < K; ++i) sumM += m_of(S.np4[i % NN]);
  if (m_of(S.pX) < sumM) return false;
  {
    const double dx = 0.0 - ((S.p1.x + S.p2.x) + S.pX.x), dy = 0.0 - ((S.p1.y + S.p2.y) + S.pX.y);
    const double dz = P.bsum_z - ((S.p1.z + S.p2.z) + S.pX.z), de = P.bsum_e - ((S.p1.e + S.p2.e) + S.pX.e);
    if (fabs(dx) > 1e-6 || fabs(dy) > 1e-6 || fabs(dz) > 1e-6 || de > 1e-6) return false;
  }
  {
    v4 leg[K];
#pragma unroll
    for (int i = 0; i < K; ++i) leg[i] = S.np4[i % NN];
    if (!ladder_scalars<K>(S.p1, S.p2, S.pX, leg, S.t1, S.t2, S.s1, S.s2, X)) return false;
  }
  S.t_hat = 0.0; S.u_hat = 0.0; S.ss13 = X.s1a[0]; S.ss14 = X.s1a[1]; S.ss23 = X.s2a[0]; S.ss24 = X.s2a[1];
  S.m2 = m2_of(S.pX);
  S.Y  = rap_of(S.pX);
  S.Pt = msqrt(pt2_of(S.pX));
  return true;
}

// ff(x)^2-type pieces of the ladder: Meson_FF and Meson_prop (MRegge.cc:1155-1218) by run-time form-factor type
__device__ __forceinline__ double meson_ff1(double x, double M2) {
  const DevProc &P  = c_proc;
  const double   tp = x - M2;
  if (P.offshellFF == 2) return 1.0 / (1.0 - tp / P.b_POW);
  if (P.offshellFF == 0) return exp(-P.b_EXP * fabs(tp));
  return exp(-P.b_OREAR * msqrt(fabs(tp) + pow2(P.a_OREAR)) + P.a_OREAR * P.b_OREAR);
}
__device__ __forceinline__ cx regge_prop(double s, double t);
// reference src/MRegge.cc:470-540 ME6: sum over the 24 orderings (a, b, c, d) of the four legs (lexicographic order of
// math::Permutations) of  Prop(s_1a, t1) F_A ff(tt_a) g / (tt_a - M_A^2) ff_A(tt_ab) g Prop(s_bc, tt_ab) ff_B(tt_ab) g
// / (tt_d - M_B^2) ff(tt_d) g Prop(s_2d, t2) F_B  with tt_a = (q1 - p_a)^2, tt_ab = (q1 - p_a - p_b)^2, tt_d = (q2 - p_d)^2
__device__ __forceinline__ cx regge_me6(const double t1, const double t2, const Ladder4 &X) {
  const DevProc &P = c_proc;
  const double FA = P.gN_P * (1.0 / (1 - t1 / P.fc1)) * (1.0 / (1 - t1 / P.fc2));
  const double FB = P.gN_P * (1.0 / (1 - t2 / P.fc1)) * (1.0 / (1 - t2 / P.fc2));
  const double g  = P.gpp_P;
  cx P1[4], P2[4];
  double G1[4], G2[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    P1[a] = regge_prop(X.s1a[a], t1);
    P2[a] = regge_prop(X.s2a[a], t2);
    const double fa = meson_ff1(X.tt1[a], P.M2_leg0), fd = meson_ff1(X.tt2[a], P.M2_leg2);
    G1[a] = fa * g * (1.0 / (X.tt1[a] - P.M2_leg0));
    G2[a] = g * (1.0 / (X.tt2[a] - P.M2_leg2)) * fd * g;
  }
  cx A = {0.0, 0.0};
#pragma unroll 1
  for (int a = 0; a < 4; ++a)
#pragma unroll 1
    for (int b = 0; b < 4; ++b) {
      if (b == a) continue;
      const double tab = X.txy[pair4(a, b)];
      const double Fab = meson_ff1(tab, P.M2_leg0) * g, Fab2 = meson_ff1(tab, P.M2_leg2);
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        if (c == a || c == b) continue;
        const int d  = 6 - a - b - c;
        const cx  Pm = regge_prop(X.sab[pair4(b, c)], tab);
        // (((P1 FA G1 ...) Fab) Pm) Fab2 G2) P2 FB
        cx sub = {P1[a].re * FA * G1[a] * Fab, P1[a].im * FA * G1[a] * Fab};
        sub = {sub.re * Pm.re - sub.im * Pm.im, sub.re * Pm.im + sub.im * Pm.re};
        const double r = Fab2 * G2[d];
        sub = {sub.re * r, sub.im * r};
        sub = {(sub.re * P2[d].re - sub.im * P2[d].im) * FB, (sub.re * P2[d].im + sub.im * P2[d].re) * FB};
        A.re += sub.re;
        A.im += sub.im;
      }
    }
  return A;
}

// reference src/MRegge.cc:560-633 ME8: sum over the 720 orderings (a, b, c, d, e, f) of the six legs of
//   Prop(s_1a, t1) F_A ff_A(tt_a) g / (tt_a - M_A^2) . ff_A(tt_ab) g Prop(s_bc, tt_ab) ff_B(tt_ab) g . 1 / (tt_bc - M_B^2)
//   . ff_B(tt_cd) g Prop(s_de, tt_cd) ff_C(tt_cd) g . 1 / (tt_f - M_C^2) ff_C(tt_f) g Prop(s_2f, t2) F_B
// with the scalars the reference reads: tt_a = tt_1[a], tt_ab = tt_xy[a][b], tt_bc = tt_xy[b][c], tt_cd = tt_xy[c][d] (two-leg
// sums; no form factor on tt_bc), tt_f = tt_2[f].  The two middle blocks depend on three legs each (120 distinct values): the
// lower one, M2(c, d, e), is tabulated once (local memory), the upper one is formed in the (a, b, c) loops.
__device__ __forceinline__ cx regge_me8(const double t1, const double t2, const Ladder<6> &X) {
  const DevProc &P = c_proc;
  const double FA = P.gN_P * (1.0 / (1 - t1 / P.fc1)) * (1.0 / (1 - t1 / P.fc2));
  const double FB = P.gN_P * (1.0 / (1 - t2 / P.fc1)) * (1.0 / (1 - t2 / P.fc2));
  const double g  = P.gpp_P;
  cx U[6], V[6];
#pragma unroll 1
  for (int a = 0; a < 6; ++a) {
    const cx     P1 = regge_prop(X.s1a[a], t1), P2 = regge_prop(X.s2a[a], t2);
    const double ua = FA * meson_ff1(X.tt1[a], P.M2_leg0) * g * (1.0 / (X.tt1[a] - P.M2_leg0));
    const double vf = (1.0 / (X.tt2[a] - P.M2_leg4)) * meson_ff1(X.tt2[a], P.M2_leg4) * g * FB;
    U[a] = {P1.re * ua, P1.im * ua};
    V[a] = {P2.re * vf, P2.im * vf};
  }
  cx M2[120];  // index c * 20 + rank(d) * 4 + rank(e)
#pragma unroll 1
  for (int c = 0; c < 6; ++c)
#pragma unroll 1
    for (int d = 0; d < 6; ++d) {
      if (d == c) continue;
      const double tcd = X.txy[pairK<6>(c, d)];
      const double f   = meson_ff1(tcd, P.M2_leg2) * g * meson_ff1(tcd, P.M2_leg4) * g;
#pragma unroll 1
      for (int e = 0; e < 6; ++e) {
        if (e == c || e == d) continue;
        const cx Pm = regge_prop(X.sab[pairK<6>(d, e)], tcd);
        M2[c * 20 + (d - (d > c)) * 4 + (e - (e > c) - (e > d))] = {Pm.re * f, Pm.im * f};
      }
    }
  cx A = {0.0, 0.0};
#pragma unroll 1
  for (int a = 0; a < 6; ++a)
#pragma unroll 1
    for (int b = 0; b < 6; ++b) {
      if (b == a) continue;
      const double tab = X.txy[pairK<6>(a, b)];
      const double fab = meson_ff1(tab, P.M2_leg0) * g * meson_ff1(tab, P.M2_leg2) * g;
#pragma unroll 1
      for (int c = 0; c < 6; ++c) {
        if (c == a || c == b) continue;
        const cx     Pm = regge_prop(X.sab[pairK<6>(b, c)], tab);
        const double r  = fab * (1.0 / (X.txy[pairK<6>(b, c)] - P.M2_leg2));
        cx Y = {U[a].re * Pm.re - U[a].im * Pm.im, U[a].re * Pm.im + U[a].im * Pm.re};
        Y = {Y.re * r, Y.im * r};
        cx Z = {0.0, 0.0};   // sum over the orderings (d, e, f) of the three remaining legs
#pragma unroll 1
        for (int d = 0; d < 6; ++d) {
          if (d == a || d == b || d == c) continue;
#pragma unroll 1
          for (int e = 0; e < 6; ++e) {
            if (e == a || e == b || e == c || e == d) continue;
            const int f  = 15 - a - b - c - d - e;
            const cx  m  = M2[c * 20 + (d - (d > c)) * 4 + (e - (e > c) - (e > d))];
            Z.re += m.re * V[f].re - m.im * V[f].im;
            Z.im += m.re * V[f].im + m.im * V[f].re;
          }
        }
        A.re += Y.re * Z.re - Y.im * Z.im;
        A.im += Y.re * Z.im + Y.im * Z.re;
      }
    }
  return A;
}
template <int K>
__device__ __forceinline__ cx regge_ladder(const double t1, const double t2, const Ladder<K> &X);
template <>
__device__ __forceinline__ cx regge_ladder<4>(const double t1, const double t2, const Ladder<4> &X) { return regge_me6(t1, t2, X); }
template <>
__device__ __forceinline__ cx regge_ladder<6>(const double t1, const double t2, const Ladder<6> &X) { return regge_me8(t1, t2, X); }

// Screening loop of the ladders (reference MContinuum::LoopKinematics, src/MContinuum.cc:94-161, with Kf = 4, 6): the
// legs keep their masses and rapidities (re-derived from the saved 4-vectors), their transverse momenta follow the linear
// system of the shifted protons; ME6 / ME8 read only lts scalars, so every scalar is rebuilt from the updated momenta.
template <int K>
struct LoopInvK {
  double p1x, p1y, p2x, p2y, m1, m2, m1sq, m2sq;
  double qx[K - 1], qy[K - 1], m[K], sh[K], ch[K];
};
using LoopInv4 = LoopInvK<4>;
struct PzConst;
template <bool SAME, int K>
__device__ __forceinline__ bool loop_node_CK(const LoopInvK<K> &L, const PzConst &pc, double ktx, double kty, double &t1, double &t2, Ladder<K> &X);
template <bool SAME, int K>
__device__ __forceinline__ double screened_amp2_ladder(const LoopInvK<K> &L, const cx &h0);

// ---- cuts (reference src/MProcess.cc:1069-1175, src/MUserCuts.cc:21-75) ------------------------------------
__device__ __forceinline__ bool leaf_fid_ok(const v4 &p) {
  const DevProc &P    = c_proc;
  const double   pt2  = pt2_of(p);
  const double   pt   = msqrt(pt2);
  const double   et   = msqrt(m2_of(p) + pt2);
  const double   pmod = msqrt(p.x * p.x + p.y * p.y + p.z * p.z);
  const double   eta  = 0.5 * log((pmod + p.z) / (pmod - p.z));
  const double   rap  = rap_of(p);
  return pt >= P.f_pt_min && pt <= P.f_pt_max && et >= P.f_Et_min && et <= P.f_Et_max && eta >= P.f_eta_min &&
         eta <= P.f_eta_max && rap >= P.f_rap_min && rap <= P.f_rap_max;
}
__device__ __forceinline__ bool leaf_veto_hit(const v4 &p) {
  const DevProc &P    = c_proc;
  const double   pt   = msqrt(pt2_of(p));
  const double   pmod = msqrt(p.x * p.x + p.y * p.y + p.z * p.z);
  const double   eta  = 0.5 * log((pmod + p.z) / (pmod - p.z));
  bool           hit  = false;
  for (int i = 0; i < P.n_veto; ++i)
    if (pt >= P.veto[i].pt_min && pt <= P.veto[i].pt_max && eta >= P.veto[i].eta_min && eta <= P.veto[i].eta_max)
      hit = true;
  return hit;
}

template <int NN>
__device__ __forceinline__ bool fiducial_cuts(const PointState<NN> &S) {
  const DevProc &P = c_proc;
  if (!P.fid_active) return true;
  const int id = P.usercuts;
  if (id != 0) {
    if (id == 3 || id == -3) {
      const double dx = S.p1.x - S.p2.x, dy = S.p1.y - S.p2.y;
      const double dpt = msqrt(dx * dx + dy * dy);
      if (id == -3 && !(dpt < 0.3)) return false;
      if (id == 3 && !(dpt > 0.3)) return false;
    } else if (id == 111 || id == -111) {
      const double dot = S.p1.x * S.p2.x + S.p1.y * S.p2.y;
      if (id == -111 && !(dot < 0)) return false;
      if (id == 111 && !(dot > 0)) return false;
    }
  }
  if (!(fabs(S.t1) >= P.f_t_min && fabs(S.t1) <= P.f_t_max && fabs(S.t2) >= P.f_t_min && fabs(S.t2) <= P.f_t_max))
    return false;
  if (S.exc) {  // forward system mass of an excited leg, MProcess.cc:1105-1119
    if ((S.exc & 1) && !(m_of(S.p1) >= P.f_fM_min && m_of(S.p1) <= P.f_fM_max)) return false;
    if ((S.exc & 2) && !(m_of(S.p2) >= P.f_fM_min && m_of(S.p2) <= P.f_fM_max)) return false;
  }
  const double M = msqrt(S.m2);
  if (!(M >= P.f_M_min && M <= P.f_M_max && S.Y >= P.f_Y_min && S.Y <= P.f_Y_max && S.Pt >= P.f_Pt_min &&
        S.Pt <= P.f_Pt_max))
    return false;
  bool ok = true;
#pragma unroll
  for (int i = 0; i < NN; ++i)
    if (i < P.n_nodes && P.nodes[i].n_children == 0) ok = ok && leaf_fid_ok(S.np4[i]);
  return ok;
}
template <int NN>
__device__ __forceinline__ bool veto_cuts(const PointState<NN> &S) {
  const DevProc &P = c_proc;
  if (!P.veto_active) return true;
  bool ok = true;
#pragma unroll
  for (int i = 0; i < NN; ++i)
    if (i < P.n_nodes && P.nodes[i].n_children == 0 && leaf_veto_hit(S.np4[i])) ok = false;
  return ok;
}

// ---- Regge continuum amplitude (reference src/MRegge.cc:394-438 ME4) ---------------------------------------
// A = A_t + sign*A_u, A_x = Prop(s_a,t1) FF_A ff(x)^2 gpp^2 /(x - M^2) Prop(s_b,t2) FF_B with
// Prop(s,t) = sum_k eta_k(0) e^{-i pi/2 ap_k t} (s/s0)^{a0_k + ap_k t}.  With one active trajectory (TUNE0)
// the phase depends on t1+t2 only and the amplitude is  eta0^2 cis(pc (t1+t2)) * real.
struct NodeKin {
  double ss13, ss24, ss14, ss23, t1, t2, t_hat, u_hat;
  v4     p1, p2;  // outgoing protons lts.pfinal[1], [2] (needed by the tensor-pomeron vertices)
  v4     pX;      // central system lts.pfinal[0] (needed by the resonance amplitudes)
  int    exc;     // lts.excite1/2 (bits 0, 1)
};
typedef NodeKin ReggeIn;

// ff(x)^2 * prop(x) of MRegge.cc:1155-1218 as numerator / denominator (the caller divides once)
__device__ __forceinline__ void meson_G_nd(double x, double &num, double &den) {
  const DevProc &P  = c_proc;
  const double   tp = x - P.M2_leg0;
  if (P.offshellFF == 2) {
    const double d = fma(-tp, P.inv_b_POW, 1.0);  // 1 - tp/b
    num = 1.0;
    den = d * d * tp;
  } else if (P.offshellFF == 0) {
    const double f = exp(-P.b_EXP * fabs(tp));
    num = f * f;
    den = tp;
  } else {
    const double f = exp(-P.b_OREAR * msqrt(fabs(tp) + pow2(P.a_OREAR)) + P.a_OREAR * P.b_OREAR);
    num = f * f;
    den = tp;
  }
}
__device__ __forceinline__ double meson_G(double x) {
  double n, d;
  meson_G_nd(x, n, d);
  return n / d;
}

__device__ __forceinline__ cx regge_prop(double s, double t) {  // general PropOnly, MRegge.cc:888-906
  const DevProc &P = c_proc;
  cx             M = {0.0, 0.0};
  const double   ls = log(s) + P.ln_inv_s0;
#pragma unroll 1
  for (int k = 0; k < 3; ++k)
    if (P.regge_n[k]) {
      const double mag = exp((P.a0[k] + P.ap[k] * t) * ls);
      double       sn, cs;
      sincos(P.phase_c[k] * t, &sn, &cs);
      const double er = P.eta0_re[k] * cs - P.eta0_im[k] * sn, ei = P.eta0_re[k] * sn + P.eta0_im[k] * cs;
      M.re += er * mag;
      M.im += ei * mag;
    }
  return M;
}

// The amplitude algebra downstream of the kinematics is well conditioned and is organised for the fp64 pipe:
// one division per evaluation (all form-factor and propagator denominators are multiplied up first),
// contractions written as fma().
__device__ __forceinline__ cx regge_me4(const ReggeIn &in) {
  const DevProc &P = c_proc;
  // S3F(t1) S3F(t2) = 1 / [(1-t1/fc1)(1-t1/fc2)(1-t2/fc1)(1-t2/fc2)], MForm.cc:433
  const double dF = (fma(-in.t1, P.inv_fc1, 1.0) * fma(-in.t1, P.inv_fc2, 1.0)) *
                    (fma(-in.t2, P.inv_fc1, 1.0) * fma(-in.t2, P.inv_fc2, 1.0));
  double dt = in.t_hat - P.M2_leg0, du = in.u_hat - P.M2_leg0;  // meson propagator denominators, MRegge.cc:1199-1218
  cx     A;
  if (P.regge_n[0] && !P.regge_n[1] && !P.regge_n[2]) {
    // single trajectory: one sincos, two exp, four log (table-driven, grb_fastmath.cuh).  The off-shell meson form
    // factor ff(x)^2 = exp(gx) of the exponential / Orear parametrisations rides in the same exponent.
    const double a1 = fma(P.ap[0], in.t1, P.a0[0]), a2 = fma(P.ap[0], in.t2, P.a0[0]);
    double       gt = 0.0, gu = 0.0;
    if (P.offshellFF == 0) {
      gt = -2.0 * P.b_EXP * fabs(dt);
      gu = -2.0 * P.b_EXP * fabs(du);
    } else if (P.offshellFF == 1) {
      gt = 2.0 * fma(-P.b_OREAR, msqrt(fabs(dt) + pow2(P.a_OREAR)), P.a_OREAR * P.b_OREAR);
      gu = 2.0 * fma(-P.b_OREAR, msqrt(fabs(du) + pow2(P.a_OREAR)), P.a_OREAR * P.b_OREAR);
    } else {
      const double d1 = fma(-dt, P.inv_b_POW, 1.0), d2 = fma(-du, P.inv_b_POW, 1.0);
      dt *= d1 * d1;
      du *= d2 * d2;
    }
    double L13, L24, L14, L23, Et, Eu;
    fast_log4(in.ss13, in.ss24, in.ss14, in.ss23, L13, L24, L14, L23);
    fast_exp2(fma(a1, L13 + P.ln_inv_s0, fma(a2, L24 + P.ln_inv_s0, gt)), fma(a1, L14 + P.ln_inv_s0, fma(a2, L23 + P.ln_inv_s0, gu)), Et, Eu);
    // g2 (Et / dt + sign Eu / du) / dF
    const double R = div_checked(P.g2_const * fma(Et, du, P.me4_sign * Eu * dt), dt * du * dF);
    double       sn, cs;
    fast_sincos(P.phase_c[0] * (in.t1 + in.t2), sn, cs);
    A.re = fma(P.eta0sq_re, cs, -P.eta0sq_im * sn) * R;
    A.im = fma(P.eta0sq_re, sn, P.eta0sq_im * cs) * R;
  } else {
    double nt, nu;
    meson_G_nd(in.t_hat, nt, dt);
    meson_G_nd(in.u_hat, nu, du);
    nu *= P.me4_sign;
    const cx P13 = regge_prop(in.ss13, in.t1), P24 = regge_prop(in.ss24, in.t2);
    const cx P14 = regge_prop(in.ss14, in.t1), P23 = regge_prop(in.ss23, in.t2);
    const double tr = P13.re * P24.re - P13.im * P24.im, ti = P13.re * P24.im + P13.im * P24.re;
    const double ur = P14.re * P23.re - P14.im * P23.im, ui = P14.re * P23.im + P14.im * P23.re;
    const double g2 = P.g2_const / dF, Gt = nt / dt, Gu = nu / du;
    A.re = g2 * (tr * Gt + ur * Gu);
    A.im = g2 * (ti * Gt + ui * Gu);
  }
  if (P.excitation) {
    // PomPomProtonVertex (MRegge.cc:637-644): an excited leg carries sqrt(g3P gN_P) S3FINEL(t, M^2) instead of the
    // elastic gN_P S3F(t) that sits in g2_const / dF; S3FINEL = [|t| / (M^2 (|t| + a))]^((1 + Delta_P)/2), MForm.cc:443-449
    double f = 1.0;
    if (in.exc & 1) f *= P.exc_vertex * pow(fabs(in.t1) / (m2_of(in.p1) * (fabs(in.t1) + 0.5616)), 0.5 * (1 + 0.0808)) *
                         (fma(-in.t1, P.inv_fc1, 1.0) * fma(-in.t1, P.inv_fc2, 1.0));
    if (in.exc & 2) f *= P.exc_vertex * pow(fabs(in.t2) / (m2_of(in.p2) * (fabs(in.t2) + 0.5616)), 0.5 * (1 + 0.0808)) *
                         (fma(-in.t2, P.inv_fc1, 1.0) * fma(-in.t2, P.inv_fc2, 1.0));
    A.re *= f;
    A.im *= f;
  }
  return A;
}

// ---- tensor pomeron continuum, 2 x pseudoscalar (reference src/MTensorPomeron.cc:572-763) -----------------
// The reference contracts rank-2 vertices with rank-4 propagators through FTensor loops.  With
//   i Delta(s,t) = 1/(4s) (-i s a')^(alpha(t)-1) [g g + g g - 1/2 g g],   i Gamma_pp = -3i g_PNN F1(t) P P  (HE limit),
//   i Gamma_P,ps,ps(k',k) = -2i g F_M((k'-k)^2) [K K - 1/4 g K^2],  K = k' + k,
// every vertex.propagator.vertex chain collapses to  2 (P.K)^2 - 1/2 K^2 P^2  and the amplitude is
//   M = 36 g_PNN^2 g^2 F1(t1) F1(t2) F_M(t1) F_M(t2) e^{-i pi (a1+a2)/2}
//       [ E13 E24 R_t ff(t^)^2/(t^ - m^2) + E14 E23 R_u ff(u^)^2/(u^ - m^2) ],   E_ij = (s_ij a')^{a}/(4 s_ij),
// a_k = delta_P + a' t_k.  The four helicity-conserving proton combinations carry the same amplitude
// (:658-662,762), so one complex number per point is enough (|A|^2 = 4|M|^2/4).
__device__ __forceinline__ double mdot(const v4 &a, const v4 &b) { return a.e * b.e - (a.x * b.x + a.y * b.y + a.z * b.z); }

__device__ __forceinline__ double proton_F1(double t) {  // reference src/MForm.cc:542-548,607-645
  const DevProc &P   = c_proc;
  const double   Q2  = fabs(t);
  const double   tau = Q2 / P.four_mp2;
  double         GE, GM;
  if (P.em_kelly) {
    GE = (1.0 - 0.24 * tau) / (1.0 + 10.98 * tau + 12.82 * (tau * tau) + 21.97 * (tau * (tau * tau)));
    GM = 2.792847337 * (1.0 + 0.12 * tau) / (1.0 + 10.97 * tau + 18.86 * (tau * tau) + 6.55 * (tau * (tau * tau)));
  } else {
    GM = 2.792847337 / pow2(1.0 + Q2 / 0.71);
    GE = GM / 2.792847337;
  }
  return 1.0 / (tau + 1) * GE + tau / (tau + 1) * GM;
}

__device__ __forceinline__ cx tensor_me4_2xS(const NodeKin &in, const v4 &p3, const v4 &p4) {
  const DevProc &P = c_proc;
  v4 pa, pb;
  pa.x = 0.0; pa.y = 0.0; pa.z = P.pb1z; pa.e = P.pb1e;
  pb.x = 0.0; pb.y = 0.0; pb.z = P.pb2z; pb.e = P.pb2e;
  v4 P1, P2, pt, pu;
  P1.x = in.p1.x + pa.x; P1.y = in.p1.y + pa.y; P1.z = in.p1.z + pa.z; P1.e = in.p1.e + pa.e;
  P2.x = in.p2.x + pb.x; P2.y = in.p2.y + pb.y; P2.z = in.p2.z + pb.z; P2.e = in.p2.e + pb.e;
  pt.x = (pa.x - in.p1.x) - p3.x; pt.y = (pa.y - in.p1.y) - p3.y; pt.z = (pa.z - in.p1.z) - p3.z; pt.e = (pa.e - in.p1.e) - p3.e;
  pu.x = (p4.x - pa.x) + in.p1.x; pu.y = (p4.y - pa.y) + in.p1.y; pu.z = (p4.z - pa.z) + in.p1.z; pu.e = (p4.e - pa.e) + in.p1.e;
  v4 Kta, Ktb, Kua, Kub;
  Kta.x = pt.x - p3.x; Kta.y = pt.y - p3.y; Kta.z = pt.z - p3.z; Kta.e = pt.e - p3.e;
  Ktb.x = p4.x + pt.x; Ktb.y = p4.y + pt.y; Ktb.z = p4.z + pt.z; Ktb.e = p4.e + pt.e;
  Kua.x = p4.x + pu.x; Kua.y = p4.y + pu.y; Kua.z = p4.z + pu.z; Kua.e = p4.e + pu.e;
  Kub.x = pu.x - p3.x; Kub.y = pu.y - p3.y; Kub.z = pu.z - p3.z; Kub.e = pu.e - p3.e;
  const double P1sq = m2_of(P1), P2sq = m2_of(P2);
  const double d1t = mdot(P1, Kta), d2t = mdot(P2, Ktb), d1u = mdot(P1, Kua), d2u = mdot(P2, Kub);
  const double Rt = (2.0 * d1t * d1t - 0.5 * m2_of(Kta) * P1sq) * (2.0 * d2t * d2t - 0.5 * m2_of(Ktb) * P2sq);
  const double Ru = (2.0 * d1u * d1u - 0.5 * m2_of(Kua) * P1sq) * (2.0 * d2u * d2u - 0.5 * m2_of(Kub) * P2sq);
  const double pt2 = m2_of(pt), pu2 = m2_of(pu);
  const double a1 = P.tp_delta + P.tp_ap * in.t1, a2 = P.tp_delta + P.tp_ap * in.t2;
  const double Et = exp(a1 * (log(in.ss13) + P.tp_ln_ap) + a2 * (log(in.ss24) + P.tp_ln_ap)) / (16.0 * in.ss13 * in.ss24);
  const double Eu = exp(a1 * (log(in.ss14) + P.tp_ln_ap) + a2 * (log(in.ss23) + P.tp_ln_ap)) / (16.0 * in.ss14 * in.ss23);
  // F_M takes the pomeron momentum as the reference forms it at each vertex, (k'-k)^2 of iG_Ppsps(k',k):
  // (pt+p3)^2, (p4-pt)^2, (p4-pu)^2, (pu+p3)^2.  These equal t1/t2 only up to the O(1e-11) by which the
  // solved kinematics conserves 4-momentum, and dlnA/dt ~ 8/GeV^2 makes that visible at the 1e-10 level.
  const double qa_t = m2_of(pt.x + p3.x, pt.y + p3.y, pt.z + p3.z, pt.e + p3.e);
  const double qb_t = m2_of(p4.x - pt.x, p4.y - pt.y, p4.z - pt.z, p4.e - pt.e);
  const double qa_u = m2_of(p4.x - pu.x, p4.y - pu.y, p4.z - pu.z, p4.e - pu.e);
  const double qb_u = m2_of(pu.x + p3.x, pu.y + p3.y, pu.z + p3.z, pu.e + p3.e);
  const double FMt  = (1.0 / (1.0 - qa_t / 0.5)) * (1.0 / (1.0 - qb_t / 0.5));
  const double FMu  = (1.0 / (1.0 - qa_u / 0.5)) * (1.0 / (1.0 - qb_u / 0.5));
  const double pre  = P.tp_coupling * proton_F1(in.t1) * proton_F1(in.t2);
  const double R    = pre * (Et * Rt * FMt * meson_G(pt2) + Eu * Ru * FMu * meson_G(pu2));
  double       sn, cs;
  fast_sincos(P.tp_phase * (a1 + a2), sn, cs);  // tp_phase = -pi/2
  cx A;
  A.re = cs * R;
  A.im = sn * R;
  return A;
}

// ---- tensor pomeron continuum, 2 x vector meson (reference src/MTensorPomeron.cc:797-888) --------------------
// The reference builds rank-4 vertices iG_Pvv = i F_M [2 g1 Gamma0 - g2 Gamma2] (:2224-2238, 2500-2536) and contracts
// them with the proton vertex and the pomeron propagator through FTensor.  With X^{kl} = 2 P^k P^l - 1/2 P^2 g^{kl}
// (proton vertex x propagator, as in the 2xS case) the (k,l) contraction is done analytically:
//   Gamma0.X = S0 (k1.k2 g_uv - k2_u k1_v),                       S0 = 4 (P.k1)(P.k2) - P^2 (k1.k2)
//   Gamma2.X = 2 k1.k2 (2 P_u P_v - 1/2 P^2 g_uv) + g_uv S0 - 2 k1_v (2 P_u P.k2 - 1/2 P^2 k2_u)
//                                                        - 2 k2_u (2 P_v P.k1 - 1/2 P^2 k1_v)
// which leaves one real 4x4 block per vertex; the two blocks of a channel are joined by the Reggeized vector
// propagator -g_{r1 r2} (:2353-2368) and the result is contracted with the conjugated massive polarisation vectors
// (src/MDirac.cc:393-470), helicities (-1,0,1) x (-1,0,1).  Nine complex amplitudes; the four helicity-conserving
// proton combinations carry the same nine (:658-662), so sum_h |A_h|^2 / 4 = sum_9 |A|^2.
struct Lv { double c[4]; };  // covariant components [t,x,y,z]
__device__ __forceinline__ Lv lower(const v4 &a) { Lv r; r.c[0] = a.e; r.c[1] = -a.x; r.c[2] = -a.y; r.c[3] = -a.z; return r; }
__device__ __forceinline__ void pvv_block(const v4 &Pv, const v4 &k1, const v4 &k2, double g1, double g2, double a[4][4]) {
  const double Pk1 = mdot(Pv, k1), Pk2 = mdot(Pv, k2), P2 = m2_of(Pv), k12 = mdot(k1, k2);
  const double S0 = 4.0 * Pk1 * Pk2 - P2 * k12;
  const Lv     Pl = lower(Pv), k1l = lower(k1), k2l = lower(k2);
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const double guv = (u == v) ? (u == 0 ? 1.0 : -1.0) : 0.0;
      const double A0  = S0 * (k12 * guv - k2l.c[u] * k1l.c[v]);
      const double A2  = 2.0 * k12 * (2.0 * Pl.c[u] * Pl.c[v] - 0.5 * P2 * guv) + guv * S0 -
                        2.0 * k1l.c[v] * (2.0 * Pl.c[u] * Pk2 - 0.5 * P2 * k2l.c[u]) -
                        2.0 * k2l.c[u] * (2.0 * Pl.c[v] * Pk1 - 0.5 * P2 * k1l.c[v]);
      a[u][v] = 2.0 * g1 * A0 - g2 * A2;
    }
}
// conjugated polarisation vectors, contravariant [t,x,y,z]: e[m+1][mu]
__device__ __forceinline__ void eps_massive_conj(const v4 &k, cx e[3][4]) {
  const double pt2 = k.x * k.x + k.y * k.y, pt = msqrt(pt2), pm = msqrt(pt2 + k.z * k.z);
  const double ct = pm > 0.0 ? k.z / pm : 1.0, st = pm > 0.0 ? pt / pm : 0.0;
  const double cp = pt > 0.0 ? k.x / pt : 1.0, sp = pt > 0.0 ? k.y / pt : 0.0;
  const double N = 0.70710678118654752440, M = m_of(k), E = k.e;
  // m = -1: neg = +1, pos = -1;  m = +1: neg = -1, pos = +1;  conj flips the imaginary parts
  e[0][0] = {0.0, 0.0}; e[0][1] = {ct * cp * N, -sp * N}; e[0][2] = {ct * sp * N, cp * N}; e[0][3] = {-st * N, 0.0};
  e[2][0] = {0.0, 0.0}; e[2][1] = {-ct * cp * N, -sp * N}; e[2][2] = {-ct * sp * N, cp * N}; e[2][3] = {st * N, 0.0};
  e[1][0] = {pm / M, 0.0}; e[1][1] = {E * st * cp / M, 0.0}; e[1][2] = {E * st * sp / M, 0.0}; e[1][3] = {E * ct / M, 0.0};
}
__device__ __forceinline__ void tensor_me4_2xV(const NodeKin &in, const v4 &p3, const v4 &p4, cx h[9]) {
  const DevProc &P = c_proc;
  v4 pa, pb;
  pa.x = 0.0; pa.y = 0.0; pa.z = P.pb1z; pa.e = P.pb1e;
  pb.x = 0.0; pb.y = 0.0; pb.z = P.pb2z; pb.e = P.pb2e;
  v4 P1, P2, pt, pu, mpt, mpu;
  P1.x = in.p1.x + pa.x; P1.y = in.p1.y + pa.y; P1.z = in.p1.z + pa.z; P1.e = in.p1.e + pa.e;
  P2.x = in.p2.x + pb.x; P2.y = in.p2.y + pb.y; P2.z = in.p2.z + pb.z; P2.e = in.p2.e + pb.e;
  pt.x = (pa.x - in.p1.x) - p3.x; pt.y = (pa.y - in.p1.y) - p3.y; pt.z = (pa.z - in.p1.z) - p3.z; pt.e = (pa.e - in.p1.e) - p3.e;
  pu.x = (p4.x - pa.x) + in.p1.x; pu.y = (p4.y - pa.y) + in.p1.y; pu.z = (p4.z - pa.z) + in.p1.z; pu.e = (p4.e - pa.e) + in.p1.e;
  mpt.x = -pt.x; mpt.y = -pt.y; mpt.z = -pt.z; mpt.e = -pt.e;
  mpu.x = -pu.x; mpu.y = -pu.y; mpu.z = -pu.z; mpu.e = -pu.e;
  // iG_Pvv(prime, p): Gamma(k1 = prime, k2 = -p).  tA = (pt, -p3), tB = (p4, pt), uA = (p4, pu), uB = (pu, -p3)
  double atA[4][4], btB[4][4], auA[4][4], buB[4][4];
  pvv_block(P1, pt, p3, P.tp_g1, P.tp_g2, atA);
  pvv_block(P2, p4, mpt, P.tp_g1, P.tp_g2, btB);
  pvv_block(P1, p4, mpu, P.tp_g1, P.tp_g2, auA);
  pvv_block(P2, pu, p3, P.tp_g1, P.tp_g2, buB);
  // channel matrices M_t(r3,r4) = sum_r1 A(r1,r3) (-g_r1) B(r4,r1),  M_u(r3,r4) = sum_r1 A(r4,r1) (-g_r1) B(r1,r3)
  double mt[4][4], mu[4][4];
#pragma unroll
  for (int r3 = 0; r3 < 4; ++r3)
#pragma unroll
    for (int r4 = 0; r4 < 4; ++r4) {
      mt[r3][r4] = -atA[0][r3] * btB[r4][0] + atA[1][r3] * btB[r4][1] + atA[2][r3] * btB[r4][2] + atA[3][r3] * btB[r4][3];
      mu[r3][r4] = -auA[r4][0] * buB[0][r3] + auA[r4][1] * buB[1][r3] + auA[r4][2] * buB[2][r3] + auA[r4][3] * buB[3][r3];
    }
  // scalar factors: 9 gPNN^2 F1 F1 F_M F_M x pomeron propagators x Reggeized vector propagator x ff^2, then -i
  const double pt2 = m2_of(pt), pu2 = m2_of(pu);
  const double a1 = P.tp_delta + P.tp_ap * in.t1, a2 = P.tp_delta + P.tp_ap * in.t2;
  const double Et = exp(a1 * (log(in.ss13) + P.tp_ln_ap) + a2 * (log(in.ss24) + P.tp_ln_ap)) / (16.0 * in.ss13 * in.ss24);
  const double Eu = exp(a1 * (log(in.ss14) + P.tp_ln_ap) + a2 * (log(in.ss23) + P.tp_ln_ap)) / (16.0 * in.ss14 * in.ss23);
  const double qa_t = m2_of(pt.x + p3.x, pt.y + p3.y, pt.z + p3.z, pt.e + p3.e);
  const double qb_t = m2_of(p4.x - pt.x, p4.y - pt.y, p4.z - pt.z, p4.e - pt.e);
  const double qa_u = m2_of(p4.x - pu.x, p4.y - pu.y, p4.z - pu.z, p4.e - pu.e);
  const double qb_u = m2_of(pu.x + p3.x, pu.y + p3.y, pu.z + p3.z, pu.e + p3.e);
  const double FMt  = (1.0 / (1.0 - qa_t / 0.5)) * (1.0 / (1.0 - qb_t / 0.5));
  const double FMu  = (1.0 / (1.0 - qa_u / 0.5)) * (1.0 / (1.0 - qb_u / 0.5));
  const double pre  = P.tp_vv_coupling * proton_F1(in.t1) * proton_F1(in.t2);
  double       fft, ffu, dum;
  meson_G_nd(pt2, fft, dum);  // numerator = ff^2 (power law: 1 / dum-part handled below)
  meson_G_nd(pu2, ffu, dum);
  if (P.offshellFF == 2) {
    fft = 1.0 / pow2(fma(-(pt2 - P.M2_leg0), P.inv_b_POW, 1.0));
    ffu = 1.0 / pow2(fma(-(pu2 - P.M2_leg0), P.inv_b_POW, 1.0));
  }
  // iD_V: (e^{i phi} s34 / s_thr)^(alpha_R - 1) / (m2 - M0^2), alpha_R = 0.5 + 0.9 m2, phi = pi/2 (e^{(s_thr-s34)/s_thr} - 1)
  const double s34 = m2_of(in.pX), sthr = 4.0 * P.M2_leg0;
  const double phi = GRB_PI / 2.0 * exp((sthr - s34) / sthr) - GRB_PI / 2.0, lnr = log(s34 / sthr);
  const double bt = (0.5 + 0.9 * pt2) - 1.0, bu = (0.5 + 0.9 * pu2) - 1.0;
  const double Rt = pre * Et * FMt * fft * exp(bt * lnr) / (pt2 - P.M2_leg0);
  const double Ru = pre * Eu * FMu * ffu * exp(bu * lnr) / (pu2 - P.M2_leg0);
  // phases: pomerons e^{-i pi (a1+a2)/2}, vector propagator e^{i b phi}, overall -i
  double sn, cs;
  sincos(P.tp_phase * (a1 + a2) + bt * phi - GRB_PI / 2.0, &sn, &cs);
  const cx Ct = {Rt * cs, Rt * sn};
  sincos(P.tp_phase * (a1 + a2) + bu * phi - GRB_PI / 2.0, &sn, &cs);
  const cx Cu = {Ru * cs, Ru * sn};
  cx e3[3][4], e4[3][4];
  eps_massive_conj(p3, e3);
  eps_massive_conj(p4, e4);
#pragma unroll
  for (int h3 = 0; h3 < 3; ++h3) {
    // w_t[r4] = sum_r3 e3[r3] mt[r3][r4]
    cx wt[4], wu[4];
#pragma unroll
    for (int r4 = 0; r4 < 4; ++r4) {
      double tr = 0, ti = 0, ur = 0, ui = 0;
#pragma unroll
      for (int r3 = 0; r3 < 4; ++r3) {
        tr = fma(e3[h3][r3].re, mt[r3][r4], tr); ti = fma(e3[h3][r3].im, mt[r3][r4], ti);
        ur = fma(e3[h3][r3].re, mu[r3][r4], ur); ui = fma(e3[h3][r3].im, mu[r3][r4], ui);
      }
      wt[r4] = {tr, ti}; wu[r4] = {ur, ui};
    }
#pragma unroll
    for (int h4 = 0; h4 < 3; ++h4) {
      cx st = {0, 0}, su = {0, 0};
#pragma unroll
      for (int r4 = 0; r4 < 4; ++r4) {
        const cx &q = e4[h4][r4];
        st.re += wt[r4].re * q.re - wt[r4].im * q.im; st.im += wt[r4].re * q.im + wt[r4].im * q.re;
        su.re += wu[r4].re * q.re - wu[r4].im * q.im; su.im += wu[r4].re * q.im + wu[r4].im * q.re;
      }
      h[3 * h3 + h4] = {Ct.re * st.re - Ct.im * st.im + Cu.re * su.re - Cu.im * su.im,
                        Ct.re * st.im + Ct.im * st.re + Cu.re * su.im + Cu.im * su.re};
    }
  }
}
__device__ __forceinline__ double tensor_2xV_amp2(const NodeKin &in, const v4 &p3, const v4 &p4, cx &h0) {
  cx h[9];
  tensor_me4_2xV(in, p3, p4, h);
  double a2 = 0.0;
#pragma unroll
  for (int k = 0; k < 9; ++k) a2 += h[k].re * h[k].re + h[k].im * h[k].im;
  h0 = h[0];
  return a2;
}

// ---- tensor pomeron resonances (reference src/MTensorPomeron.cc:164-564 ME3; closed form in grb_tres.cuh) -------------
// common factor K of grb_tres.cuh: 9 g_PNN^2 F1(t1) F1(t2) F_M(t1) F_M(t2) (s1 a')^a1 (s2 a')^a2 / (16 s1 s2) e^{-i pi (a1+a2)/2}.
// s1, s2 = (p_1 + p_X)^2, (p_2 + p_X)^2 as lts.s1 / lts.s2 (MKinematics.h GetLorentzScalars); F_M takes q_k^2 = t_k.
__device__ __forceinline__ cx tensor_me3(const NodeKin &in, const v4 &p3, const v4 &p4) {
  const DevProc &P = c_proc;
  typedef tres::V4<double> T4;
  const T4 pa = {0.0, 0.0, P.pb1z, P.pb1e}, pb = {0.0, 0.0, P.pb2z, P.pb2e};
  const T4 p1 = {in.p1.x, in.p1.y, in.p1.z, in.p1.e}, p2 = {in.p2.x, in.p2.y, in.p2.z, in.p2.e};
  const T4 pX = {in.pX.x, in.pX.y, in.pX.z, in.pX.e}, d3 = {p3.x, p3.y, p3.z, p3.e}, d4 = {p4.x, p4.y, p4.z, p4.e};
  double   re, im;
  tres::resonance_sum<double>(c_res.tp, c_res.n, pa, pb, p1, p2, pX, d3, d4, re, im);
  const double s1 = m2_of(in.p1.x + in.pX.x, in.p1.y + in.pX.y, in.p1.z + in.pX.z, in.p1.e + in.pX.e);
  const double s2 = m2_of(in.p2.x + in.pX.x, in.p2.y + in.pX.y, in.p2.z + in.pX.z, in.p2.e + in.pX.e);
  const double a1 = P.tp_delta + P.tp_ap * in.t1, a2 = P.tp_delta + P.tp_ap * in.t2;
  const double E  = exp(a1 * (log(s1) + P.tp_ln_ap) + a2 * (log(s2) + P.tp_ln_ap)) / (16.0 * s1 * s2);
  const double K  = P.tp_vv_coupling * proton_F1(in.t1) * proton_F1(in.t2) * (1.0 / (1.0 - in.t1 / 1.0)) * (1.0 / (1.0 - in.t2 / 1.0)) * E;
  double       sn, cs;
  sincos(P.tp_phase * (a1 + a2), &sn, &cs);  // tp_phase = -pi/2
  cx A;
  A.re = K * (cs * re - sn * im);
  A.im = K * (cs * im + sn * re);
  return A;
}

// PP[RES] / PP[RES+CON]: ME4 + coherent resonance sum (reference PROC_9, PROC_17)
__device__ __forceinline__ cx regge_resonances(const NodeKin &in, const v4 &p3, const v4 &p4) {
  const DevProc &P = c_proc;
  ResKin K;
  K.t1 = in.t1; K.t2 = in.t2;
  K.pX = in.pX; K.p1 = in.p1; K.p2 = in.p2; K.d0 = p3; K.d1 = p4; K.exc = in.exc;
  K.s1 = m2_of(in.p1.x + in.pX.x, in.p1.y + in.pX.y, in.p1.z + in.pX.z, in.p1.e + in.pX.e);
  K.s2 = m2_of(in.p2.x + in.pX.x, in.p2.y + in.pX.y, in.p2.z + in.pX.z, in.p2.e + in.pX.e);
  K.m2 = m2_of(in.pX);
  K.q1.x = 0.0 - in.p1.x; K.q1.y = 0.0 - in.p1.y; K.q1.z = P.pb1z - in.p1.z; K.q1.e = P.pb1e - in.p1.e;
  K.x1 = 1 - in.p1.z / P.pb1z;
  K.x2 = 1 - in.p2.z / P.pb2z;
  K.qt1 = msqrt(K.q1.x * K.q1.x + K.q1.y * K.q1.y);
  K.qt2 = msqrt(in.p2.x * in.p2.x + in.p2.y * in.p2.y);
  return resonance_sum(K);
}

}  // namespace grb

#include "grb_yy.cuh"
#include "grb_ferm.cuh"

namespace grb {

enum { AMP_REGGE = 0, AMP_TENSOR_2XS = 1, AMP_REGGE_RES = 2, AMP_REGGE_RES_CON = 3, AMP_YY_LL = 4, AMP_TENSOR_2XV = 5,
       AMP_TENSOR_2XF = 6, AMP_YY_QED = 7, AMP_REGGE_ME6 = 8, AMP_REGGE_ME8 = 9, AMP_TENSOR_RES = 10, AMP_TENSOR_RES_CON = 11 };
template <int AMP>
__device__ __forceinline__ cx amplitude(const NodeKin &in, const v4 &p3, const v4 &p4) {
  if (AMP == AMP_YY_LL) return {sqrt(yy_ll_amp2(in, p3, p4)), 0.0};  // unscreened only: |A|^2 is all that is used
  if (AMP == AMP_TENSOR_2XV) {  // unscreened only: |A|^2 = sum over the nine helicity amplitudes
    cx h0;
    return {sqrt(tensor_2xV_amp2(in, p3, p4, h0)), 0.0};
  }
  if (AMP == AMP_TENSOR_2XF) return {sqrt(tensor_2xF_amp2(in, p3, p4)), 0.0};  // unscreened only
  if (AMP == AMP_YY_QED) return {sqrt(yy_qed_amp2(in, p3, p4)), 0.0};          // unscreened only
  if (AMP == AMP_REGGE_ME6 || AMP == AMP_REGGE_ME8) return {0.0, 0.0};  // ladders: evaluated from their own scalars (regge_ladder)
  if (AMP == AMP_TENSOR_2XS) return tensor_me4_2xS(in, p3, p4);
  if (AMP == AMP_TENSOR_RES) return tensor_me3(in, p3, p4);
  if (AMP == AMP_TENSOR_RES_CON) {  // PROC_14: helicity amplitudes of ME4 and ME3 added index by index (four equal copies each)
    const cx a = tensor_me4_2xS(in, p3, p4), b = tensor_me3(in, p3, p4);
    return {a.re + b.re, a.im + b.im};
  }
  if (AMP == AMP_REGGE_RES) return regge_resonances(in, p3, p4);
  if (AMP == AMP_REGGE_RES_CON) {
    const cx a = regge_me4(in), b = regge_resonances(in, p3, p4);
    return {a.re + b.re, a.im + b.im};
  }
  return regge_me4(in);
}

// Processes whose helicity amplitudes differ (reference lts.hamp.size() > 1 after dropping identical copies): the
// screening loop integrates every amplitude separately (MProcess.cc:262-283).  NH = distinct amplitudes,
// amp2 = SCALE * sum_h |A_h|^2:  yy -> l+ l- (HELAS and yy[QED]): 16 amplitudes, /4;  tensor 2xV: 36 = 4 x 9 amplitudes,
// /4 -> 9, x1;  tensor 2xF: 16 = 4 x 4 -> 4, x1.
template <int AMP> struct AmpHel { static constexpr int NH = 1; };
template <> struct AmpHel<AMP_YY_LL> { static constexpr int NH = 16; };
template <> struct AmpHel<AMP_TENSOR_2XV> { static constexpr int NH = 9; };
template <> struct AmpHel<AMP_TENSOR_2XF> { static constexpr int NH = 4; };
template <> struct AmpHel<AMP_YY_QED> { static constexpr int NH = 16; };
template <int AMP>
__device__ __forceinline__ double amp_hel_scale() { return (AMP == AMP_YY_LL || AMP == AMP_YY_QED) ? 0.25 : 1.0; }
template <int AMP>
__device__ __forceinline__ void amplitude_vec(const NodeKin &in, const v4 &p3, const v4 &p4, cx *h) {
  if (AMP == AMP_YY_LL) yy_ll_core<true>(in, p3, p4, h);
  else if (AMP == AMP_TENSOR_2XV) tensor_me4_2xV(in, p3, p4, h);
  else if (AMP == AMP_TENSOR_2XF) tensor_me4_2xF(in, p3, p4, h);
  else if (AMP == AMP_YY_QED) yy_qed_amp(in, p3, p4, h);
  else h[0] = amplitude<AMP>(in, p3, p4);
}

// ---- screening loop: per-point invariants and per-node work -------------------------------------------------
// reference src/MFactorized.cc:84-139 LoopKinematics: the loop re-derives m1, m2, mX, yX from the saved
// 4-vectors (pfinal_orig) -- done once per point here -- shifts the proton pT by -/+ kt, rebuilds pX and
// re-solves pz.  Quirk kept: the central daughters p3, p4 are NOT updated inside the loop.
struct LoopInv {
  double p1x, p1y, p2x, p2y;
  double m1, m2, m1sq, m2sq, mXsq, shY, chY, emax;  // emax = sqrt_s - (m1 + m2)
  v4     p3, p4;
  bool   same_mass;  // |m1 - m2| < 1e-9: SolvePz3_A branch (MKinematics.h:220-229)
  // <C> class (MContinuum::LoopKinematics): q0 and, per central leg, mass and sinh / cosh of the rapidity re-derived
  // from the saved 4-vectors (pfinal_orig[3], [4])
  double qx, qy, m3, m4, sh3, ch3, sh4, ch4;
  int    exc;  // lts.excite1/2 of the point
};

template <int NN>
__device__ __forceinline__ void make_loop_inv(const PointState<NN> &S, LoopInv &L) {
  L.p1x = S.p1.x; L.p1y = S.p1.y; L.p2x = S.p2.x; L.p2y = S.p2.y;
  L.m1 = m_of(S.p1); L.m2 = m_of(S.p2);
  L.m1sq = pow2(L.m1); L.m2sq = pow2(L.m2);
  const double mX = m_of(S.pX);
  const double yX = rap_of(S.pX);
  L.mXsq = pow2(mX);
  L.shY  = sinh(yX);
  L.chY  = cosh(yX);
  L.emax = c_proc.sqrt_s - (L.m1 + L.m2);
  L.same_mass = fabs(L.m1 - L.m2) < 1e-9;
  L.p3 = S.np4[0];
  L.p4 = S.np4[1];
  L.exc = S.exc;
  if (c_proc.ps_class == GRB_CLASS_C) {
    L.qx = S.qx; L.qy = S.qy;
    L.m3 = m_of(S.np4[0]); L.m4 = m_of(S.np4[1]);
    const double y3 = rap_of(S.np4[0]), y4 = rap_of(S.np4[1]);
    L.sh3 = sinh(y3); L.ch3 = cosh(y3); L.sh4 = sinh(y4); L.ch4 = cosh(y4);
  } else {
    L.qx = L.qy = L.m3 = L.m4 = L.sh3 = L.ch3 = L.sh4 = L.ch4 = 0.0;
  }
}

// One node of the loop.  The tests that can never fire are not evaluated: px, py, pz are conserved by
// construction (X = -(n1+n2), p2z = -(Xz+p1z)), and ss[0][1], ss[0][2] >= (m_p + m_X)^2; the energy test of
// CheckEMC, the direction test, ss13..ss24 >= 0 and t1, t2 <= 0 are kept.
template <bool SAME>
__device__ __forceinline__ bool loop_node(const LoopInv &L, const PzConst &pc, double ktx, double kty, NodeKin &o) {
  const DevProc &P = c_proc;
  const double n1x = L.p1x - ktx, n1y = L.p1y - kty, n2x = L.p2x + ktx, n2y = L.p2y + kty;
  const double Xx = -(n1x + n2x), Xy = -(n1y + n2y);
  // the three independent square roots share one range test (grb_fastmath.cuh sqrt3: same roundings as sqrt())
  double pt1, pt2, mtX;
  sqrt3(fmax(0.0, n1x * n1x + n1y * n1y), fmax(0.0, n2x * n2x + n2y * n2y), fmax(0.0, L.mXsq + (Xx * Xx + Xy * Xy)), pt1, pt2, mtX);
  const double Xz = mtX * L.shY, Xe = mtX * L.chY;
  if (Xe > L.emax) return false;
  // SolvePz picks its branch on |m1 - m2| < 1e-9 (MKinematics.h:220-229); m1, m2 are re-derived from the saved
  // 4-vectors (MFactorized.cc:97-98) whose E^2 - p^2 carries ~3e-9 of rounding noise, so ~2/3 of all points take
  // the unequal-mass branch.  The work list is partitioned by that flag so that a warp runs one branch only.
  const double p1z = SAME ? solve_pz3_A(L.m1, pt1, pt2, Xz, Xe, pc) : solve_pz3_B(L.m1, L.m2, pt1, pt2, Xz, Xe, pc);
  const double p2z = -(Xz + p1z);
  if (isnan(p1z) || p1z < 0 || p2z > 0) return false;
  double e1, e2;
  sqrt2(fmax(0.0, L.m1sq + pow2(pt1) + pow2(p1z)), fmax(0.0, L.m2sq + pow2(pt2) + pow2(p2z)), e1, e2);
  if (P.bsum_e - ((e1 + e2) + Xe) > 1e-6) return false;  // CheckEMC energy component (not abs, MMath.h:191-198)
  const v4 &p3 = L.p3, &p4 = L.p4;
  o.ss13 = m2_of(n1x + p3.x, n1y + p3.y, p1z + p3.z, e1 + p3.e);
  o.ss14 = m2_of(n1x + p4.x, n1y + p4.y, p1z + p4.z, e1 + p4.e);
  o.ss23 = m2_of(n2x + p3.x, n2y + p3.y, p2z + p3.z, e2 + p3.e);
  o.ss24 = m2_of(n2x + p4.x, n2y + p4.y, p2z + p4.z, e2 + p4.e);
  if (o.ss13 < 0 || o.ss14 < 0 || o.ss23 < 0 || o.ss24 < 0) return false;
  const double q1x = 0.0 - n1x, q1y = 0.0 - n1y, q1z = P.pb1z - p1z, q1e = P.pb1e - e1;
  o.t1 = m2_of(q1x, q1y, q1z, q1e);
  o.t2 = m2_of(0.0 - n2x, 0.0 - n2y, P.pb2z - p2z, P.pb2e - e2);
  if (o.t1 > 0 || o.t2 > 0) return false;
  o.t_hat = m2_of(q1x - p3.x, q1y - p3.y, q1z - p3.z, q1e - p3.e);
  o.u_hat = m2_of(q1x - p4.x, q1y - p4.y, q1z - p4.z, q1e - p4.e);
  o.p1.x = n1x; o.p1.y = n1y; o.p1.z = p1z; o.p1.e = e1;
  o.p2.x = n2x; o.p2.y = n2y; o.p2.z = p2z; o.p2.e = e2;
  o.pX.x = Xx; o.pX.y = Xy; o.pX.z = Xz; o.pX.e = Xe;
  o.exc = L.exc;
  return true;
}

// One node of the <C> class loop (reference src/MContinuum.cc:94-161): the central legs move with the protons (same
// masses and rapidities, transverse momenta from the linear system), so ss13..ss24 see the updated legs, while t_hat /
// u_hat and the amplitudes' vertex algebra keep reading the Born-level lts.decaytree[i].p4 (L.p3, L.p4) -- as the
// reference does.  No crude energy test here (:128-134 go straight to SolvePz).
template <bool SAME>
__device__ __forceinline__ bool loop_node_C(const LoopInv &L, const PzConst &pc, double ktx, double kty, NodeKin &o) {
  const DevProc &P = c_proc;
  const double n1x = L.p1x - ktx, n1y = L.p1y - kty, n2x = L.p2x + ktx, n2y = L.p2y + kty;
  const double wx = n1x + n2x, wy = n1y + n2y;
  const double p3x = (L.qx - wx) * (1.0 / 2.0), p3y = (L.qy - wy) * (1.0 / 2.0);
  const double p4x = (L.qx * (-1.0) - wx) * (1.0 / 2.0), p4y = (L.qy * (-1.0) - wy) * (1.0 / 2.0);
  const double mt3 = msqrt(pow2(L.m3) + (p3x * p3x + p3y * p3y)), mt4 = msqrt(pow2(L.m4) + (p4x * p4x + p4y * p4y));
  const double p3z = mt3 * L.sh3, p3e = mt3 * L.ch3, p4z = mt4 * L.sh4, p4e = mt4 * L.ch4;
  const double Xx = p3x + p4x, Xy = p3y + p4y, Xz = p3z + p4z, Xe = p3e + p4e;
  const double pt1 = msqrt(n1x * n1x + n1y * n1y), pt2 = msqrt(n2x * n2x + n2y * n2y);
  const double p1z = SAME ? solve_pz3_A(L.m1, pt1, pt2, Xz, Xe, pc) : solve_pz3_B(L.m1, L.m2, pt1, pt2, Xz, Xe, pc);
  const double p2z = -(Xz + p1z);
  if (isnan(p1z) || p1z < 0 || p2z > 0) return false;
  const double e1 = msqrt(L.m1sq + pow2(pt1) + pow2(p1z)), e2 = msqrt(L.m2sq + pow2(pt2) + pow2(p2z));
  {
    const double dx = 0.0 - ((n1x + n2x) + Xx), dy = 0.0 - ((n1y + n2y) + Xy);
    const double dz = P.bsum_z - ((p1z + p2z) + Xz), de = P.bsum_e - ((e1 + e2) + Xe);
    if (fabs(dx) > 1e-6 || fabs(dy) > 1e-6 || fabs(dz) > 1e-6 || de > 1e-6) return false;
  }
  o.ss13 = m2_of(n1x + p3x, n1y + p3y, p1z + p3z, e1 + p3e);
  o.ss14 = m2_of(n1x + p4x, n1y + p4y, p1z + p4z, e1 + p4e);
  o.ss23 = m2_of(n2x + p3x, n2y + p3y, p2z + p3z, e2 + p3e);
  o.ss24 = m2_of(n2x + p4x, n2y + p4y, p2z + p4z, e2 + p4e);
  const double s01 = m2_of(Xx + n1x, Xy + n1y, Xz + p1z, Xe + e1), s02 = m2_of(Xx + n2x, Xy + n2y, Xz + p2z, Xe + e2);
  const double s34 = m2_of(p3x + p4x, p3y + p4y, p3z + p4z, p3e + p4e);
  if (o.ss13 < 0 || o.ss14 < 0 || o.ss23 < 0 || o.ss24 < 0 || s01 < 0 || s02 < 0 || s34 < 0) return false;
  const double q1x = 0.0 - n1x, q1y = 0.0 - n1y, q1z = P.pb1z - p1z, q1e = P.pb1e - e1;
  o.t1 = m2_of(q1x, q1y, q1z, q1e);
  o.t2 = m2_of(0.0 - n2x, 0.0 - n2y, P.pb2z - p2z, P.pb2e - e2);
  if (o.t1 > 0 || o.t2 > 0) return false;
  const v4 &p3 = L.p3, &p4 = L.p4;
  o.t_hat = m2_of(q1x - p3.x, q1y - p3.y, q1z - p3.z, q1e - p3.e);
  o.u_hat = m2_of(q1x - p4.x, q1y - p4.y, q1z - p4.z, q1e - p4.e);
  o.p1.x = n1x; o.p1.y = n1y; o.p1.z = p1z; o.p1.e = e1;
  o.p2.x = n2x; o.p2.y = n2y; o.p2.z = p2z; o.p2.e = e2;
  o.pX.x = Xx; o.pX.y = Xy; o.pX.z = Xz; o.pX.e = Xe;
  o.exc = L.exc;
  return true;
}
template <bool SAME>
__device__ __forceinline__ bool loop_node_any(const LoopInv &L, const PzConst &pc, double ktx, double kty, NodeKin &o) {
  return c_proc.ps_class == GRB_CLASS_C ? loop_node_C<SAME>(L, pc, ktx, kty, o) : loop_node<SAME>(L, pc, ktx, kty, o);
}

// reference src/MProcess.cc:157-288 S3ScreenedAmp2 for a one-amplitude process (H = 1): returns
// |A_0 + i/(8 pi^2 s) * Simpson38_2D[ A_eik(kt^2) A(kt) kt ]|^2, nodes in the reference's row-major order
// (phi = 2 pi folded onto phi = 0, see grb_set_eikonal).  Rejected nodes contribute 0 but keep their weight.
template <int AMP, bool SAME>
__device__ __forceinline__ double screened_amp2_h1(const LoopInv &L, const cx &h0) {
  const DevLoop &Q = c_loop;
  PzConst        pc;
  pc.s = c_proc.s; pc.t7 = c_proc.sqrt_s; pc.t9 = c_proc.s_pow15; pc.t14 = c_proc.s * c_proc.s;
  double Ire = 0.0, Iim = 0.0;
  for (int i = 0; i < Q.nphi_nodes; ++i) {
    const double cp = Q.cphi[i], sp = Q.sphi[i], wi = Q.wphi[i];
    double       Jre = 0.0, Jim = 0.0;
    for (int j = 0; j <= Q.nkt; ++j) {
      const double kt = Q.kt[j];
      NodeKin      in;
      if (!loop_node_any<SAME>(L, pc, kt * cp, kt * sp, in)) continue;
      const cx A = amplitude<AMP>(in, L.p3, L.p4);
      // f = A_eik * A * kt * w_j, with b_j = A_eik(kt_j^2) kt_j w_j tabulated
      Jre = fma(Q.b_re[j], A.re, fma(-Q.b_im[j], A.im, Jre));
      Jim = fma(Q.b_re[j], A.im, fma(Q.b_im[j], A.re, Jim));
    }
    Ire = fma(wi, Jre, Ire);
    Iim = fma(wi, Jim, Iim);
  }
  // hamp_loop = i * norm * I
  const double lr = -Q.norm * Iim, li = Q.norm * Ire;
  const double ar = h0.re + lr, ai = h0.im + li;
  return ar * ar + ai * ai;
}

// ---- single-trajectory Regge continuum: one loop node as ONE basic block ----------------------------------------------
// loop_node() + regge_me4() leave a node through ~20 branches (range guards of every sqrt / division / log / exp /
// sincos, five physics rejections), which costs issue slots and keeps the scheduler from overlapping the independent
// chains.  Here every MUFU-seeded sequence runs unconditionally: `valid` collects the physics rejections, the return value
// says whether a range guard failed on a node that was still valid at that point -- then (rare) the caller re-evaluates
// the node on the checked path.  Same expressions in the same order as the checked path: identical bits when the guards
// hold.  FF = PARAM_REGGE::offshellFF resolved at compile time.
// S0ONE: PARAM_REGGE::s0 = 1 GeV^2 (TUNE0), ln(1/s0) = 0 exactly: the four additions of it are identities and are skipped
template <bool SAME, int FF, bool S0ONE>
__device__ __forceinline__ bool regge_node_fast(const LoopInv &L, const PzInv &pzi, double ktx, double kty, bool &valid, cx &A) {
  const DevProc &P = c_proc;
  const double n1x = L.p1x - ktx, n1y = L.p1y - kty, n2x = L.p2x + ktx, n2y = L.p2y + kty;
  const double Xx = -(n1x + n2x), Xy = -(n1y + n2y);
  // msqrt's clamp max(0, .) is dropped on this path: every radicand is a sum of squares (>= 0, or NaN -- and a NaN
  // fails the range guard, which sends the node to the checked path)
  const double a1s = n1x * n1x + n1y * n1y, a2s = n2x * n2x + n2y * n2y, aXs = L.mXsq + (Xx * Xx + Xy * Xy);
  // Range guards: every MUFU-seeded square root (and every table-driven log) needs a normal positive argument >= 2^-970;
  // their keys (fast_key) are min-combined and tested ONCE at the end, the other guards are and-combined without
  // short-circuit evaluation: no branch, no convergence barrier.  A failed guard sends the node to the checked path
  // whether or not the node is still valid at that point (identical result either way; guards fail on a handful of nodes
  // per million, essentially only where the kinematics has no solution).
  unsigned     key = min(min(fast_key(a1s), fast_key(a2s)), fast_key(aXs));
  const double pt1 = sqrt_fast_core(a1s), pt2 = sqrt_fast_core(a2s), mtX = sqrt_fast_core(aXs);
  const double Xz = mtX * L.shY, Xe = mtX * L.chY;
  valid = !(Xe > L.emax);
  bool         dom = true;
  const double p1z = SAME ? solve_pz3_A_fast(pzi, pt1, pt2, Xz, Xe, key, dom) : solve_pz3_B_fast(pzi, pt1, pt2, Xz, Xe, key, dom);
  const double p2z = -(Xz + p1z);
  valid = valid & !(isnan(p1z) | (p1z < 0) | (p2z > 0));
  const double e1s = L.m1sq + pow2(pt1) + pow2(p1z), e2s = L.m2sq + pow2(pt2) + pow2(p2z);
  key = min(key, min(fast_key(e1s), fast_key(e2s)));
  const double e1 = sqrt_fast_core(e1s), e2 = sqrt_fast_core(e2s);
  valid = valid & !(P.bsum_e - ((e1 + e2) + Xe) > 1e-6);
  // (Measured and not taken: the sub-energies through dot products, (a + b)^2 = m_a^2 + m_b^2 + 2 a.b -- 2.4 % faster, but
  // the reference's (E1 + E3)^2 - |p1 + p3|^2 carries ~4e-13 of relative rounding noise that such a form does not
  // reproduce: screened weights moved from a median 4e-15 to 7e-13 against the reference.)
  //
  // M4Vec::M2 is e^2 - ((x^2 + y^2) + z^2).  The transverse parts are shared bit for bit between the scalars:
  //   t1:  q1 = p_beam1 - p1 has q1x = 0 - n1x, so q1x^2 + q1y^2 = n1x^2 + n1y^2 = a1s (negation is exact); t2 likewise;
  //   t^:  (q1x - p3x) = -(n1x + p3x) exactly, so its transverse part is the one of ss13 = (p1 + p3)^2; u^ <-> ss14.
  // Same products, same sums, same order: identical bits, 20 fp64 instructions fewer per node.
  const v4 &p3 = L.p3, &p4 = L.p4;
  const double x13 = n1x + p3.x, y13 = n1y + p3.y, x14 = n1x + p4.x, y14 = n1y + p4.y;
  const double x23 = n2x + p3.x, y23 = n2y + p3.y, x24 = n2x + p4.x, y24 = n2y + p4.y;
  const double T13 = x13 * x13 + y13 * y13, T14 = x14 * x14 + y14 * y14, T23 = x23 * x23 + y23 * y23, T24 = x24 * x24 + y24 * y24;
  const double ss13 = pow2(e1 + p3.e) - (T13 + pow2(p1z + p3.z));
  const double ss14 = pow2(e1 + p4.e) - (T14 + pow2(p1z + p4.z));
  const double ss23 = pow2(e2 + p3.e) - (T23 + pow2(p2z + p3.z));
  const double ss24 = pow2(e2 + p4.e) - (T24 + pow2(p2z + p4.z));
  // ss < 0 for any of the four <=> a sign bit is set (a difference of equal numbers is +0, never -0): one integer test
  valid = valid & ((__double2hiint(ss13) | __double2hiint(ss14) | __double2hiint(ss23) | __double2hiint(ss24)) >= 0);
  const double q1z = P.pb1z - p1z, q1e = P.pb1e - e1;
  const double t1 = pow2(q1e) - (a1s + pow2(q1z));
  const double t2 = pow2(P.pb2e - e2) - (a2s + pow2(P.pb2z - p2z));
  valid = valid & !(t1 > 0);
  valid = valid & !(t2 > 0);
  const double t_hat = pow2(q1e - p3.e) - (T13 + pow2(q1z - p3.z));
  const double u_hat = pow2(q1e - p4.e) - (T14 + pow2(q1z - p4.z));
  // amplitude: regge_me4(), single-trajectory branch
  const double dF = (fma(-t1, P.inv_fc1, 1.0) * fma(-t1, P.inv_fc2, 1.0)) * (fma(-t2, P.inv_fc1, 1.0) * fma(-t2, P.inv_fc2, 1.0));
  double       dt = t_hat - P.M2_leg0, du = u_hat - P.M2_leg0;
  const double al1 = fma(P.ap[0], t1, P.a0[0]), al2 = fma(P.ap[0], t2, P.a0[0]);
  double       gt = 0.0, gu = 0.0;
  if (FF == 0) {
    gt = -2.0 * P.b_EXP * fabs(dt);
    gu = -2.0 * P.b_EXP * fabs(du);
  } else if (FF == 1) {
    gt = 2.0 * fma(-P.b_OREAR, msqrt(fabs(dt) + pow2(P.a_OREAR)), P.a_OREAR * P.b_OREAR);
    gu = 2.0 * fma(-P.b_OREAR, msqrt(fabs(du) + pow2(P.a_OREAR)), P.a_OREAR * P.b_OREAR);
  } else {
    const double d1 = fma(-dt, P.inv_b_POW, 1.0), d2 = fma(-du, P.inv_b_POW, 1.0);
    dt *= d1 * d1;
    du *= d2 * d2;
  }
  key = min(min(key, fast_key(ss13)), min(min(fast_key(ss24), fast_key(ss14)), fast_key(ss23)));
  const double L13 = fast_log_core(ss13), L24 = fast_log_core(ss24), L14 = fast_log_core(ss14), L23 = fast_log_core(ss23);
  const double c0 = S0ONE ? 0.0 : P.ln_inv_s0;
  const double xt = S0ONE ? fma(al1, L13, fma(al2, L24, gt)) : fma(al1, L13 + c0, fma(al2, L24 + c0, gt));
  const double xu = S0ONE ? fma(al1, L14, fma(al2, L23, gu)) : fma(al1, L14 + c0, fma(al2, L23 + c0, gu));
  const double Et = fast_exp_core(xt), Eu = fast_exp_core(xu);
  const double num = P.g2_const * fma(Et, du, P.me4_sign * Eu * dt), den = dt * du * dF;
  const double R   = div_fast_core(num, den);
  const double ph  = P.phase_c[0] * (t1 + t2);
  const bool redo = !(fast_key_ok(key) & dom & (fabs(xt) < 690.0) & (fabs(xu) < 690.0) & div_fast_ok(num, R) & (fabs(ph) < 1.0e5));
  double sn, cs;
  fast_sincos_core(ph, sn, cs);
  A.re = fma(P.eta0sq_re, cs, -P.eta0sq_im * sn) * R;
  A.im = fma(P.eta0sq_re, sn, P.eta0sq_im * cs) * R;
  return redo;
}
// the checked path of one node (taken when a range guard failed: a handful of nodes per million)
template <bool SAME>
__device__ __forceinline__ bool regge_node_checked(const LoopInv &L, const PzConst &pc, double ktx, double kty, cx &A) {
  NodeKin in;
  if (!loop_node<SAME>(L, pc, ktx, kty, in)) return false;
  A = regge_me4(in);
  return true;
}
template <bool SAME, int FF, bool S0ONE>
__device__ __forceinline__ double screened_amp2_regge_fast(const LoopInv &L, const cx &h0) {
  const DevLoop &Q = c_loop;
  PzConst        pc;
  pc.s = c_proc.s; pc.t7 = c_proc.sqrt_s; pc.t9 = c_proc.s_pow15; pc.t14 = c_proc.s * c_proc.s;
  const PzInv pzi = make_pz_inv(L.m1, L.m2, pc, SAME);
  double Ire = 0.0, Iim = 0.0;
  for (int i = 0; i < Q.nphi_nodes; ++i) {
    const double cp = Q.cphi[i], sp = Q.sphi[i], wi = Q.wphi[i];
    double       Jre = 0.0, Jim = 0.0;
    for (int j = 0; j <= Q.nkt; ++j) {
      const double kt = Q.kt[j];
      bool         valid;
      cx           A;
      if (__builtin_expect(regge_node_fast<SAME, FF, S0ONE>(L, pzi, kt * cp, kt * sp, valid, A), 0)) valid = regge_node_checked<SAME>(L, pc, kt * cp, kt * sp, A);
      if (valid) {  // rejected nodes contribute 0 but keep their Simpson weight (predicated, no branch)
        Jre = fma(Q.b_re[j], A.re, fma(-Q.b_im[j], A.im, Jre));
        Jim = fma(Q.b_re[j], A.im, fma(Q.b_im[j], A.re, Jim));
      }
    }
    Ire = fma(wi, Jre, Ire);
    Iim = fma(wi, Jim, Iim);
  }
  const double lr = -Q.norm * Iim, li = Q.norm * Ire;
  const double ar = h0.re + lr, ai = h0.im + li;
  return ar * ar + ai * ai;
}

// ---- <C> class, single-trajectory Regge continuum: the one-basic-block node with the moving central legs of
// loop_node_C (two more square roots for the leg transverse masses, no crude energy test, full CheckEMC) ---------------
template <bool SAME, int FF>
__device__ __forceinline__ bool regge_node_fast_C(const LoopInv &L, const PzInv &pzi, double ktx, double kty, bool &valid, cx &A) {
  const DevProc &P = c_proc;
  const double n1x = L.p1x - ktx, n1y = L.p1y - kty, n2x = L.p2x + ktx, n2y = L.p2y + kty;
  const double wx = n1x + n2x, wy = n1y + n2y;
  const double p3x = (L.qx - wx) * (1.0 / 2.0), p3y = (L.qy - wy) * (1.0 / 2.0);
  const double p4x = (L.qx * (-1.0) - wx) * (1.0 / 2.0), p4y = (L.qy * (-1.0) - wy) * (1.0 / 2.0);
  const double a3s = pow2(L.m3) + (p3x * p3x + p3y * p3y), a4s = pow2(L.m4) + (p4x * p4x + p4y * p4y);
  const double a1s = n1x * n1x + n1y * n1y, a2s = n2x * n2x + n2y * n2y;
  bool         redo = !(sqrt_fast_ok(a1s) && sqrt_fast_ok(a2s) && sqrt_fast_ok(a3s) && sqrt_fast_ok(a4s));
  const double mt3 = sqrt_fast_core(a3s), mt4 = sqrt_fast_core(a4s), pt1 = sqrt_fast_core(a1s), pt2 = sqrt_fast_core(a2s);
  const double p3z = mt3 * L.sh3, p3e = mt3 * L.ch3, p4z = mt4 * L.sh4, p4e = mt4 * L.ch4;
  const double Xx = p3x + p4x, Xy = p3y + p4y, Xz = p3z + p4z, Xe = p3e + p4e;
  bool         dom = true;
  const double p1z = SAME ? solve_pz3_A_fast(pzi, pt1, pt2, Xz, Xe, dom) : solve_pz3_B_fast(pzi, pt1, pt2, Xz, Xe, dom);
  redo = redo || !dom;
  const double p2z = -(Xz + p1z);
  valid = !(isnan(p1z) || p1z < 0 || p2z > 0);
  const double e1s = L.m1sq + pow2(pt1) + pow2(p1z), e2s = L.m2sq + pow2(pt2) + pow2(p2z);
  redo = redo || (valid && !(sqrt_fast_ok(e1s) && sqrt_fast_ok(e2s)));
  const double e1 = sqrt_fast_core(e1s), e2 = sqrt_fast_core(e2s);
  {
    const double dx = 0.0 - ((n1x + n2x) + Xx), dy = 0.0 - ((n1y + n2y) + Xy);
    const double dz = P.bsum_z - ((p1z + p2z) + Xz), de = P.bsum_e - ((e1 + e2) + Xe);
    valid = valid && !(fabs(dx) > 1e-6 || fabs(dy) > 1e-6 || fabs(dz) > 1e-6 || de > 1e-6);
  }
  const double ss13 = m2_of(n1x + p3x, n1y + p3y, p1z + p3z, e1 + p3e);
  const double ss14 = m2_of(n1x + p4x, n1y + p4y, p1z + p4z, e1 + p4e);
  const double ss23 = m2_of(n2x + p3x, n2y + p3y, p2z + p3z, e2 + p3e);
  const double ss24 = m2_of(n2x + p4x, n2y + p4y, p2z + p4z, e2 + p4e);
  const double s01 = m2_of(Xx + n1x, Xy + n1y, Xz + p1z, Xe + e1), s02 = m2_of(Xx + n2x, Xy + n2y, Xz + p2z, Xe + e2);
  const double s34 = m2_of(p3x + p4x, p3y + p4y, p3z + p4z, p3e + p4e);
  valid = valid && !(ss13 < 0 || ss14 < 0 || ss23 < 0 || ss24 < 0 || s01 < 0 || s02 < 0 || s34 < 0);
  const double q1x = 0.0 - n1x, q1y = 0.0 - n1y, q1z = P.pb1z - p1z, q1e = P.pb1e - e1;
  const double t1 = m2_of(q1x, q1y, q1z, q1e);
  const double t2 = m2_of(0.0 - n2x, 0.0 - n2y, P.pb2z - p2z, P.pb2e - e2);
  valid = valid && !(t1 > 0 || t2 > 0);
  const v4 &p3 = L.p3, &p4 = L.p4;   // Born-level lts.decaytree[i].p4 (not updated in the reference's loop)
  const double t_hat = m2_of(q1x - p3.x, q1y - p3.y, q1z - p3.z, q1e - p3.e);
  const double u_hat = m2_of(q1x - p4.x, q1y - p4.y, q1z - p4.z, q1e - p4.e);
  // amplitude: regge_me4(), single-trajectory branch (as in regge_node_fast)
  const double dF = (fma(-t1, P.inv_fc1, 1.0) * fma(-t1, P.inv_fc2, 1.0)) * (fma(-t2, P.inv_fc1, 1.0) * fma(-t2, P.inv_fc2, 1.0));
  double       dt = t_hat - P.M2_leg0, du = u_hat - P.M2_leg0;
  const double al1 = fma(P.ap[0], t1, P.a0[0]), al2 = fma(P.ap[0], t2, P.a0[0]);
  double       gt = 0.0, gu = 0.0;
  if (FF == 0) {
    gt = -2.0 * P.b_EXP * fabs(dt);
    gu = -2.0 * P.b_EXP * fabs(du);
  } else if (FF == 1) {
    gt = 2.0 * fma(-P.b_OREAR, msqrt(fabs(dt) + pow2(P.a_OREAR)), P.a_OREAR * P.b_OREAR);
    gu = 2.0 * fma(-P.b_OREAR, msqrt(fabs(du) + pow2(P.a_OREAR)), P.a_OREAR * P.b_OREAR);
  } else {
    const double d1 = fma(-dt, P.inv_b_POW, 1.0), d2 = fma(-du, P.inv_b_POW, 1.0);
    dt *= d1 * d1;
    du *= d2 * d2;
  }
  redo = redo || (valid && !(fast_log_in_domain(ss13) && fast_log_in_domain(ss24) && fast_log_in_domain(ss14) && fast_log_in_domain(ss23)));
  const double L13 = fast_log_core(ss13), L24 = fast_log_core(ss24), L14 = fast_log_core(ss14), L23 = fast_log_core(ss23);
  const double xt = fma(al1, L13 + P.ln_inv_s0, fma(al2, L24 + P.ln_inv_s0, gt)), xu = fma(al1, L14 + P.ln_inv_s0, fma(al2, L23 + P.ln_inv_s0, gu));
  const double Et = fast_exp_core(xt), Eu = fast_exp_core(xu);
  const double num = P.g2_const * fma(Et, du, P.me4_sign * Eu * dt), den = dt * du * dF;
  const double R   = div_fast_core(num, den);
  const double ph  = P.phase_c[0] * (t1 + t2);
  redo = redo || (valid && !(fabs(xt) < 690.0 && fabs(xu) < 690.0 && div_fast_ok(num, R) && fabs(ph) < 1.0e5));
  double sn, cs;
  fast_sincos_core(ph, sn, cs);
  A.re = fma(P.eta0sq_re, cs, -P.eta0sq_im * sn) * R;
  A.im = fma(P.eta0sq_re, sn, P.eta0sq_im * cs) * R;
  return redo;
}
template <bool SAME, int FF>
__device__ __forceinline__ double screened_amp2_regge_fast_C(const LoopInv &L, const cx &h0) {
  const DevLoop &Q = c_loop;
  PzConst        pc;
  pc.s = c_proc.s; pc.t7 = c_proc.sqrt_s; pc.t9 = c_proc.s_pow15; pc.t14 = c_proc.s * c_proc.s;
  const PzInv pzi = make_pz_inv(L.m1, L.m2, pc, SAME);
  double Ire = 0.0, Iim = 0.0;
  for (int i = 0; i < Q.nphi_nodes; ++i) {
    const double cp = Q.cphi[i], sp = Q.sphi[i], wi = Q.wphi[i];
    double       Jre = 0.0, Jim = 0.0;
    for (int j = 0; j <= Q.nkt; ++j) {
      const double kt = Q.kt[j];
      bool         valid;
      cx           A;
      if (__builtin_expect(regge_node_fast_C<SAME, FF>(L, pzi, kt * cp, kt * sp, valid, A), 0)) {
        NodeKin in;
        valid = loop_node_C<SAME>(L, pc, kt * cp, kt * sp, in);
        if (valid) A = regge_me4(in);
      }
      if (valid) {
        Jre = fma(Q.b_re[j], A.re, fma(-Q.b_im[j], A.im, Jre));
        Jim = fma(Q.b_re[j], A.im, fma(Q.b_im[j], A.re, Jim));
      }
    }
    Ire = fma(wi, Jre, Ire);
    Iim = fma(wi, Jim, Iim);
  }
  const double ar = h0.re - Q.norm * Iim, ai = h0.im + Q.norm * Ire;
  return ar * ar + ai * ai;
}

// ---- tensor pomeron 2xS: the same one-basic-block treatment ------------------------------------------------------------
// Kinematics as in regge_node_fast (kept as its own copy so that the headline kernel's code generation is untouched);
// the amplitude of tensor_me4_2xS() with table-driven log / exp / sincos and ALL its quotients -- two 1/(16 s s), four
// F_M, two meson propagators, ten inside the two Kelly form factors F1(t) -- multiplied up into one division:
//   R = pre (Nt Du + Nu Dt) / (Dt Du D_F1),  term_x = E'_x R_x num_x / [16 s s (1 - 2 qa)(1 - 2 qb) den_x].
// The vertex algebra runs on Lorentz products instead of the reference's component vectors: with q1 = p_a - p_1 (pomeron 1),
// P1 = p_a + p_1, P2 = p_b + p_2 and the central legs p3, p4 (fixed over the loop in the <F> class)
//   p_t = q1 - p3, p_u = p4 - q1, K_ta = q1 - 2 p3, K_tb = q1 + p4 - p3, K_ua = 2 p4 - q1, K_ub = p4 - p3 - q1,
// so every K^2, P.K, p_t^2, p_u^2 is a combination of t1 = q1^2, the ten products q1.p3, q1.p4, P1.p3, P1.p4, P2.p3, P2.p4,
// P1.q1, P2.q1, P1^2, P2^2 (fused dot products: 4 instructions each instead of 4 + 7 for a component vector and its M2) and
// the per-point constants m3^2, m4^2, p3.p4.  The F_M arguments are (p_t + p3)^2 = (p4 - p_u)^2 = q1^2 and (p4 - p_t)^2 =
// (p_u + p3)^2 = (p3 + p4 - q1)^2 (NOT replaced by t2: the solved kinematics conserves momentum only to O(1e-11), see
// tensor_me4_2xS).  None of these regroupings cancels more than the reference's own forms do (the products are taken with
// the small components of q1 directly, never as differences of beam-sized products); the result differs from
// tensor_me4_2xS by ~1e-13 relative (test_fast_node_equals_checked_node), the polynomial form factors are in Horner form.
// 220 -> 170 fp64 instructions for the vertex algebra.
struct TensorInv { double m3sq, m4sq, p3p4; };
__device__ __forceinline__ double fdot(double ae, double ax, double ay, double az, const v4 &b) {  // a.b, fused
  return fma(ae, b.e, -fma(ax, b.x, fma(ay, b.y, az * b.z)));
}
template <bool SAME>
__device__ __forceinline__ bool tensor_node_fast(const LoopInv &L, const TensorInv &TI, const PzInv &pzi, double ktx, double kty, bool &valid, cx &A) {
  const DevProc &P = c_proc;
  const double n1x = L.p1x - ktx, n1y = L.p1y - kty, n2x = L.p2x + ktx, n2y = L.p2y + kty;
  const double Xx = -(n1x + n2x), Xy = -(n1y + n2y);
  const double a1s = n1x * n1x + n1y * n1y, a2s = n2x * n2x + n2y * n2y, aXs = L.mXsq + (Xx * Xx + Xy * Xy);
  unsigned     key = min(min(fast_key(a1s), fast_key(a2s)), fast_key(aXs));   // range guards as in regge_node_fast
  const double pt1 = sqrt_fast_core(a1s), pt2 = sqrt_fast_core(a2s), mtX = sqrt_fast_core(aXs);
  const double Xz = mtX * L.shY, Xe = mtX * L.chY;
  valid = !(Xe > L.emax);
  bool         dom = true;
  const double p1z = SAME ? solve_pz3_A_fast(pzi, pt1, pt2, Xz, Xe, key, dom) : solve_pz3_B_fast(pzi, pt1, pt2, Xz, Xe, key, dom);
  const double p2z = -(Xz + p1z);
  valid = valid & !(isnan(p1z) | (p1z < 0) | (p2z > 0));
  const double e1s = L.m1sq + pow2(pt1) + pow2(p1z), e2s = L.m2sq + pow2(pt2) + pow2(p2z);
  key = min(key, min(fast_key(e1s), fast_key(e2s)));
  const double e1 = sqrt_fast_core(e1s), e2 = sqrt_fast_core(e2s);
  valid = valid & !(P.bsum_e - ((e1 + e2) + Xe) > 1e-6);
  const v4 &p3 = L.p3, &p4 = L.p4;
  const double ss13 = m2_of(n1x + p3.x, n1y + p3.y, p1z + p3.z, e1 + p3.e);
  const double ss14 = m2_of(n1x + p4.x, n1y + p4.y, p1z + p4.z, e1 + p4.e);
  const double ss23 = m2_of(n2x + p3.x, n2y + p3.y, p2z + p3.z, e2 + p3.e);
  const double ss24 = m2_of(n2x + p4.x, n2y + p4.y, p2z + p4.z, e2 + p4.e);
  valid = valid & ((__double2hiint(ss13) | __double2hiint(ss14) | __double2hiint(ss23) | __double2hiint(ss24)) >= 0);
  // t1, t2: the transverse part of q = p_beam - p is the one of p (a1s, a2s), see regge_node_fast
  const double q1z = P.pb1z - p1z, q1e = P.pb1e - e1;
  const double t1 = pow2(q1e) - (a1s + pow2(q1z));
  const double t2 = pow2(P.pb2e - e2) - (a2s + pow2(P.pb2z - p2z));
  valid = valid & !(t1 > 0);
  valid = valid & !(t2 > 0);
  // Lorentz products (q1 = (-n1x, -n1y, q1z, q1e); P1 = (n1x, n1y, P1z, P1e); P2 = (n2x, n2y, P2z, P2e)).
  // u channel: the reference forms p_u = (p4 - p_a) + p_1, whose first difference is rounded at the beam scale (half an ulp
  // of 6500 GeV = 4.5e-13 GeV in z and E).  Multiplied by beam-sized partners (P.K ~ 1e4 GeV^2, q1e up to 100 GeV) that
  // deterministic offset reaches 1e-11 ... 5e-10 of the u-channel term: it is part of the reference's result and is kept --
  // the u channel takes its own pomeron momentum q1u = p4 - p_u built in the reference's order (t channel: p_a - p_1 is exact).
  const double P1z = p1z + P.pb1z, P1e = e1 + P.pb1e, P2z = p2z + P.pb2z, P2e = e2 + P.pb2e;
  const double quz = p4.z - ((p4.z - P.pb1z) + p1z), que = p4.e - ((p4.e - P.pb1e) + e1);
  const double tu  = fma(que, que, -fma(quz, quz, a1s));
  const double q1p3 = fdot(q1e, -n1x, -n1y, q1z, p3), q1p4 = fdot(q1e, -n1x, -n1y, q1z, p4);
  const double qup3 = fdot(que, -n1x, -n1y, quz, p3), qup4 = fdot(que, -n1x, -n1y, quz, p4);
  const double P1p3 = fdot(P1e, n1x, n1y, P1z, p3), P1p4 = fdot(P1e, n1x, n1y, P1z, p4);
  const double P2p3 = fdot(P2e, n2x, n2y, P2z, p3), P2p4 = fdot(P2e, n2x, n2y, P2z, p4);
  const double P1q1 = fma(P1e, q1e, -fma(P1z, q1z, -a1s)), P1qu = fma(P1e, que, -fma(P1z, quz, -a1s));
  const double nn   = fma(n2x, n1x, n2y * n1y);
  const double P2q1 = fma(P2e, q1e, fma(-P2z, q1z, nn)), P2qu = fma(P2e, que, fma(-P2z, quz, nn));
  const double hP1 = 0.5 * fma(P1e, P1e, -fma(P1z, P1z, a1s)), hP2 = 0.5 * fma(P2e, P2e, -fma(P2z, P2z, a2s));  // P^2 / 2
  const double m34 = TI.m3sq + TI.m4sq, t34 = t1 + m34, u34 = tu + m34;
  const double pt2v = fma(-2.0, q1p3, t1 + TI.m3sq), pu2v = fma(-2.0, qup4, tu + TI.m4sq);                        // t^, u^
  const double Kta2 = fma(-2.0, q1p3, fma(3.0, TI.m3sq, pt2v)), Kua2 = fma(-2.0, qup4, fma(3.0, TI.m4sq, pu2v));
  const double Ktb2 = fma(2.0, q1p4 - q1p3, fma(-2.0, TI.p3p4, t34)), Kub2 = fma(-2.0, qup4 - qup3, fma(-2.0, TI.p3p4, u34));
  // (p3 + p4 - q)^2: pomeron 2 at the meson vertices
  const double qbt = fma(-2.0, q1p3 + q1p4, fma(2.0, TI.p3p4, t34)), qbu = fma(-2.0, qup3 + qup4, fma(2.0, TI.p3p4, u34));
  const double d1t = fma(-2.0, P1p3, P1q1), d1u = fma(2.0, P1p4, -P1qu);
  const double dP2 = P2p4 - P2p3, d2t = dP2 + P2q1, d2u = dP2 - P2qu;
  const double Rt = fma(2.0 * d1t, d1t, -(Kta2 * hP1)) * fma(2.0 * d2t, d2t, -(Ktb2 * hP2));
  const double Ru = fma(2.0 * d1u, d1u, -(Kua2 * hP1)) * fma(2.0 * d2u, d2u, -(Kub2 * hP2));
  const double al1 = fma(P.tp_ap, t1, P.tp_delta), al2 = fma(P.tp_ap, t2, P.tp_delta);
  key = min(min(key, fast_key(ss13)), min(min(fast_key(ss24), fast_key(ss14)), fast_key(ss23)));
  const double L13 = fast_log_core(ss13), L24 = fast_log_core(ss24), L14 = fast_log_core(ss14), L23 = fast_log_core(ss23);
  const double xt = fma(al1, L13 + P.tp_ln_ap, al2 * (L24 + P.tp_ln_ap)), xu = fma(al1, L14 + P.tp_ln_ap, al2 * (L23 + P.tp_ln_ap));
  const double Et = fast_exp_core(xt), Eu = fast_exp_core(xu);
  double nt, dt, nu, du;
  meson_G_nd(pt2v, nt, dt);
  meson_G_nd(pu2v, nu, du);
  // term_x = E_x R_x n_x / (16 s s g_x d_x), g_x = (1 - 2 q^2)(1 - 2 qb_x): F_M with Lambda^2 = 0.5 at both meson vertices
  const double Nt = Et * Rt * nt, Dt = (ss13 * ss24) * (fma(-2.0, t1, 1.0) * fma(-2.0, qbt, 1.0)) * dt;
  const double Nu = Eu * Ru * nu, Du = (ss14 * ss23) * (fma(-2.0, tu, 1.0) * fma(-2.0, qbu, 1.0)) * du;
  // F1(t1) F1(t2) as one rational function (Kelly) or dipole (MForm.cc:542-645)
  const double inv4mp2 = 1.0 / P.four_mp2;
  const double tau1 = fabs(t1) * inv4mp2, tau2 = fabs(t2) * inv4mp2;
  double       nF, dF;
  if (P.em_kelly) {
    const double dE1 = fma(fma(fma(21.97, tau1, 12.82), tau1, 10.98), tau1, 1.0), dM1 = fma(fma(fma(6.55, tau1, 18.86), tau1, 10.97), tau1, 1.0);
    const double dE2 = fma(fma(fma(21.97, tau2, 12.82), tau2, 10.98), tau2, 1.0), dM2 = fma(fma(fma(6.55, tau2, 18.86), tau2, 10.97), tau2, 1.0);
    const double nE1 = fma(-0.24, tau1, 1.0), nM1 = fma(2.792847337 * 0.12, tau1, 2.792847337);
    const double nE2 = fma(-0.24, tau2, 1.0), nM2 = fma(2.792847337 * 0.12, tau2, 2.792847337);
    nF = fma(nE1, dM1, (tau1 * nM1) * dE1) * fma(nE2, dM2, (tau2 * nM2) * dE2);
    dF = (dE1 * dM1 * (tau1 + 1.0)) * (dE2 * dM2 * (tau2 + 1.0));
  } else {
    const double g1 = pow2(1.0 + fabs(t1) / 0.71), g2 = pow2(1.0 + fabs(t2) / 0.71);   // GM = mu / g, GE = 1 / g
    nF = fma(tau1, 2.792847337, 1.0) * fma(tau2, 2.792847337, 1.0);
    dF = (g1 * (tau1 + 1.0)) * (g2 * (tau2 + 1.0));
  }
  const double num = (0.0625 * P.tp_coupling) * nF * fma(Nt, Du, Nu * Dt), den = (Dt * Du) * dF;
  const double R   = div_fast_core(num, den);
  const double ph  = P.tp_phase * (al1 + al2);
  const bool redo = !(fast_key_ok(key) & dom & (fabs(xt) < 690.0) & (fabs(xu) < 690.0) & div_fast_ok(num, R) & (fabs(ph) < 1.0e5));
  double sn, cs;
  fast_sincos_core(ph, sn, cs);
  A.re = cs * R;
  A.im = sn * R;
  return redo;
}
template <bool SAME>
__device__ __forceinline__ double screened_amp2_tensor_fast(const LoopInv &L, const cx &h0) {
  const DevLoop &Q = c_loop;
  PzConst        pc;
  pc.s = c_proc.s; pc.t7 = c_proc.sqrt_s; pc.t9 = c_proc.s_pow15; pc.t14 = c_proc.s * c_proc.s;
  const PzInv pzi = make_pz_inv(L.m1, L.m2, pc, SAME);
  TensorInv   TI;
  TI.m3sq = m2_of(L.p3); TI.m4sq = m2_of(L.p4); TI.p3p4 = mdot(L.p3, L.p4);
  double Ire = 0.0, Iim = 0.0;
  for (int i = 0; i < Q.nphi_nodes; ++i) {
    const double cp = Q.cphi[i], sp = Q.sphi[i], wi = Q.wphi[i];
    double       Jre = 0.0, Jim = 0.0;
    for (int j = 0; j <= Q.nkt; ++j) {
      const double kt = Q.kt[j];
      bool         valid;
      cx           A;
      if (__builtin_expect(tensor_node_fast<SAME>(L, TI, pzi, kt * cp, kt * sp, valid, A), 0)) {
        NodeKin in;
        valid = loop_node<SAME>(L, pc, kt * cp, kt * sp, in);
        if (valid) A = tensor_me4_2xS(in, L.p3, L.p4);
      }
      if (valid) {
        Jre = fma(Q.b_re[j], A.re, fma(-Q.b_im[j], A.im, Jre));
        Jim = fma(Q.b_re[j], A.im, fma(Q.b_im[j], A.re, Jim));
      }
    }
    Ire = fma(wi, Jre, Ire);
    Iim = fma(wi, Jim, Iim);
  }
  const double ar = h0.re - Q.norm * Iim, ai = h0.im + Q.norm * Ire;
  return ar * ar + ai * ai;
}

// The same loop for NH > 1 amplitudes.  The Born amplitudes are re-evaluated from the Born-level kinematics carried
// in the work list (same device function, same inputs as k_points: identical bits), so the list does not grow with NH.
template <int AMP, bool SAME>
__device__ __forceinline__ double screened_amp2_multi(const LoopInv &L, const NodeKin &born) {
  constexpr int  NH = AmpHel<AMP>::NH;
  const DevLoop &Q  = c_loop;
  PzConst        pc;
  pc.s = c_proc.s; pc.t7 = c_proc.sqrt_s; pc.t9 = c_proc.s_pow15; pc.t14 = c_proc.s * c_proc.s;
  double Ire[NH], Iim[NH];
#pragma unroll
  for (int h = 0; h < NH; ++h) { Ire[h] = 0.0; Iim[h] = 0.0; }
#pragma unroll 1
  for (int i = 0; i < Q.nphi_nodes; ++i) {
    const double cp = Q.cphi[i], sp = Q.sphi[i], wi = Q.wphi[i];
#pragma unroll 1
    for (int j = 0; j <= Q.nkt; ++j) {
      const double kt = Q.kt[j];
      NodeKin      in;
      if (!loop_node_any<SAME>(L, pc, kt * cp, kt * sp, in)) continue;
      cx A[NH];
      amplitude_vec<AMP>(in, L.p3, L.p4, A);
      const double br = wi * Q.b_re[j], bi = wi * Q.b_im[j];
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        Ire[h] = fma(br, A[h].re, fma(-bi, A[h].im, Ire[h]));
        Iim[h] = fma(br, A[h].im, fma(bi, A[h].re, Iim[h]));
      }
    }
  }
  cx h0[NH];
  amplitude_vec<AMP>(born, L.p3, L.p4, h0);
  double amp2 = 0.0;
#pragma unroll
  for (int h = 0; h < NH; ++h) {
    const double ar = h0[h].re - Q.norm * Iim[h], ai = h0[h].im + Q.norm * Ire[h];  // A_0 + i norm I
    amp2 += ar * ar + ai * ai;
  }
  return amp_hel_scale<AMP>() * amp2;
}

// ---- four- / six-leg ladder: loop node and loop integral ----------------------------------------------------------------
template <bool SAME, int K>
__device__ __forceinline__ bool loop_node_CK(const LoopInvK<K> &L, const PzConst &pc, double ktx, double kty, double &t1, double &t2, Ladder<K> &X) {
  const DevProc &P = c_proc;
  v4 p1, p2, pX, leg[K];
  p1.x = L.p1x - ktx; p1.y = L.p1y - kty; p2.x = L.p2x + ktx; p2.y = L.p2y + kty;
  const double wx = p1.x + p2.x, wy = p1.y + p2.y;
  double bx[K], by[K];
  bx[0] = L.qx[0] - wx; by[0] = L.qy[0] - wy;
#pragma unroll
  for (int i = 0; i < K - 1; ++i) { bx[i + 1] = L.qx[i] * (-1.0) - wx; by[i + 1] = L.qy[i] * (-1.0) - wy; }
  pX.x = 0.0; pX.y = 0.0; pX.z = 0.0; pX.e = 0.0;
#pragma unroll
  for (int i = 0; i < K; ++i) {
    v4 &p = leg[i];
    p.x = blin_row<K>(i, bx);
    p.y = blin_row<K>(i, by);
    const double mt = msqrt(pow2(L.m[i]) + (p.x * p.x + p.y * p.y));
    p.z = mt * L.sh[i]; p.e = mt * L.ch[i];
    pX.x += p.x; pX.y += p.y; pX.z += p.z; pX.e += p.e;
  }
  const double pt1 = msqrt(p1.x * p1.x + p1.y * p1.y), pt2 = msqrt(p2.x * p2.x + p2.y * p2.y);
  const double p1z = SAME ? solve_pz3_A(L.m1, pt1, pt2, pX.z, pX.e, pc) : solve_pz3_B(L.m1, L.m2, pt1, pt2, pX.z, pX.e, pc);
  const double p2z = -(pX.z + p1z);
  if (isnan(p1z) || p1z < 0 || p2z > 0) return false;
  p1.z = p1z; p1.e = msqrt(L.m1sq + pow2(pt1) + pow2(p1z));
  p2.z = p2z; p2.e = msqrt(L.m2sq + pow2(pt2) + pow2(p2z));
  {
    const double dx = 0.0 - ((p1.x + p2.x) + pX.x), dy = 0.0 - ((p1.y + p2.y) + pX.y);
    const double dz = P.bsum_z - ((p1.z + p2.z) + pX.z), de = P.bsum_e - ((p1.e + p2.e) + pX.e);
    if (fabs(dx) > 1e-6 || fabs(dy) > 1e-6 || fabs(dz) > 1e-6 || de > 1e-6) return false;
  }
  double s1, s2;
  return ladder_scalars<K>(p1, p2, pX, leg, t1, t2, s1, s2, X);
}
template <bool SAME, int K>
__device__ __forceinline__ double screened_amp2_ladder(const LoopInvK<K> &L, const cx &h0) {
  const DevLoop &Q = c_loop;
  PzConst        pc;
  pc.s = c_proc.s; pc.t7 = c_proc.sqrt_s; pc.t9 = c_proc.s_pow15; pc.t14 = c_proc.s * c_proc.s;
  double Ire = 0.0, Iim = 0.0;
#pragma unroll 1
  for (int i = 0; i < Q.nphi_nodes; ++i) {
    const double cp = Q.cphi[i], sp = Q.sphi[i], wi = Q.wphi[i];
    double       Jre = 0.0, Jim = 0.0;
#pragma unroll 1
    for (int j = 0; j <= Q.nkt; ++j) {
      const double kt = Q.kt[j];
      double       t1, t2;
      Ladder<K>    X;
      if (!loop_node_CK<SAME, K>(L, pc, kt * cp, kt * sp, t1, t2, X)) continue;
      const cx A = regge_ladder<K>(t1, t2, X);
      Jre = fma(Q.b_re[j], A.re, fma(-Q.b_im[j], A.im, Jre));
      Jim = fma(Q.b_re[j], A.im, fma(Q.b_im[j], A.re, Jim));
    }
    Ire = fma(wi, Jre, Ire);
    Iim = fma(wi, Jim, Iim);
  }
  const double ar = h0.re - Q.norm * Iim, ai = h0.im + Q.norm * Ire;
  return ar * ar + ai * ai;
}

}  // namespace grb
