// grb_tres.cuh -- tensor-pomeron resonance production  p p -> p + R(-> two pseudoscalars) + p,  R = 0+ or 2+ :
// the contracted closed form of the reference's MTensorPomeron::ME3 (src/MTensorPomeron.cc:164-564).
//
// The reference multiplies explicit Lorentz tensors through FTensor loops: proton vertices iG_PppHE (rank 2, :1388-1399),
// pomeron propagators iD_P (rank 4, :2288-2298), the pomeron-pomeron-scalar vertex iG_PPS_total (rank 4, :1492-1567) or the
// pomeron-pomeron-tensor vertex iG_PPT_total (rank 6, seven coupling structures, :1646-2015) with the spin-2 propagator
// iD_TMES (:2422-2446) and the decay vertex iG_f2psps (:2121-2137).  Here every contraction is done on paper:
//
//   proton vertex x propagator   X_k^{ab} = c_k Xh_k^{ab},  Xh_k = 2 P_k^a P_k^b - 1/2 g^{ab} P_k^2,  P_1 = p_1 + p_a, P_2 = p_2 + p_b
//                                c_k = -3i g_PNN F_1(t_k) (1/(4 s_k)) (-i s_k alpha')^{alpha(t_k) - 1}
//   spin-2 propagator x decay    Y^{rs} = Delta (g_d/2) F(m_34^2) Yh^{rs},  Yh = kh kh + 1/3 kh^2 gh,
//                                gh = -g + p p / p^2,  kh = gh k,  k = p_3 - p_4,  p = p_X,  Delta = 1/(p^2 - M0^2 + i M0 Gamma)
//
// Xh_1, Xh_2, Yh are symmetric and traceless, so the reference's auxiliary tensor R = 1/2 (g g + g g) + 1/4 g g (CalcRTensor,
// :2538-2575; the +1/4 is the reference's, kept by the CPU checker) acts on them as the identity and each coupling structure
// collapses to a Lorentz invariant of (Xh_1, Xh_2, Yh, q_1, q_2) -- products of the three linear maps
//   Xh_k v = 2 P_k (P_k.v) - 1/2 P_k^2 v,      Yh v = kh (kh.v) + 1/3 kh^2 (-v + p (p.v)/p^2)
// applied to 4-vectors.  With T = Tr(Xh_1 Xh_2 Yh), XX = Xh_1 : Xh_2 = 4 (P_1.P_2)^2 - P_1^2 P_2^2 :
//   #0 (l,s)=(0,2)        2 g_0 T
//   #1, #2 (2,0)-/+(2,2)  -2 g_m [ q_1.q_2 T -/+ ( q_2 Xh_1 Xh_2 Yh q_1 + q_1 Xh_2 Xh_1 Yh q_2 ) + XX (q_1 Yh q_2) ]
//   #3 (2,4)              -2 g_3 (Xh_1 q_2) Yh (Xh_2 q_1)
//   #4 (4,2)              -4 g_4 [ q_2 Xh_1 Xh_2 q_1 - q_1.q_2 XX ] (q_1 Yh q_2)
//   #5 (4,4)              g_5 [ (q_2 Xh_1 Yh q_2)(q_1 Xh_2 q_1) + (q_1 Xh_2 Yh q_1)(q_2 Xh_1 q_2) ]
//   #6 (6,4)              -2 g_6 (q_2 Xh_1 q_2)(q_1 Xh_2 q_1)(q_1 Yh q_2)
// (structures 1..6 enter only where |g_m| > 1e-6, :1660-1672) and for the scalar vertex
//   0+ :  2 g_0 XX + 2 g_1 [ q_2 Xh_1 Xh_2 q_1 - q_1.q_2 XX ].
// All i's multiply out to a real structure times  c_1 c_2 Delta :
//   A(2+) = -K F(m_X^2) (g_d/2) F(m_34^2) S_T Delta,     A(0+) = +K F(m_X^2) g_d F(m_34^2) S_S Delta,
//   K = 9 g_PNN^2 F_1(t_1) F_1(t_2) F_M(t_1) F_M(t_2) (s_1 a')^{a_1} (s_2 a')^{a_2} / (16 s_1 s_2) e^{-i pi (a_1 + a_2)/2},
//   a_k = delta_P + a' t_k,  F(x) = exp(-(x - M0^2)^2 / Lambda^4), Lambda = 1 GeV,  F_M(t) = 1/(1 - t / 1 GeV^2).
// The four helicity-conserving proton combinations carry the same amplitude (:232-252), so one complex number per point.
//
// The header is plain C++ under a host compiler (tests/host_tres.cc instantiates it in double and long double to pin the
// algebra against the explicit-tensor CPU restatement and to measure the rounding noise of the reference's explicit contraction) and device
// code under nvcc; the product only ever calls it from the CUDA kernels.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define GRB_TRES_FN __host__ __device__ __forceinline__
#else
#define GRB_TRES_FN inline
#endif

namespace grb {
namespace tres {

template <typename R>
struct V4 {
  R x, y, z, e;
};
template <typename R>
GRB_TRES_FN R dot(const V4<R> &a, const V4<R> &b) { return a.e * b.e - (a.x * b.x + a.y * b.y + a.z * b.z); }
template <typename R>
GRB_TRES_FN V4<R> lin(R ca, const V4<R> &a, R cb, const V4<R> &b) {  // ca a + cb b
  V4<R> r;
  r.x = ca * a.x + cb * b.x; r.y = ca * a.y + cb * b.y; r.z = ca * a.z + cb * b.z; r.e = ca * a.e + cb * b.e;
  return r;
}

GRB_TRES_FN double r_exp(double x) { return exp(x); }
#if !defined(__CUDACC__)
inline long double r_exp(long double x) { return expl(x); }
#endif

// one resonance as the kernels see it (constant memory): only what ME3 reads
struct ResPar {
  int    J, on;        // spin (0 or 2); bit m of `on` set where |g[m]| > 1e-6 (structures 1..6)
  double M0sq, M0G;    // M0^2, M0 Gamma
  double g[7];         // production couplings g_Tensor
  double gd;           // decay coupling g_decay_tensor[0]
};

// the linear maps of the header comment
template <typename R>
struct Maps {
  V4<R> P1, P2, p, kh;
  R     P1sq, P2sq, kh2, inv_p2;
  GRB_TRES_FN V4<R> X1(const V4<R> &v) const { return lin<R>(R(2) * dot(P1, v), P1, R(-0.5) * P1sq, v); }
  GRB_TRES_FN V4<R> X2(const V4<R> &v) const { return lin<R>(R(2) * dot(P2, v), P2, R(-0.5) * P2sq, v); }
  GRB_TRES_FN V4<R> Y(const V4<R> &v) const {
    const R   third = kh2 / R(3);
    V4<R>     t     = lin<R>(dot(kh, v), kh, -third, v);
    const R   c     = third * (dot(p, v) * inv_p2);
    t.x += c * p.x; t.y += c * p.y; t.z += c * p.z; t.e += c * p.e;
    return t;
  }
};

// real Lorentz structure of the pomeron-pomeron-tensor vertex contracted with Xh_1, Xh_2, Yh
template <typename R>
GRB_TRES_FN R tensor_structure(const Maps<R> &m, const V4<R> &q1, const V4<R> &q2, const double *g, int on) {
  const R      q1q2 = dot(q1, q2), P12 = dot(m.P1, m.P2);
  const R      XX   = R(4) * (P12 * P12) - m.P1sq * m.P2sq;
  const V4<R> X1q2 = m.X1(q2), X2q1 = m.X2(q1), Yq1 = m.Y(q1), Yq2 = m.Y(q2);
  // T = Tr(Xh_1 Xh_2 Yh) = 2 (Xh_2 P_1).(Yh P_1) - P_1^2 (P_2 Yh P_2)
  const R T = R(2) * dot(m.X2(m.P1), m.Y(m.P1)) - m.P1sq * dot(m.P2, m.Y(m.P2));
  R       S = R(2) * R(g[0]) * T;
  const R q1Yq2 = dot(q1, Yq2);
  if (on & 6) {
    const R chain = dot(m.X2(X1q2), Yq1) + dot(m.X1(X2q1), Yq2);  // q_2 Xh_1 Xh_2 Yh q_1 + q_1 Xh_2 Xh_1 Yh q_2
    const R base  = q1q2 * T + XX * q1Yq2;
    if (on & 2) S += R(-2) * R(g[1]) * (base - chain);
    if (on & 4) S += R(-2) * R(g[2]) * (base + chain);
  }
  const R q2X1X2q1 = dot(X1q2, X2q1);
  const R q2X1q2 = dot(q2, X1q2), q1X2q1 = dot(q1, X2q1);
  if (on & 8) S += R(-2) * R(g[3]) * dot(X1q2, m.Y(X2q1));
  if (on & 16) S += R(-4) * R(g[4]) * (q2X1X2q1 - q1q2 * XX) * q1Yq2;
  if (on & 32) S += R(g[5]) * (dot(X1q2, Yq2) * q1X2q1 + dot(X2q1, Yq1) * q2X1q2);
  if (on & 64) S += R(-2) * R(g[6]) * (q2X1q2 * q1X2q1 * q1Yq2);
  return S;
}

// sum over resonances of  (real structure) x F(m_X^2) F(m_34^2) x Delta : returns (re, im) WITHOUT the common factor K.
//   pa, pb beams; p1, p2 outgoing protons; pX central system (lts.pfinal[0]); p3, p4 its daughters (stale inside the
//   screening loop, like the reference's lts.decaytree: src/MFactorized.cc:84-139)
template <typename R>
GRB_TRES_FN void resonance_sum(const ResPar *res, int n_res, const V4<R> &pa, const V4<R> &pb, const V4<R> &p1, const V4<R> &p2,
                               const V4<R> &pX, const V4<R> &p3, const V4<R> &p4, R &out_re, R &out_im) {
  Maps<R> m;
  m.P1 = lin<R>(R(1), p1, R(1), pa);
  m.P2 = lin<R>(R(1), p2, R(1), pb);
  m.P1sq = dot(m.P1, m.P1);
  m.P2sq = dot(m.P2, m.P2);
  const V4<R> q1 = lin<R>(R(1), pa, R(-1), p1), q2 = lin<R>(R(1), pb, R(-1), p2);
  const V4<R> qs = lin<R>(R(1), q1, R(1), q2), ps = lin<R>(R(1), p3, R(1), p4), k = lin<R>(R(1), p3, R(-1), p4);
  const R      mq2 = dot(qs, qs);   // (q1 + q2)^2: argument of the production form factor F~
  const R      m34 = dot(ps, ps);   // (p3 + p4)^2: argument of the decay form factor
  const R      mX2 = dot(pX, pX);   // propagator mass
  m.p      = pX;
  m.inv_p2 = R(1) / mX2;
  m.kh     = lin<R>(R(-1), k, dot(pX, k) * m.inv_p2, pX);
  m.kh2    = dot(m.kh, m.kh);
  // scalar structure (computed once: independent of the resonance up to g_0, g_1)
  const R      q1q2 = dot(q1, q2), P12 = dot(m.P1, m.P2);
  const R      XX   = R(4) * (P12 * P12) - m.P1sq * m.P2sq;
  const R      sc1  = dot(m.X1(q2), m.X2(q1)) - q1q2 * XX;
  R re = R(0), im = R(0);
  for (int i = 0; i < n_res; ++i) {
    const ResPar &r  = res[i];
    const R       M2 = R(r.M0sq);
    const R       dq = mq2 - M2, d34 = m34 - M2;
    const R       ff = r_exp(-(dq * dq) - d34 * d34);
    R             S;
    if (r.J == 0) S = R(r.gd) * (R(2) * R(r.g[0]) * XX + R(2) * R(r.g[1]) * sc1);
    else S = R(-0.5) * R(r.gd) * tensor_structure<R>(m, q1, q2, r.g, r.on);
    // Delta = 1 / (dm + i M0 Gamma) = (dm - i M0 Gamma) / (dm^2 + (M0 Gamma)^2)
    const R dm = mX2 - M2, w = R(r.M0G), inv = (S * ff) / (dm * dm + w * w);
    re += inv * dm;
    im -= inv * w;
  }
  out_re = re;
  out_im = im;
}

}  // namespace tres
}  // namespace grb
