// grb_ferm.cuh -- spin-1/2 central pairs of MTensorPomeron::ME4: tensor pomeron 2xF (p pbar, Lambda Lambdabar;
// reference src/MTensorPomeron.cc:765-804) and the yy[QED] lepton branch (:664-713), Dirac algebra of
// src/MDirac.cc in the DIRAC basis.
//
// The reference builds 4x4 gamma matrices, rank-4 vertex tensors iG_PppbarP (256 spinor sandwiches per helicity
// pair) and contracts them with FTensor.  With the high-energy proton vertex i Gamma_pp = -3i g F1 P P and
// i Delta = E(s,t) [g g + g g - 1/2 g g], a pomeron leg contracted into the traceless fermion vertex
//   Gamma_{mu nu}(k',k) = 1/2 (gamma_mu K_nu + gamma_nu K_mu) - 1/4 g_{mu nu} Kslash,  K = k' + k,
// collapses to  Vslash  with the real 4-vector  V = 2 (P.K) P - 1/2 P^2 K.  Every helicity amplitude is then a chain
//   ubar(p4) Vslash_L (qslash + M) Vslash_R v(p3)
// of three slash applications on a 4-spinor: no matrices, no tensors.  In the Dirac basis
//   aslash = [[a0, 0, -az, -(ax - i ay)], [0, a0, -(ax + i ay), az], [az, ax - i ay, -a0, 0], [ax + i ay, -az, 0, -a0]].
// The QED branch contracts complex currents the same way (slash of a complex 4-vector).
#pragma once

namespace grb {

struct sp4 { cx c[4]; };
struct cv4 { cx e, x, y, z; };  // complex contravariant 4-vector

__device__ __forceinline__ cx fcmk(double re, double im) { cx r; r.re = re; r.im = im; return r; }
__device__ __forceinline__ cx fcadd(cx a, cx b) { return fcmk(a.re + b.re, a.im + b.im); }
__device__ __forceinline__ cx fcsub(cx a, cx b) { return fcmk(a.re - b.re, a.im - b.im); }
__device__ __forceinline__ cx fcmul(cx a, cx b) { return fcmk(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
__device__ __forceinline__ cx fcmul(double a, cx b) { return fcmk(a * b.re, a * b.im); }
__device__ __forceinline__ cx fcmuli(cx a) { return fcmk(-a.im, a.re); }  // i * a
__device__ __forceinline__ cx fcneg(cx a) { return fcmk(-a.re, -a.im); }

// aslash psi for a vector with components of type T (double or cx); am = ax - i ay, ap = ax + i ay
__device__ __forceinline__ sp4 slash_apply(double a0, double ax, double ay, double az, const sp4 &p) {
  const cx am = fcmk(ax, -ay), ap = fcmk(ax, ay);
  sp4      o;
  o.c[0] = fcsub(fcsub(fcmul(a0, p.c[0]), fcmul(az, p.c[2])), fcmul(am, p.c[3]));
  o.c[1] = fcadd(fcsub(fcmul(a0, p.c[1]), fcmul(ap, p.c[2])), fcmul(az, p.c[3]));
  o.c[2] = fcsub(fcadd(fcmul(az, p.c[0]), fcmul(am, p.c[1])), fcmul(a0, p.c[2]));
  o.c[3] = fcsub(fcsub(fcmul(ap, p.c[0]), fcmul(az, p.c[1])), fcmul(a0, p.c[3]));
  return o;
}
__device__ __forceinline__ sp4 slash_apply(const v4 &a, const sp4 &p) { return slash_apply(a.e, a.x, a.y, a.z, p); }
__device__ __forceinline__ sp4 slash_apply(const cv4 &a, const sp4 &p) {
  const cx am = fcmk(a.x.re + a.y.im, a.x.im - a.y.re), ap = fcmk(a.x.re - a.y.im, a.x.im + a.y.re);
  sp4      o;
  o.c[0] = fcsub(fcsub(fcmul(a.e, p.c[0]), fcmul(a.z, p.c[2])), fcmul(am, p.c[3]));
  o.c[1] = fcadd(fcsub(fcmul(a.e, p.c[1]), fcmul(ap, p.c[2])), fcmul(a.z, p.c[3]));
  o.c[2] = fcsub(fcadd(fcmul(a.z, p.c[0]), fcmul(am, p.c[1])), fcmul(a.e, p.c[2]));
  o.c[3] = fcsub(fcsub(fcmul(ap, p.c[0]), fcmul(a.z, p.c[1])), fcmul(a.e, p.c[3]));
  return o;
}
// row spinor times aslash (real vector)
__device__ __forceinline__ sp4 slash_apply_left(const sp4 &b, const v4 &a) {
  const cx am = fcmk(a.x, -a.y), ap = fcmk(a.x, a.y);
  sp4      o;
  o.c[0] = fcadd(fcadd(fcmul(a.e, b.c[0]), fcmul(a.z, b.c[2])), fcmul(ap, b.c[3]));
  o.c[1] = fcsub(fcadd(fcmul(a.e, b.c[1]), fcmul(am, b.c[2])), fcmul(a.z, b.c[3]));
  o.c[2] = fcneg(fcadd(fcadd(fcmul(a.z, b.c[0]), fcmul(ap, b.c[1])), fcmul(a.e, b.c[2])));
  o.c[3] = fcsub(fcsub(fcmul(a.z, b.c[1]), fcmul(am, b.c[0])), fcmul(a.e, b.c[3]));
  return o;
}
__device__ __forceinline__ cx sp_dot(const sp4 &b, const sp4 &p) {
  return fcadd(fcadd(fcmul(b.c[0], p.c[0]), fcmul(b.c[1], p.c[1])), fcadd(fcmul(b.c[2], p.c[2]), fcmul(b.c[3], p.c[3])));
}
// J[mu] = bar gamma_lo[mu] u  (covariant index: gamma_0 = gamma^0, gamma_i = -gamma^i), MDirac.cc:47-120
__device__ __forceinline__ void current_lo(const sp4 &b, const sp4 &u, cx J[4]) {
  J[0] = fcsub(fcadd(fcmul(b.c[0], u.c[0]), fcmul(b.c[1], u.c[1])), fcadd(fcmul(b.c[2], u.c[2]), fcmul(b.c[3], u.c[3])));
  J[1] = fcneg(fcsub(fcadd(fcmul(b.c[0], u.c[3]), fcmul(b.c[1], u.c[2])), fcadd(fcmul(b.c[2], u.c[1]), fcmul(b.c[3], u.c[0]))));
  J[2] = fcmuli(fcsub(fcadd(fcmul(b.c[0], u.c[3]), fcmul(b.c[3], u.c[0])), fcadd(fcmul(b.c[1], u.c[2]), fcmul(b.c[2], u.c[1]))));
  J[3] = fcneg(fcadd(fcsub(fcmul(b.c[0], u.c[2]), fcmul(b.c[1], u.c[3])), fcsub(fcmul(b.c[3], u.c[1]), fcmul(b.c[2], u.c[0]))));
}

// Helicity spinors of MDirac::uHelDirac (:224-246): cos/sin of theta/2 and e^{i phi} from the momentum components
// (the reference goes through atan2 and cos/sin: same values to a few ulp).
struct HelBasis {
  double N, r;     // sqrt(E+m), sqrt(E+m) |p| / (E+m)
  double c, s;     // cos(theta/2), sin(theta/2)
  cx     ph;       // e^{i phi}
};
__device__ __forceinline__ HelBasis hel_basis(const v4 &p) {
  HelBasis     B;
  const double m = m_of(p), pt2 = pt2_of(p), pm = sqrt(pt2 + p.z * p.z), pt = sqrt(pt2);
  B.N = msqrt(p.e + m);
  B.r = B.N * pm / (p.e + m);
  if (pm == 0.0) { B.c = 1.0; B.s = 0.0; }
  else if (p.z >= 0.0) { const double d = pm + p.z; B.c = sqrt(d / (2.0 * pm)); B.s = pt / sqrt(2.0 * pm * d); }
  else { const double d = pm - p.z; B.s = sqrt(d / (2.0 * pm)); B.c = pt / sqrt(2.0 * pm * d); }
  B.ph = (pt == 0.0) ? fcmk(1.0, 0.0) : fcmk(p.x / pt, p.y / pt);
  return B;
}
__device__ __forceinline__ sp4 u_hel(const HelBasis &B, int hel) {
  sp4 u;
  if (hel > 0) {
    u.c[0] = fcmk(B.N * B.c, 0.0); u.c[1] = fcmul(B.N * B.s, B.ph);
    u.c[2] = fcmk(B.r * B.c, 0.0); u.c[3] = fcmul(B.r * B.s, B.ph);
  } else {
    u.c[0] = fcmk(-B.N * B.s, 0.0); u.c[1] = fcmul(B.N * B.c, B.ph);
    u.c[2] = fcmk(B.r * B.s, 0.0);  u.c[3] = fcmul(-B.r * B.c, B.ph);
  }
  return u;
}
__device__ __forceinline__ sp4 v_hel(const HelBasis &B, int hel) {  // :248-257
  const sp4 u = u_hel(B, -hel);
  sp4       v;
  v.c[0] = u.c[2]; v.c[1] = u.c[3]; v.c[2] = u.c[0]; v.c[3] = u.c[1];
  return v;
}
__device__ __forceinline__ sp4 sp_bar(const sp4 &s) {  // conj(s) . gamma^0
  sp4 b;
  b.c[0] = fcmk(s.c[0].re, -s.c[0].im); b.c[1] = fcmk(s.c[1].re, -s.c[1].im);
  b.c[2] = fcmk(-s.c[2].re, s.c[2].im); b.c[3] = fcmk(-s.c[3].re, s.c[3].im);
  return b;
}

__device__ __forceinline__ void proton_GEGM(double t, double &GE, double &GM, double &tau) {  // MForm.cc:566-645
  const DevProc &P  = c_proc;
  const double   Q2 = fabs(t);
  tau = Q2 / P.four_mp2;
  if (P.em_kelly) {
    GE = (1.0 - 0.24 * tau) / (1.0 + 10.98 * tau + 12.82 * (tau * tau) + 21.97 * (tau * (tau * tau)));
    GM = 2.792847337 * (1.0 + 0.12 * tau) / (1.0 + 10.97 * tau + 18.86 * (tau * tau) + 6.55 * (tau * (tau * tau)));
  } else {
    GM = 2.792847337 / pow2(1.0 + Q2 / 0.71);
    GE = GM / 2.792847337;
  }
}

__device__ __forceinline__ v4 v4_lin(double a, const v4 &A, double b, const v4 &B) {  // a A - b B
  v4 r;
  r.x = a * A.x - b * B.x; r.y = a * A.y - b * B.y; r.z = a * A.z - b * B.z; r.e = a * A.e - b * B.e;
  return r;
}
__device__ __forceinline__ v4 v4_add(const v4 &A, const v4 &B) { v4 r; r.x = A.x + B.x; r.y = A.y + B.y; r.z = A.z + B.z; r.e = A.e + B.e; return r; }
__device__ __forceinline__ v4 v4_sub(const v4 &A, const v4 &B) { v4 r; r.x = A.x - B.x; r.y = A.y - B.y; r.z = A.z - B.z; r.e = A.e - B.e; return r; }

// ---- tensor pomeron 2 x fermion: four distinct helicity amplitudes (h3, h4); the four helicity-conserving proton
// combinations carry the same ones (lts.hamp has 16 entries, |A|^2 = sum_16 / 4 = sum_4) ------------------------
__device__ __forceinline__ void tensor_me4_2xF(const NodeKin &in, const v4 &p3, const v4 &p4, cx h[4]) {
  const DevProc &P = c_proc;
  v4 pa, pb;
  pa.x = 0.0; pa.y = 0.0; pa.z = P.pb1z; pa.e = P.pb1e;
  pb.x = 0.0; pb.y = 0.0; pb.z = P.pb2z; pb.e = P.pb2e;
  const v4 P1 = v4_add(in.p1, pa), P2 = v4_add(in.p2, pb);
  const v4 pt = v4_sub(v4_sub(pa, in.p1), p3), pu = v4_add(v4_sub(p4, pa), in.p1);
  const v4 KtL = v4_add(p4, pt), KtR = v4_sub(pt, p3), KuL = v4_add(p4, pu), KuR = v4_sub(pu, p3);
  const double P1sq = m2_of(P1), P2sq = m2_of(P2);
  // t channel: left vertex on pomeron 2, right vertex on pomeron 1; u channel the other way round (:783-796)
  const v4 VtL = v4_lin(2.0 * mdot(P2, KtL), P2, 0.5 * P2sq, KtL), VtR = v4_lin(2.0 * mdot(P1, KtR), P1, 0.5 * P1sq, KtR);
  const v4 VuL = v4_lin(2.0 * mdot(P1, KuL), P1, 0.5 * P1sq, KuL), VuR = v4_lin(2.0 * mdot(P2, KuR), P2, 0.5 * P2sq, KuR);
  const double a1 = P.tp_delta + P.tp_ap * in.t1, a2 = P.tp_delta + P.tp_ap * in.t2;
  const double Et = exp(a1 * (log(in.ss13) + P.tp_ln_ap) + a2 * (log(in.ss24) + P.tp_ln_ap)) / (16.0 * in.ss13 * in.ss24);
  const double Eu = exp(a1 * (log(in.ss14) + P.tp_ln_ap) + a2 * (log(in.ss23) + P.tp_ln_ap)) / (16.0 * in.ss14 * in.ss23);
  // F1 at each fermion vertex takes (k'-k)^2 as the reference forms it (:1445-1446)
  const double F1t = proton_F1(m2_of(v4_sub(p4, pt))) * proton_F1(m2_of(v4_add(pt, p3)));
  const double F1u = proton_F1(m2_of(v4_sub(p4, pu))) * proton_F1(m2_of(v4_add(pu, p3)));
  const double pre = P.tp_ff_coupling * proton_F1(in.t1) * proton_F1(in.t2);
  const double St = pre * F1t * Et * meson_G(m2_of(pt)), Su = pre * F1u * Eu * meson_G(m2_of(pu));
  double       sn, cs;
  sincos(P.tp_phase * (a1 + a2), &sn, &cs);
  const HelBasis B3 = hel_basis(p3), B4 = hel_basis(p4);
  const double   M  = P.mass_leg0;
#pragma unroll
  for (int h3 = 0; h3 < 2; ++h3) {
    const sp4 v3 = v_hel(B3, h3 == 0 ? -1 : 1);
    sp4       wt = slash_apply(VtR, v3), wu = slash_apply(VuR, v3);
    sp4       xt = slash_apply(pt, wt), xu = slash_apply(pu, wu);
#pragma unroll
    for (int k = 0; k < 4; ++k) { xt.c[k] = fcadd(xt.c[k], fcmul(M, wt.c[k])); xu.c[k] = fcadd(xu.c[k], fcmul(M, wu.c[k])); }
    wt = slash_apply(VtL, xt);
    wu = slash_apply(VuL, xu);
#pragma unroll
    for (int h4 = 0; h4 < 2; ++h4) {
      const sp4 ub = sp_bar(u_hel(B4, h4 == 0 ? -1 : 1));
      const cx  ct = sp_dot(ub, wt), cu = sp_dot(ub, wu);
      const cx  R  = fcmk(St * ct.re + Su * cu.re, St * ct.im + Su * cu.im);
      h[2 * h3 + h4] = fcmk(R.re * cs - R.im * sn, R.re * sn + R.im * cs);
    }
  }
}
__device__ __forceinline__ double tensor_2xF_amp2(const NodeKin &in, const v4 &p3, const v4 &p4) {
  cx h[4];
  tensor_me4_2xF(in, p3, p4, h);
  double a2 = 0.0;
#pragma unroll
  for (int k = 0; k < 4; ++k) a2 += h[k].re * h[k].re + h[k].im * h[k].im;
  return a2;
}

// ---- yy[QED] -> l+ l-: full proton-photon-proton spinor vertices (iG_ypp :1342-1363, F1 gamma_mu + i F2
// sigma_{mu nu} q^nu / 2 m_p), photon propagators, lepton line.  16 distinct amplitudes in the reference's order
// (ha, hb, h3, h4), |A|^2 = sum / 4. ---------------------------------------------------------------------------
__device__ __forceinline__ void qed_proton_current(const v4 &prime, const v4 &pbeam, const double *bsp, int hel, double tk,
                                                   cv4 &L) {
  const DevProc &P = c_proc;
  const v4       q = v4_sub(prime, pbeam);
  double         GE, GM, tau;
  proton_GEGM(m2_of(q), GE, GM, tau);
  const double F1 = 1.0 / (tau + 1) * GE + tau / (tau + 1) * GM, F2 = -1.0 / (tau + 1) * GE + 1.0 / (tau + 1) * GM;
  sp4          u;
#pragma unroll
  for (int k = 0; k < 4; ++k) u.c[k] = fcmk(bsp[k], 0.0);
  const sp4 ub = sp_bar(u_hel(hel_basis(prime), hel));
  const sp4 Qu = slash_apply(q, u), bQ = slash_apply_left(ub, q);
  cx        Jg[4], Ja[4], Jb[4], G[4];
  current_lo(ub, u, Jg);
  current_lo(ub, Qu, Ja);
  current_lo(bQ, u, Jb);
  const double f2 = F2 * P.inv_2mp;
#pragma unroll
  for (int mu = 0; mu < 4; ++mu) {
    // sigma term: (i/2)(Ja - Jb); times i F2/(2 mp): -(1/2) f2 (Ja - Jb)
    const cx S = fcmul(-0.5 * f2, fcsub(Ja[mu], Jb[mu]));
    const cx V = fcadd(fcmul(F1, Jg[mu]), S);
    // (-i e) V, then the photon propagator -i g_mumu / t: together -e g_mumu / t
    const double g = (mu == 0) ? 1.0 : -1.0;
    G[mu] = fcmul(-P.qed_e * g / tk, V);
  }
  L.e = G[0]; L.x = G[1]; L.y = G[2]; L.z = G[3];
}

__device__ __forceinline__ void yy_qed_amp(const NodeKin &in, const v4 &p3, const v4 &p4, cx h[16]) {
  const DevProc &P = c_proc;
  v4 pa, pb;
  pa.x = 0.0; pa.y = 0.0; pa.z = P.pb1z; pa.e = P.pb1e;
  pb.x = 0.0; pb.y = 0.0; pb.z = P.pb2z; pb.e = P.pb2e;
  const v4 pt = v4_sub(v4_sub(pa, in.p1), p3), pu = v4_add(v4_sub(p4, pa), in.p1);
  const double   M  = P.mass_leg0;
  const double   e2 = P.qed_e * P.qed_e;
  const double   St = -e2 / (m2_of(pt) - P.M2_leg0), Su = -e2 / (m2_of(pu) - P.M2_leg0);
  const HelBasis B3 = hel_basis(p3), B4 = hel_basis(p4);
  sp4 v3[2], ub4[2];
  v3[0] = v_hel(B3, -1); v3[1] = v_hel(B3, 1);
  ub4[0] = sp_bar(u_hel(B4, -1)); ub4[1] = sp_bar(u_hel(B4, 1));
#pragma unroll 1
  for (int ha = 0; ha < 2; ++ha) {
    cv4 L1;
    qed_proton_current(in.p1, pa, &P.bsp[0][ha][0], ha == 0 ? -1 : 1, in.t1, L1);
#pragma unroll 1
    for (int hb = 0; hb < 2; ++hb) {
      cv4 L2;
      qed_proton_current(in.p2, pb, &P.bsp[1][hb][0], hb == 0 ? -1 : 1, in.t2, L2);
#pragma unroll
      for (int h3 = 0; h3 < 2; ++h3) {
        // t: ubar4 L2slash (ptslash + M) L1slash v3;  u: ubar4 L1slash (puslash + M) L2slash v3
        sp4 wt = slash_apply(L1, v3[h3]), wu = slash_apply(L2, v3[h3]);
        sp4 xt = slash_apply(pt, wt), xu = slash_apply(pu, wu);
#pragma unroll
        for (int k = 0; k < 4; ++k) { xt.c[k] = fcadd(xt.c[k], fcmul(M, wt.c[k])); xu.c[k] = fcadd(xu.c[k], fcmul(M, wu.c[k])); }
        wt = slash_apply(L2, xt);
        wu = slash_apply(L1, xu);
#pragma unroll
        for (int h4 = 0; h4 < 2; ++h4) {
          const cx ct = sp_dot(ub4[h4], wt), cu = sp_dot(ub4[h4], wu);
          h[((ha * 2 + hb) * 2 + h3) * 2 + h4] = fcmk(St * ct.re + Su * cu.re, St * ct.im + Su * cu.im);
        }
      }
    }
  }
}
__device__ __forceinline__ double yy_qed_amp2(const NodeKin &in, const v4 &p3, const v4 &p4) {
  cx h[16];
  yy_qed_amp(in, p3, p4, h);
  double a2 = 0.0;
#pragma unroll
  for (int k = 0; k < 16; ++k) a2 += h[k].re * h[k].re + h[k].im * h[k].im;
  return 0.25 * a2;
}

}  // namespace grb
