// grb_res.cuh -- resonance amplitudes of the Regge model on the device: MRegge::ME3 / PhotoME3
// (reference src/MRegge.cc:677-706,781-882) with the Jacob-Wick spin weights of spin::ProdAmp / DecayAmp
// (src/MSpin.cc:280-483), Breit-Wigner propagators (src/MForm.cc:743-827) and the kt-EPA coherent flux
// (src/MForm.cc:664-684).
//
// Not a translation of the reference's MMatrix / pow() code: Wigner small-d functions are evaluated from a
// host-built coefficient table with integer powers of cos(theta/2), sin(theta/2); the rotated density matrix
// D^+ rho D is formed once per resonance and shared by the production and decay weights; the two Regge
// propagators and the proton form factors are shared by all resonances of a point.
#pragma once
#include "grb_dev.cuh"

namespace grb {

extern __constant__ DevRes c_res;  // defined in grb_consts.cu

__device__ __forceinline__ cx cmul(const cx &a, const cx &b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
__device__ __forceinline__ cx cmulc(const cx &a, const cx &b) { return {a.re * b.re + a.im * b.im, a.im * b.re - a.re * b.im}; }  // a * conj(b)
__device__ __forceinline__ cx cadd(const cx &a, const cx &b) { return {a.re + b.re, a.im + b.im}; }
__device__ __forceinline__ cx cscale(const cx &a, double s) { return {a.re * s, a.im * s}; }
__device__ __forceinline__ cx cdiv(const cx &a, const cx &b) {
  const double d = 1.0 / (b.re * b.re + b.im * b.im);
  return {(a.re * b.re + a.im * b.im) * d, (a.im * b.re - a.re * b.im) * d};
}
__device__ __forceinline__ double theta_of(const v4 &a) {  // M4Vec::Theta
  return (a.x == 0.0 && a.y == 0.0 && a.z == 0.0) ? 0.0 : atan2(msqrt(pt2_of(a)), a.z);
}
__device__ __forceinline__ double phi_of(const v4 &a) { return (a.x == 0.0 && a.y == 0.0) ? 0.0 : atan2(a.y, a.x); }

// LorentzBoost(mother, M0, p, sign = -1): into the rest frame of `mother` (reference MKinematics.h:383-401)
__device__ __forceinline__ v4 boost_to_rest(const v4 &mother, double M0, const v4 &p) {
  const double b0 = -mother.x / mother.e, b1 = -mother.y / mother.e, b2 = -mother.z / mother.e;
  const double gamma  = mother.e / M0;
  const double kappa1 = (b2 * p.z + b1 * p.y) + b0 * p.x;
  const double kappa2 = gamma * (gamma * kappa1 / (1.0 + gamma) + p.e);
  v4 r;
  r.x = p.x + kappa2 * b0; r.y = p.y + kappa2 * b1; r.z = p.z + kappa2 * b2; r.e = gamma * (p.e + kappa1);
  return r;
}

// Wigner small d^J_{m,mp}(theta), index i = m + J, j = mp + J (reference src/MSpin.cc:927-970); N = 2J+1
template <int N>
__device__ __forceinline__ void wigner_small_d(double theta, double d[N][N]) {
  constexpr int J2 = N - 1;  // 2J
  double sh, ch;
  sincos(0.5 * theta, &sh, &ch);
  double cp[N + N], sp[N + N];  // powers 0..2J
  cp[0] = 1.0; sp[0] = 1.0;
#pragma unroll
  for (int k = 1; k <= J2; ++k) { cp[k] = cp[k - 1] * ch; sp[k] = sp[k - 1] * sh; }
  const double *coef = c_res.wd[J2 / 2 - 1];
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < N; ++k) {
        // exponents: cos^(2J + m - mp - 2k) sin^(mp - m + 2k), m = i - J, mp = j - J
        const int a = J2 + i - j - 2 * k, b = j - i + 2 * k;
        if (a >= 0 && b >= 0 && a <= J2 && b <= J2) s = fma(coef[(i * 5 + j) * 5 + k], cp[a] * sp[b], s);
      }
      d[i][j] = s;
    }
}

// e_j = exp(i M_j phi), M_j = j - J (N = 2J + 1 = 3, 5): one sincos and the double-angle identities instead of N calls of
// sincos(M phi) (the reference evaluates each phase on its own; the identities agree with that to 2 ulp, and the kernel is
// bound by instruction fetch: each inlined double-precision sincos is ~80 instructions of code)
template <int N>
__device__ __forceinline__ void azimuthal_phases(double phi, cx e[N]) {
  constexpr int J = (N - 1) / 2;
  double s1, c1;
  sincos(phi, &s1, &c1);
  e[J] = {1.0, 0.0};
  e[J + 1] = {c1, s1};
  e[J - 1] = {c1, -s1};
  if (N == 5) {
    const double c2 = (c1 - s1) * (c1 + s1), s2 = 2.0 * s1 * c1;
    e[J + 2 < N ? J + 2 : 0] = {c2, s2};
    e[J - 2 >= 0 ? J - 2 : 0] = {c2, -s2};
  }
}

// rho_ROT = D^+ rho D with D_{m mp} = d_{m mp}(theta) e^{i mp phi}
template <int N>
__device__ __forceinline__ void rotated_rho(const DevResonance &r, double th, double ph, cx out[N][N]) {
  double d[N][N];
  wigner_small_d<N>(th, d);
  cx e[N];
  azimuthal_phases<N>(ph, e);
  // t = rho * D : t[k][j] = sum_l rho[k][l] d[l][j] e_j
  cx t[N][N];
#pragma unroll
  for (int k = 0; k < N; ++k)
#pragma unroll
    for (int j = 0; j < N; ++j) {
      double sr = 0.0, si = 0.0;
#pragma unroll
      for (int l = 0; l < N; ++l) { sr = fma(r.rho_re[k * 5 + l], d[l][j], sr); si = fma(r.rho_im[k * 5 + l], d[l][j], si); }
      t[k][j] = cmul({sr, si}, e[j]);
    }
  // out[i][j] = sum_k conj(D[k][i]) t[k][j] = conj(e_i) sum_k d[k][i] t[k][j]
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) {
      double sr = 0.0, si = 0.0;
#pragma unroll
      for (int k = 0; k < N; ++k) { sr = fma(d[k][i], t[k][j].re, sr); si = fma(d[k][i], t[k][j].im, si); }
      out[i][j] = cmulc({sr, si}, e[i]);
    }
}

// Re[ f rho f^+ ] for one row f_j = d_{lambda, M_j}(theta) e^{i M_j phi} * T  (|T|^2 factored out by the caller)
template <int N>
__device__ __forceinline__ double row_weight(const cx rho[N][N], const double d[N][N], int row, double phi) {
  cx f[N], e[N];
  azimuthal_phases<N>(phi, e);
#pragma unroll
  for (int j = 0; j < N; ++j) {
    f[j].re = d[row][j] * e[j].re;
    f[j].im = d[row][j] * e[j].im;
  }
  double acc = 0.0;
#pragma unroll
  for (int j = 0; j < N; ++j) {
    cx t = {0.0, 0.0};
#pragma unroll
    for (int k = 0; k < N; ++k) t = cadd(t, cmul(f[k], rho[k][j]));
    acc += t.re * f[j].re + t.im * f[j].im;  // Re(t * conj(f_j))
  }
  return acc;
}

struct ResKin {  // what a resonance needs from the point (current loop node or Born kinematics)
  double s1, s2, t1, t2, m2, x1, x2, qt1, qt2;
  v4     pX, q1, p1, p2, d0, d1;  // central system, upper propagator, protons, daughters (not loop-updated)
  int    exc;                     // lts.excite1/2 (bits 0, 1): forward N* systems
};

// rotation angles of the density matrix frame (reference src/MSpin.cc:544-601)
__device__ __forceinline__ void rho_rotation(const ResKin &K, double MX, int frame, double &th, double &ph) {
  const DevProc &P = c_proc;
  if (frame == 0) { th = 0.0; ph = 0.0; }
  else if (frame == 1) { th = theta_of(K.pX); ph = phi_of(K.pX); }
  else if (frame == 2) {  // Collins-Soper: bisector of the boosted beams
    v4 ba, bb;
    ba.x = 0.0; ba.y = 0.0; ba.z = P.pb1z; ba.e = P.pb1e;
    bb.x = 0.0; bb.y = 0.0; bb.z = P.pb2z; bb.e = P.pb2e;
    const v4 a = boost_to_rest(K.pX, MX, ba), b = boost_to_rest(K.pX, MX, bb);
    const double na = msqrt(a.x * a.x + a.y * a.y + a.z * a.z), nb = msqrt(b.x * b.x + b.y * b.y + b.z * b.z);
    v4 z;
    z.x = a.x / na - b.x / nb; z.y = a.y / na - b.y / nb; z.z = a.z / na - b.z / nb; z.e = 0.0;
    const double nz = msqrt(z.x * z.x + z.y * z.y + z.z * z.z);
    z.x /= nz; z.y /= nz; z.z /= nz;
    th = theta_of(z); ph = phi_of(z);
  } else {
    const v4 q = boost_to_rest(K.pX, MX, K.q1);
    th = theta_of(q); ph = phi_of(q);
  }
}

// ProdAmp * DecayAmp spin weight for J = (N-1)/2 > 0 (both share D^+ rho D)
template <int N>
__device__ __forceinline__ cx spin_weights(const DevResonance &r, const ResKin &K, double MX) {
  cx A = {r.g_re, r.g_im};
  double Adec = r.g_decay;
  const bool need_prod = r.spingen && !(r.spinX2 == 0);
  const bool need_dec  = r.spindec;
  if (!need_prod && !need_dec) return cscale(A, Adec);
  double th, ph;
  rho_rotation(K, MX, r.frame, th, ph);
  cx rho[N][N];
  rotated_rho<N>(r, th, ph, rho);
  const double norm = sqrt((double)N);
  constexpr int J = (N - 1) / 2;
  if (need_prod) {
    const v4 q = boost_to_rest(K.pX, MX, K.q1);
    double d[N][N];
    wigner_small_d<N>(theta_of(q), d);
    const double phq = phi_of(q);
    double tr;
    if (N == 3 && r.P == 1) {  // 1+: s1 = s2 = 1, T0 of MSpin.cc:339-343 (rows lambda1, cols lambda2)
      const double T0[3][3] = {{0.0, -0.5, 0.0}, {0.5, 0.0, -0.5}, {0.0, 0.5, 0.0}};
      tr = 0.0;
#pragma unroll
      for (int mu = 0; mu < 3; ++mu)
#pragma unroll
        for (int nu = 0; nu < 3; ++nu) {
          const int lam = (mu - 1) - (nu - 1);
          if (T0[mu][nu] != 0.0 && lam >= -J && lam <= J) tr += T0[mu][nu] * T0[mu][nu] * row_weight<N>(rho, d, lam + J, phq);
        }
    } else if (N == 3 && r.P == -1) {  // 1-: s1 = 0, s2 = 1, T0 = (-0.697, -0.169, 0.697); lambda = -lambda2
      tr = 0.697 * 0.697 * row_weight<N>(rho, d, 1 + J, phq) + 0.169 * 0.169 * row_weight<N>(rho, d, 0 + J, phq) +
           0.697 * 0.697 * row_weight<N>(rho, d, -1 + J, phq);
    } else {  // s1 = s2 = 0, T0 = 1
      tr = row_weight<N>(rho, d, J, phq);
    }
    A = cscale(A, norm * msqrt(tr));
  }
  if (need_dec) {  // spinless daughters: one row, lambda = 0, T = T_00
    const v4 dd = boost_to_rest(K.pX, MX, K.d0);
    double d[N][N];
    wigner_small_d<N>(theta_of(dd), d);
    const double t2 = r.T_re[0] * r.T_re[0] + r.T_im[0] * r.T_im[0];
    Adec *= norm * msqrt(t2 * row_weight<N>(rho, d, J, phi_of(dd)));
  }
  return cscale(A, Adec);
}

__device__ __forceinline__ cx spin_amp(const DevResonance &r, const ResKin &K, double MX) {
  if (r.spinX2 == 0) {
    cx A = {r.g_re, r.g_im};
    if (r.spingen && r.P == -1) {  // pseudoscalar: |t1 t2|^(1/2) sin(dphi), MSpin.cc:301-331
      double D = phi_of(K.p1) - phi_of(K.p2);
      while (D >= GRB_PI) D -= 2.0 * GRB_PI;
      while (D < -GRB_PI) D += 2.0 * GRB_PI;
      A = cscale(A, sqrt(fabs(K.t1)) * sqrt(fabs(K.t2)) * sin(fabs(D)));
    }
    const cx T = {r.T_re[0], r.T_im[0]};
    return r.spindec ? cscale(cmul(A, T), r.g_decay) : cscale(A, r.g_decay);
  }
  if (r.spinX2 == 2) return spin_weights<3>(r, K, MX);
  return spin_weights<5>(r, K, MX);
}

// Breit-Wigner propagators (reference src/MForm.cc:743-827); BW 3 needs the daughter masses
__device__ __forceinline__ cx breit_wigner(const DevResonance &r, double m2, double mA, double mB) {
  const double M0 = r.mass, G = r.width, J = 0.5 * r.spinX2;
  cx den;
  double num = -1.0;
  den.re = m2 - M0 * M0;
  if (r.BW == 1) den.im = M0 * G;
  else if (r.BW == 2) den.im = sqrt(m2) * G;
  else if (r.BW == 3) {
    const double u = msqrt((m2 - pow2(mA + mB)) * (m2 - pow2(mA - mB))) / (2 * msqrt(m2));
    const double d = msqrt((M0 * M0 - pow2(mA + mB)) * (M0 * M0 - pow2(mA - mB))) / (2 * M0);
    den.im = G * M0 * M0 / msqrt(m2) * pow(u / d, (double)(2 * (int)J + 1));
  } else {
    den.im = M0 * G;
    const double x = m2 / (M0 * M0);
    if ((int)J == 0) num = -1.0;
    else if ((int)(J * 2) == 1) num = -2.0 * msqrt(m2);
    else if ((int)J == 1) num = -(1.0 - x);
    else if ((int)(J * 2) == 3) num = -(2.0 / 3.0) * msqrt(m2) * (1.0 - x);
    else if ((int)J == 2) num = -(7.0 / 6.0 - (4.0 / 3.0) * x + (2.0 / 3.0) * x * x);
  }
  return cdiv({num, 0.0}, den);
}

__device__ __forceinline__ void sachs(double Q2, double &GE, double &GM) {  // MForm.cc:560-645
  const DevProc &P   = c_proc;
  const double   tau = Q2 / P.four_mp2;
  if (P.em_kelly) {
    GE = (1.0 - 0.24 * tau) / (1.0 + 10.98 * tau + 12.82 * (tau * tau) + 21.97 * (tau * (tau * tau)));
    GM = 2.792847337 * (1.0 + 0.12 * tau) / (1.0 + 10.97 * tau + 18.86 * (tau * tau) + 6.55 * (tau * (tau * tau)));
  } else {
    GM = 2.792847337 / pow2(1.0 + Q2 / 0.71);
    GE = GM / 2.792847337;
  }
}
// sqrt(CohFlux(xi,t,pt)/xi), reference src/MForm.cc:664-684 and MRegge.cc:795-802
__device__ __forceinline__ double coh_flux_amp(double xi, double t, double pt) {
  const DevProc &P  = c_proc;
  const double pt2 = pt * pt, xi2 = xi * xi, mp2 = 0.25 * P.four_mp2, Q2 = fabs(t);
  double GE, GM;
  sachs(Q2, GE, GM);
  const double PART1 = (4.0 * mp2 * GE * GE + Q2 * GM * GM) / (4.0 * mp2 + Q2);
  const double DELTA = pt2 / (pt2 + xi2 * mp2);
  double f = 0.0072973525693 / GRB_PI * ((1.0 - xi) * DELTA * DELTA * PART1 + (xi2 / 4.0) * DELTA * GM * GM);
  f = f / xi / pt2 * (16.0 * (GRB_PI * GRB_PI));
  return msqrt(f / xi);
}

// sqrt(IncohFlux(xi,t,pt,M2)/xi): photon flux of a dissociating proton (reference src/MForm.cc:700-719) with the proton
// structure function F2(x, Q^2) of PARAM_STRUCTURE.F2 (CKMT or Donnachie-Landshoff, :463-508), F1 = F2 / 2x
__device__ __forceinline__ double f2xq2(double xbj, double Q2) {
  if (c_proc.f2_model == 1) {
    return 0.324 * pow(xbj, -0.0808) * pow(Q2 / (Q2 + 0.561991692786383), 1 + 0.0808) +
           0.098 * pow(xbj, 1 - 0.5475) * pow(Q2 / (Q2 + 0.011133), 0.5475);
  }
  const double n_Q2 = (3.0 / 2.0) * (1 + Q2 / (Q2 + 3.5489)), DELTA_Q2 = 0.0800 * (1 + (2 * Q2) / (Q2 + 1.1170));
  const double C1 = pow(Q2 / (Q2 + 0.2631), 1.0 + DELTA_Q2), C2 = pow(Q2 / (Q2 + 0.6452), 0.4150);
  return 0.1502 * pow(xbj, -DELTA_Q2) * pow(1 - xbj, n_Q2 + 4.0) * C1 +
         pow(xbj, 1.0 - 0.4150) * (1.2064 * pow(1 - xbj, n_Q2) + 0.1798 * pow(1 - xbj, n_Q2 + 1.0)) * C2;
}
__device__ __forceinline__ double incoh_flux_amp(double xi, double t, double pt, double M2) {
  const double mp2 = 0.25 * c_proc.four_mp2, pt2 = pt * pt, xi2 = xi * xi, Q2 = fabs(t);
  const double xbj = Q2 / (Q2 + M2 - mp2), DELTA = pt2 / (pt2 + xi * (M2 - mp2) + xi2 * mp2);
  const double F2 = f2xq2(xbj, Q2), F1 = F2 / (2.0 * xbj);
  double f = 0.0072973525693 / GRB_PI * ((1.0 - xi) * (DELTA * DELTA) * F2 / (Q2 + M2 - mp2) +
                                         (xi2 / (4.0 * (xbj * xbj))) * DELTA * 2.0 * xbj * F1 / (Q2 + M2 - mp2));
  f = f / xi / pt2 * (16.0 * (GRB_PI * GRB_PI));
  return msqrt(f / xi);
}

// Regge single-trajectory propagator pieces shared by all resonances: Prop(s,t) = eta0 cis(pc t) (s/s0)^alpha(t)
__device__ __forceinline__ cx regge_prop_k(double s, double t, int k, bool with_s0 = true) {
  const DevProc &P = c_proc;
  const double mag = exp((P.a0[k] + P.ap[k] * t) * (log(s) + (with_s0 ? P.ln_inv_s0 : 0.0)));
  double sn, cs;
  sincos(P.phase_c[k] * t, &sn, &cs);
  return {(P.eta0_re[k] * cs - P.eta0_im[k] * sn) * mag, (P.eta0_re[k] * sn + P.eta0_im[k] * cs) * mag};
}
__device__ __forceinline__ cx regge_prop_all(double s, double t) {
  const DevProc &P = c_proc;
  cx M = {0.0, 0.0};
#pragma unroll 1
  for (int k = 0; k < 3; ++k)
    if (P.regge_n[k]) M = cadd(M, regge_prop_k(s, t, k));
  return M;
}

// coherent sum over the resonances (reference PROC_9 / PROC_17, MSubProc.h:233-262,399-436)
__device__ __forceinline__ cx resonance_sum(const ResKin &K) {
  const DevProc &P = c_proc;
  const DevRes & R = c_res;
  const double MX = (K.m2 > 0.0) ? msqrt(K.m2) : -msqrt(-K.m2);
  // shared by every ME3 term
  // PomPomProtonVertex (MRegge.cc:637-644): gN_P S3F(t) for an intact proton, sqrt(g3P gN_P) S3FINEL(t, M^2) for an excited system
  const double M2f1 = (K.exc & 1) ? m2_of(K.p1) : 0.0, M2f2 = (K.exc & 2) ? m2_of(K.p2) : 0.0;
  const double VX1 = (K.exc & 1) ? P.gN_P * P.exc_vertex * pow(fabs(K.t1) / (M2f1 * (fabs(K.t1) + 0.5616)), 0.5 * (1 + 0.0808)) : 0.0;
  const double VX2 = (K.exc & 2) ? P.gN_P * P.exc_vertex * pow(fabs(K.t2) / (M2f2 * (fabs(K.t2) + 0.5616)), 0.5 * (1 + 0.0808)) : 0.0;
  const double FFA = (K.exc & 1) ? VX1 : P.gN_P / (fma(-K.t1, P.inv_fc1, 1.0) * fma(-K.t1, P.inv_fc2, 1.0));
  const double FFB = (K.exc & 2) ? VX2 : P.gN_P / (fma(-K.t2, P.inv_fc1, 1.0) * fma(-K.t2, P.inv_fc2, 1.0));
  const cx P1 = regge_prop_all(K.s1, K.t1), P2 = regge_prop_all(K.s2, K.t2);
  const cx PP = cscale(cmul(P1, P2), 2.0 * FFA * FFB * exp(-P.omega * log(K.m2)));  // x (1/m2)^omega
  // shared by every PhotoME3 term (pomeron trajectory k = 0)
  cx Ph1 = {0.0, 0.0}, Ph2 = {0.0, 0.0};
  double gf1 = 0.0, gf2 = 0.0;
  if (R.any_photo) {
    Ph1 = regge_prop_k(K.s1, K.t1, 0, false);  // PhotoProp normalises to W0^2 = 1, not s0 (MRegge.cc:864-866)
    Ph2 = regge_prop_k(K.s2, K.t2, 0, false);
    gf1 = (K.exc & 1) ? incoh_flux_amp(K.x1, K.t1, K.qt1, M2f1) : coh_flux_amp(K.x1, K.t1, K.qt1);   // MRegge.cc:793-800
    gf2 = (K.exc & 2) ? incoh_flux_amp(K.x2, K.t2, K.qt2, M2f2) : coh_flux_amp(K.x2, K.t2, K.qt2);
  }
  const double mA = R.bw3 ? m_of(K.d0) : 0.0, mB = R.bw3 ? m_of(K.d1) : 0.0;
  cx A = {0.0, 0.0};
#pragma unroll 1
  for (int i = 0; i < R.n; ++i) {
    const DevResonance &r = R.r[i];
    const double dm = K.m2 - r.M2;
    const cx core = cscale(breit_wigner(r, K.m2, mA, mB), exp(-dm * dm * r.inv_gFF4));
    const cx spin = spin_amp(r, K, MX);
    cx term;
    if (r.photo) {
      // gammaflux2 * Prop(s1,t1) gN exp(B0 t1 / 2) + gammaflux1 * Prop(s2,t2) gN exp(B0 t2 / 2)
      // (PhotoProp's vertex, MRegge.cc:876-880: the inelastic vertex replaces gN exp(B0 t / 2) on an excited leg)
      const cx a1 = cscale(Ph1, gf2 * ((K.exc & 1) ? VX1 : P.gN_P * exp(0.5 * r.B0 * K.t1)));
      const cx a2 = cscale(Ph2, gf1 * ((K.exc & 2) ? VX2 : P.gN_P * exp(0.5 * r.B0 * K.t2)));
      term = cmul(core, cadd(a1, a2));
    } else {
      term = cmul(core, PP);
    }
    A = cadd(A, cmul(term, spin));
  }
  return A;
}

}  // namespace grb
