// grb_yy.cuh -- gamma gamma -> l+ l- on the device: the 16 MadGraph5/HELAS helicity amplitudes of the reference
// (src/Amplitude/AMP_MG5_yy_ll.cc:41-255, HelAmps_sm.cc) with the photons put on the light cone by
// OffShell2LightCone (include/Graniitti/MKinematics.h:799-907) and the kt-EPA coherent fluxes
// (src/MFlux.cc:22-44, src/MForm.cc:664-684).
//
// Organisation: the two photon polarisation vectors, the two lepton spinors and the four + four off-shell
// fermion lines are built once per point (2+2+2+2+4+4 wavefunctions instead of 16 x 6), then the 16 x 2
// vertex contractions.  The ALOHA vertex algebra is written in light-cone components
// V+- = V0 +- V3, Vc = V1 + i V2, Vd = V1 - i V2 (the generated expressions reduce to bilinears in those).
#pragma once
#include "grb_res.cuh"

#ifndef GRB_YY_INLINE
#define GRB_YY_INLINE static __noinline__
#endif

namespace grb {

struct wf6 {  // HELAS wavefunction: [0..1] momentum, [2..5] components
  cx c[6];
};
__device__ __forceinline__ cx csub(const cx &a, const cx &b) { return {a.re - b.re, a.im - b.im}; }
__device__ __forceinline__ cx ci(const cx &a) { return {-a.im, a.re}; }
__device__ __forceinline__ double sgn_of(double a, double b) { return (b < 0) ? -fabs(a) : fabs(a); }

// massless vector, HelAmps_sm.cc:207-254; p = (E,px,py,pz)
__device__ __forceinline__ void helas_vxxxxx(const double p[4], int nhel, int nsv, wf6 &v) {
  const double sqh = sqrt(0.5), hel = (double)nhel;
  v.c[0] = {p[0] * nsv, p[3] * nsv};
  v.c[1] = {p[1] * nsv, p[2] * nsv};
  const double pp = p[0], pt = sqrt((p[1] * p[1]) + (p[2] * p[2]));
  v.c[2] = {0.0, 0.0};
  v.c[5] = {hel * pt / pp * sqh, 0.0};
  if (pt != 0.0) {
    const double pzpt = p[3] / (pp * pt) * sqh * hel;
    v.c[3] = {-p[1] * pzpt, -nsv * p[2] / pt * sqh};
    v.c[4] = {-p[2] * pzpt, nsv * p[1] / pt * sqh};
  } else {
    v.c[3] = {-hel * sqh, 0.0};
    v.c[4] = {0.0, nsv * sgn_of(sqh, p[3])};
  }
}
// fermion wavefunctions, HelAmps_sm.cc:140-205 (oxxxxx), :265-330 (ixxxxx); OUT selects oxxxxx
template <bool OUT>
__device__ __forceinline__ void helas_fermion(const double p[4], double fmass, int nhel, int nsf, wf6 &f) {
  const double sg = OUT ? 1.0 : -1.0;
  f.c[0] = {sg * p[0] * nsf, sg * p[3] * nsf};
  f.c[1] = {sg * p[1] * nsf, sg * p[2] * nsf};
  const int nh = nhel * nsf;
  const double yim = OUT ? -p[2] : p[2];  // chi[1] carries -i py for oxxxxx, +i py for ixxxxx
  cx chi[2];
  if (fmass != 0.0) {
    const double pp = fmin(p[0], sqrt(p[1] * p[1] + p[2] * p[2] + p[3] * p[3]));
    if (pp == 0.0) {
      double sqm[2];
      sqm[0] = sqrt(fabs(fmass)); sqm[1] = sgn_of(sqm[0], fmass);
      if (OUT) {
        const int ip = -((1 - nh) / 2) * nhel, im = (1 + nh) / 2 * nhel;
        f.c[2] = {im * sqm[abs(ip)], 0.0}; f.c[3] = {ip * nsf * sqm[abs(ip)], 0.0};
        f.c[4] = {im * nsf * sqm[abs(im)], 0.0}; f.c[5] = {ip * sqm[abs(im)], 0.0};
      } else {
        const int ip = (1 + nh) / 2, im = (1 - nh) / 2;
        f.c[2] = {ip * sqm[ip], 0.0}; f.c[3] = {im * nsf * sqm[ip], 0.0};
        f.c[4] = {ip * nsf * sqm[im], 0.0}; f.c[5] = {im * sqm[im], 0.0};
      }
      return;
    }
    double sf[2], omega[2], sfo[2];
    sf[0] = (1 + nsf + (1 - nsf) * nh) * 0.5; sf[1] = (1 + nsf - (1 - nsf) * nh) * 0.5;
    omega[0] = sqrt(p[0] + pp); omega[1] = fmass / omega[0];
    const int ip = (1 + nh) / 2, im = (1 - nh) / 2;
    sfo[0] = sf[0] * omega[ip]; sfo[1] = sf[1] * omega[im];
    const double pp3 = fmax(pp + p[3], 0.0);
    chi[0] = {sqrt(pp3 * 0.5 / pp), 0.0};
    if (pp3 == 0.0) chi[1] = {(double)-nh, 0.0};
    else { const double dn = sqrt(2.0 * pp * pp3); chi[1] = {nh * p[1] / dn, yim / dn}; }
    if (OUT) {
      f.c[2] = cscale(chi[im], sfo[1]); f.c[3] = cscale(chi[ip], sfo[1]);
      f.c[4] = cscale(chi[im], sfo[0]); f.c[5] = cscale(chi[ip], sfo[0]);
    } else {
      f.c[2] = cscale(chi[im], sfo[0]); f.c[3] = cscale(chi[ip], sfo[0]);
      f.c[4] = cscale(chi[im], sfo[1]); f.c[5] = cscale(chi[ip], sfo[1]);
    }
  } else {
    double sq;
    if (p[1] == 0.0 && p[2] == 0.0 && p[3] < 0.0) sq = 0.0; else sq = sqrt(fmax(p[0] + p[3], 0.0)) * nsf;
    chi[0] = {sq, 0.0};
    if (sq == 0.0) chi[1] = {-nhel * sqrt(2.0 * p[0]), 0.0};
    else chi[1] = {nh * p[1] / sq, yim / sq};
    const cx z = {0.0, 0.0};
    if (OUT) {
      if (nh == 1) { f.c[2] = chi[0]; f.c[3] = chi[1]; f.c[4] = z; f.c[5] = z; }
      else { f.c[2] = z; f.c[3] = z; f.c[4] = chi[1]; f.c[5] = chi[0]; }
    } else {
      if (nh == 1) { f.c[2] = z; f.c[3] = z; f.c[4] = chi[0]; f.c[5] = chi[1]; }
      else { f.c[2] = chi[1]; f.c[3] = chi[0]; f.c[4] = z; f.c[5] = z; }
    }
  }
}

struct lc4 {  // light-cone components of a complex 4-vector wavefunction
  cx p, m, c, d;
};
__device__ __forceinline__ lc4 lightcone(const wf6 &V) {
  lc4 r;
  r.p = cadd(V.c[2], V.c[5]); r.m = csub(V.c[2], V.c[5]);
  r.c = cadd(V.c[3], ci(V.c[4])); r.d = csub(V.c[3], ci(V.c[4]));
  return r;
}
// FFV1_0 (HelAmps_sm.cc:678-687) without the coupling: F1 (gamma.V) F2
// (GRB_YY_INLINE: 32 inlined copies of this and 8 of FFV1_1 / FFV1_2 made a hot path of 8 900 instructions = 139 KB per point,
// far beyond the instruction cache: 42 % of the stall samples were instruction fetches; as real functions the hot code shrinks)
__device__ GRB_YY_INLINE cx ffv1_0(const wf6 &F1, const wf6 &F2, const lc4 &v) {
  cx t = cmul(F1.c[2], cadd(cmul(F2.c[4], v.p), cmul(F2.c[5], v.c)));
  t = cadd(t, cmul(F1.c[3], cadd(cmul(F2.c[4], v.d), cmul(F2.c[5], v.m))));
  t = cadd(t, cmul(F1.c[4], csub(cmul(F2.c[2], v.m), cmul(F2.c[3], v.c))));
  t = cadd(t, cmul(F1.c[5], csub(cmul(F2.c[3], v.p), cmul(F2.c[2], v.d))));
  return t;
}
// FFV1_2 / FFV1_1 (HelAmps_sm.cc:435-482, :930-977) with zero propagator mass (AMP_MG5_yy_ll.cc:247-248),
// without the coupling: i (gamma.P)(gamma.V) F / P^2 in light-cone form
__device__ GRB_YY_INLINE void ffv1_2(const wf6 &F1, const wf6 &V3, const lc4 &v, wf6 &F2) {
  F2.c[0] = cadd(F1.c[0], V3.c[0]); F2.c[1] = cadd(F1.c[1], V3.c[1]);
  const double P0 = -F2.c[0].re, P1 = -F2.c[1].re, P2 = -F2.c[1].im, P3 = -F2.c[0].im;
  const double inv = 1.0 / ((P0 * P0) - (P1 * P1) - (P2 * P2) - (P3 * P3));
  const double Pp = P0 + P3, Pm = P0 - P3;
  const cx Pc = {P1, P2}, Pd = {P1, -P2};
  const cx a = cadd(cmul(F1.c[2], v.p), cmul(F1.c[3], v.d)), b = cadd(cmul(F1.c[2], v.c), cmul(F1.c[3], v.m));
  const cx c = csub(cmul(F1.c[4], v.m), cmul(F1.c[5], v.d)), e = csub(cmul(F1.c[4], v.c), cmul(F1.c[5], v.p));
  F2.c[2] = cscale(ci(csub(cscale(a, Pm), cmul(Pd, b))), inv);
  F2.c[3] = cscale(ci(csub(cscale(b, Pp), cmul(Pc, a))), inv);
  F2.c[4] = cscale(ci(csub(cscale(c, Pp), cmul(Pd, e))), inv);
  F2.c[5] = cscale(ci(csub(cmul(Pc, c), cscale(e, Pm))), inv);
}
__device__ GRB_YY_INLINE void ffv1_1(const wf6 &F2, const wf6 &V3, const lc4 &v, wf6 &F1) {
  F1.c[0] = cadd(F2.c[0], V3.c[0]); F1.c[1] = cadd(F2.c[1], V3.c[1]);
  const double P0 = -F1.c[0].re, P1 = -F1.c[1].re, P2 = -F1.c[1].im, P3 = -F1.c[0].im;
  const double inv = 1.0 / ((P0 * P0) - (P1 * P1) - (P2 * P2) - (P3 * P3));
  const double Pp = P0 + P3, Pm = P0 - P3;
  const cx Pc = {P1, P2}, Pd = {P1, -P2};
  const cx a = csub(cmul(F2.c[2], v.m), cmul(F2.c[3], v.c)), b = csub(cmul(F2.c[2], v.d), cmul(F2.c[3], v.p));
  const cx c = cadd(cmul(F2.c[4], v.p), cmul(F2.c[5], v.c)), e = cadd(cmul(F2.c[4], v.d), cmul(F2.c[5], v.m));
  F1.c[2] = cscale(ci(csub(cmul(Pc, b), cscale(a, Pp))), inv);
  F1.c[3] = cscale(ci(csub(cscale(b, Pm), cmul(Pd, a))), inv);
  F1.c[4] = cscale(ci(csub(cmul(Pc, e), cscale(c, Pm))), inv);
  F1.c[5] = cscale(ci(csub(cmul(Pd, c), cscale(e, Pp))), inv);
}

__device__ __forceinline__ double p3mod_of(const v4 &a) { return msqrt(a.x * a.x + a.y * a.y + a.z * a.z); }

// OffShell2LightCone for two final-state particles (reference MKinematics.h:799-907)
__device__ __forceinline__ void offshell2lightcone(v4 &q1, v4 &q2, v4 &pa, v4 &pb) {
  q1.e = p3mod_of(q1);
  q2.e = p3mod_of(q2);
  v4 q;
  q.x = q1.x + q2.x; q.y = q1.y + q2.y; q.z = q1.z + q2.z; q.e = q1.e + q2.e;
  const double ma = m_of(pa), mb = m_of(pb);
  for (int iter = 1;; ++iter) {
    const double dE = (pa.e + pb.e) - q.e;
    {
      const double Ea = pa.e - dE / 2, Eb = pb.e - dE / 2;
      const double sa = msqrt(Ea * Ea - ma * ma) / p3mod_of(pa), sb = msqrt(Eb * Eb - mb * mb) / p3mod_of(pb);
      pa.x = sa * pa.x; pa.y = sa * pa.y; pa.z = sa * pa.z; pa.e = Ea;
      pb.x = sb * pb.x; pb.y = sb * pb.y; pb.z = sb * pb.z; pb.e = Eb;
    }
    const double dx = ((pa.x + pb.x) - q.x) * (1.0 / 2), dy = ((pa.y + pb.y) - q.y) * (1.0 / 2), dz = ((pa.z + pb.z) - q.z) * (1.0 / 2);
    pa.x -= dx; pa.y -= dy; pa.z -= dz; pa.e = msqrt((pa.x * pa.x + pa.y * pa.y + pa.z * pa.z) + ma * ma);
    pb.x -= dx; pb.y -= dy; pb.z -= dz; pb.e = msqrt((pb.x * pb.x + pb.y * pb.y + pb.z * pb.z) + mb * mb);
    const double IVM = fabs(m2_of(q.x - (pa.x + pb.x), q.y - (pa.y + pb.y), q.z - (pa.z + pb.z), q.e - (pa.e + pb.e)));
    if (iter > 15 || IVM < 1e-10) break;
  }
}

// |M|^2 (spin averaged) x kt-EPA fluxes for yy[CON] -> l+ l-  (reference PROC_3, MSubProc.h:165-173).
// VEC: also return the 16 helicity amplitudes as the reference leaves them in lts.hamp (MG5 amplitude x GC_3^2 x
// sqrt(flux_1/x_1) sqrt(flux_2/x_2), src/MFlux.cc:22-44) for the screening loop, where sum |.|^2 / 4 is taken last.
template <bool VEC>
__device__ __forceinline__ double yy_ll_core(const NodeKin &in, const v4 &p3, const v4 &p4, cx *hv) {
  const DevProc &P = c_proc;
  v4 q1, q2, pa = p3, pb = p4;
  q1.x = 0.0 - in.p1.x; q1.y = 0.0 - in.p1.y; q1.z = P.pb1z - in.p1.z; q1.e = P.pb1e - in.p1.e;
  q2.x = 0.0 - in.p2.x; q2.y = 0.0 - in.p2.y; q2.z = P.pb2z - in.p2.z; q2.e = P.pb2e - in.p2.e;
  const double qt1 = msqrt(q1.x * q1.x + q1.y * q1.y), qt2 = msqrt(q2.x * q2.x + q2.y * q2.y);
  const double x1 = 1 - in.p1.z / P.pb1z, x2 = 1 - in.p2.z / P.pb2z;
  const double m3 = m_of(p3), m4 = m_of(p4);
  offshell2lightcone(q1, q2, pa, pb);
  const double k1[4] = {q1.e, q1.x, q1.y, q1.z}, k2[4] = {q2.e, q2.x, q2.y, q2.z};
  const double k3[4] = {pa.e, pa.x, pa.y, pa.z}, k4[4] = {pb.e, pb.x, pb.y, pb.z};
  wf6 w2[2], w3[2];
  lc4 v1[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int hel = 2 * h - 1;
    wf6 w1;
    helas_vxxxxx(k2, hel, -1, w1);
    helas_fermion<false>(k3, m3, hel, -1, w2[h]);
    helas_fermion<true>(k4, m4, hel, +1, w3[h]);
    v1[h] = lightcone(w1);
  }
  double sum = 0.0;
  // The helicity of the first photon is a real loop (not unrolled) when only |M|^2 is needed: one polarisation vector and
  // its four off-shell fermion lines are live at a time instead of both sets (register pressure: the kernel runs at 255
  // registers with local-memory spills).  Same operations in the same order (h0, h2, h3, h1) as the unrolled form.
#pragma unroll(VEC ? 2 : 1)
  for (int h0 = 0; h0 < 2; ++h0) {
    wf6 w0;
    helas_vxxxxx(k1, 2 * h0 - 1, -1, w0);
    const lc4 v0 = lightcone(w0);
    wf6 w4[2], w5[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      ffv1_2(w2[h], w0, v0, w4[h]);
      ffv1_1(w3[h], w0, v0, w5[h]);
    }
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
      for (int h3 = 0; h3 < 2; ++h3)
#pragma unroll
        for (int h1 = 0; h1 < 2; ++h1) {
          const cx a = cadd(ffv1_0(w4[h2], w3[h3], v1[h1]), ffv1_0(w2[h2], w5[h3], v1[h1]));
          sum += a.re * a.re + a.im * a.im;
          if (VEC) hv[((h0 * 2 + h2) * 2 + h3) * 2 + h1] = a;
        }
  }
  // couplings: each diagram carries GC_3^2 (off-shell line) and the vertex -i GC_3 ... |GC_3|^4 = ee^4 overall;
  // GC_3 = -i ee, ee = 2 sqrt(pi/137.03599908) (Parameters_sm.cc:93-101)
  const double amp2 = sum * P.yy_ee4 / 4.0;
  // sqrt(flux_i / x_i): coherent flux of an intact proton, incoherent flux of an excited N* system (MFlux.cc:22-28)
  const double f1 = (in.exc & 1) ? incoh_flux_amp(x1, in.t1, qt1, m2_of(in.p1)) : coh_flux_amp(x1, in.t1, qt1);
  const double f2 = (in.exc & 2) ? incoh_flux_amp(x2, in.t2, qt2, m2_of(in.p2)) : coh_flux_amp(x2, in.t2, qt2);
  if (VEC) {
    const double sc = sqrt(P.yy_ee4) * (f1 * f2);
#pragma unroll
    for (int k = 0; k < 16; ++k) { hv[k].re *= sc; hv[k].im *= sc; }
  }
  return (f1 * f1) * (f2 * f2) * amp2;
}
__device__ __forceinline__ double yy_ll_amp2(const NodeKin &in, const v4 &p3, const v4 &p4) {
  return yy_ll_core<false>(in, p3, p4, nullptr);
}

}  // namespace grb
