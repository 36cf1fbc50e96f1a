/* oracle/grb_oracle_eikonal.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the eikonal table construction of the reference: proton opacity Omega(s,b) ("MBT") and
 * screening amplitude A_eik(s,kt^2) ("MSA") by Fourier-Bessel quadrature (reference src/MEikonal.cc:129-242
 * SingleAmpElastic / S3Density / S3Screening, :379-401 S3CalculateArray, :404-436 15-digit ASCII round trip,
 * :262-331 S3CalcXS), with the non-linear pomeron trajectory of src/MForm.cc:402-431 and the Numerical-Recipes
 * type BESSJ0 rational approximation the reference uses (libs/extmath.hpp:12-49).
 * Pinned by tests/golden/msa_pp_13000.npy (the reference's own table) in tests/test_eikonal_table.py. */
#define _GNU_SOURCE
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "grb_oracle.h"

#define E_PI 3.141592653589793238462643383279502884197169399375105820974944
static const double E_mpi = 0.13957018; /* PDG::mpi, MPDG.h:56 */
static const double E_GeV2barn = ((6.582119514E-25 * 2.99792458E8) * (6.582119514E-25 * 2.99792458E8)) / 1E-28;
static const double E_GeV2fm = (6.582119514E-25 * 2.99792458E8) * 1E15;

typedef struct { double re, im; } ecx;
static double e_pow2(double x) { return x * x; }
static double e_msqrt(double x) { return sqrt(x > 0.0 ? x : 0.0); }

/* libs/extmath.hpp:12-49 BESSJ0 (polynomial/rational approximation, ~1e-8 accurate: part of the reference) */
static double e_bessj0(double X) {
  const double P1 = 1.0, P2 = -0.1098628627E-2, P3 = 0.2734510407E-4, P4 = -0.2073370639E-5, P5 = 0.2093887211E-6,
               Q1 = -0.1562499995E-1, Q2 = 0.1430488765E-3, Q3 = -0.6911147651E-5, Q4 = 0.7621095161E-6,
               Q5 = -0.9349451520E-7, R1 = 57568490574.0, R2 = -13362590354.0, R3 = 651619640.7, R4 = -11214424.18,
               R5 = 77392.33017, R6 = -184.9052456, S1 = 57568490411.0, S2 = 1029532985.0, S3 = 9494680.718,
               S4 = 59272.64853, S5 = 267.8532712, S6 = 1.0;
  if (X == 0.0) return 1.0;
  const double AX = fabs(X);
  if (AX < 8.0) {
    const double Y = X * X;
    const double FR = R1 + Y * (R2 + Y * (R3 + Y * (R4 + Y * (R5 + Y * R6))));
    const double FS = S1 + Y * (S2 + Y * (S3 + Y * (S4 + Y * (S5 + Y * S6))));
    return FR / FS;
  }
  const double Z = 8.0 / AX, Y = Z * Z, XX = AX - 0.785398164;
  const double FP = P1 + Y * (P2 + Y * (P3 + Y * (P4 + Y * P5)));
  const double FQ = Q1 + Y * (Q2 + Y * (Q3 + Y * (Q4 + Y * Q5)));
  return sqrt(0.636619772 / AX) * (FP * cos(XX) - Z * FQ * sin(XX));
}

/* src/MForm.cc:402-431 S3PomAlpha with the pion-loop insert S3HPL */
static double e_pom_alpha(const grb_param_soft *S, double t) {
  const double BETApi = (2.0 / 3.0) * S->gN_P;
  const double PIC = (e_pow2(BETApi) * e_pow2(E_mpi)) / (32.0 * (E_PI * e_pow2(E_PI)));
  const double tau = 4.0 * e_pow2(E_mpi) / fabs(t);
  const double m = 1.0, sqrtau = e_msqrt(1 + tau);
  const double F_pi = 1.0 / (1.0 - t / S->fc3);
  const double hpl = (4.0 / tau) * e_pow2(F_pi) *
                     (2.0 * tau - pow(1.0 + tau, 3.0 / 2.0) * log((sqrtau + 1.0) / (sqrtau - 1.0)) + log((m * m) / (E_mpi * E_mpi)));
  return 1.0 + S->DELTA_P + S->ALPHA_P * t - PIC * hpl;
}
/* src/MEikonal.cc:129-153 SingleAmpElastic (pomeron, C-even) */
static ecx e_single_amp(const grb_param_soft *S, double s, double t) {
  const double F = (1.0 / (1 - t / S->fc1)) * (1.0 / (1 - t / S->fc2));
  const double a = e_pom_alpha(S, t);
  const double denom = sin(E_PI * a), y = (-1.0 * E_PI) * a;
  const ecx eta = {-(1.0 + 1 * cos(y)) / denom, -(1 * sin(y)) / denom};
  const double f = e_pow2(S->gN_P) * F * F;
  const double p = pow(s / 1.0, a);
  return (ecx){f * eta.re * p, f * eta.im * p};
}
/* MMath.h:378-396 CSIntegral */
static ecx e_simpson(const ecx *f, int N, double h) {
  ecx ev = {0, 0}, od = {0, 0};
  for (int j = 1; j <= N / 2 - 1; ++j) { ev.re += f[2 * j].re; ev.im += f[2 * j].im; }
  for (int j = 1; j <= N / 2; ++j) { od.re += f[2 * j - 1].re; od.im += f[2 * j - 1].im; }
  return (ecx){h * (f[0].re + 2.0 * ev.re + 4.0 * od.re + f[N].re) / 3.0, h * (f[0].im + 2.0 * ev.im + 4.0 * od.im + f[N].im) / 3.0};
}
/* the table is written with 15 significant digits and read back (MEikonal.cc:84-91,404-436) */
static double e_round15(double x) {
  char buf[64];
  snprintf(buf, sizeof buf, "%.15g", x);
  return strtod(buf, NULL);
}
static ecx e_interp(const double *x, const ecx *F, int N, double MIN, double MAX, double a) {
  const double STEP = (MAX - MIN) / N;
  if (a < MIN) a = MIN;
  int i = (int)floor((a - MIN) / STEP);
  if (i < 0) i = 0;
  if (i >= N) i = N - 1;
  const double dx = a - x[i], h = x[i + 1] - x[i];
  return (ecx){F[i].re + dx * ((F[i + 1].re - F[i].re) / h), F[i].im + dx * ((F[i + 1].im - F[i].im) / h)};
}

/* Builds MBT (n_bt+1 rows) and MSA (n_kt2+1 rows); outputs are x, re, im arrays; xs = sigma_tot, el, inel [barn] */
int grb_oracle_build_eikonal(const grb_param_soft *S, double s, int n_bt, int n_kt2, double *bt_x, double *bt_re,
                             double *bt_im, double *kt2_x, double *sa_re, double *sa_im, double *xs) {
  const int NFB = 10000;                                   /* FBIntegralN, MEikonal.h:40 */
  const double MinBT = 1E-6, MaxBT = 10.0 / E_GeV2fm, MinKT2 = 1E-6, MaxKT2 = 25.0, FBmin = 1E-9, FBmax = 30.0;
  if (S->odderon_on) return GRB_ERR_UNSUPPORTED;
  ecx *f = malloc(sizeof(ecx) * (NFB + 1)), *A = malloc(sizeof(ecx) * (NFB + 1)), *F = malloc(sizeof(ecx) * (n_bt + 1));
  /* ---- density: S3Density ---- */
  const double kt_STEP = (FBmax - FBmin) / NFB;
  for (int j = 0; j <= NFB; ++j) {
    const double kt = FBmin + j * kt_STEP;
    A[j] = e_single_amp(S, s, -e_pow2(kt));
  }
  const double bSTEP = (MaxBT - MinBT) / n_bt;
  for (int i = 0; i <= n_bt; ++i) {
    const double bt = MinBT + i * bSTEP;
    for (int j = 0; j <= NFB; ++j) {
      const double kt = FBmin + j * kt_STEP;
      const double b = e_bessj0(bt * kt);
      f[j] = (ecx){A[j].re * b * kt, A[j].im * b * kt};
    }
    const ecx I = e_simpson(f, NFB, kt_STEP);
    const double FACTOR = 1 / (2.0 * E_PI * s);
    bt_x[i] = e_round15(bt);
    F[i] = (ecx){e_round15(I.re * FACTOR), e_round15(I.im * FACTOR)};
    bt_re[i] = F[i].re;
    bt_im[i] = F[i].im;
  }
  /* ---- cross sections: S3CalcXS (single channel) ---- */
  if (xs) {
    const int N = 2 * 3000;
    const double STEP = (MaxBT - MinBT) / N;
    ecx *ft = malloc(sizeof(ecx) * (N + 1)), *fe = malloc(sizeof(ecx) * (N + 1)), *fi = malloc(sizeof(ecx) * (N + 1));
    for (int i = 0; i <= N; ++i) {
      const double bt = MinBT + i * STEP;
      const ecx Om = e_interp(bt_x, F, n_bt, MinBT, MaxBT, bt);
      /* A_el = i (1 - exp(i Omega / 2)) */
      const double er = exp(-Om.im / 2.0), c = cos(Om.re / 2.0), sn = sin(Om.re / 2.0);
      const ecx ex = {er * c, er * sn};
      const ecx Ael = {-(0.0 - ex.im), 1.0 - ex.re};
      const double a2 = e_pow2(hypot(Ael.re, Ael.im));
      ft[i] = (ecx){2.0 * Ael.im * bt, 0};
      fe[i] = (ecx){a2 * bt, 0};
      fi[i] = (ecx){(2.0 * Ael.im - a2) * bt, 0};
    }
    const double IC = 2.0 * E_PI;
    xs[0] = IC * e_simpson(ft, N, STEP).re * E_GeV2barn;
    xs[1] = IC * e_simpson(fe, N, STEP).re * E_GeV2barn;
    xs[2] = IC * e_simpson(fi, N, STEP).re * E_GeV2barn;
    free(ft); free(fe); free(fi);
  }
  /* ---- screening amplitude: S3Screening ---- */
  const double STEPb = (MaxBT - MinBT) / NFB;
  for (int j = 0; j <= NFB; ++j) {
    const double bt = MinBT + j * STEPb;
    const ecx Om = e_interp(bt_x, F, n_bt, MinBT, MaxBT, bt);
    const double er = exp(-Om.im / 2.0), c = cos(Om.re / 2.0), sn = sin(Om.re / 2.0);
    const ecx ex = {er * c, er * sn};
    A[j] = (ecx){-(0.0 - ex.im), 1.0 - ex.re};
  }
  const double kSTEP = (MaxKT2 - MinKT2) / n_kt2;
  for (int i = 0; i <= n_kt2; ++i) {
    const double kt2 = MinKT2 + i * kSTEP, kt = e_msqrt(kt2);
    for (int j = 0; j <= NFB; ++j) {
      const double bt = MinBT + j * STEPb;
      const double b = e_bessj0(bt * kt);
      f[j] = (ecx){A[j].re * b * bt, A[j].im * b * bt};
    }
    const ecx I = e_simpson(f, NFB, STEPb);
    const double FACTOR = 2.0 * E_PI;
    kt2_x[i] = e_round15(kt2);
    sa_re[i] = e_round15((2.0 * s) * FACTOR * I.re);
    sa_im[i] = e_round15((2.0 * s) * FACTOR * I.im);
  }
  free(f); free(A); free(F);
  return GRB_OK;
}
