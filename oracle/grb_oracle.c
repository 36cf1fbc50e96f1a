/* oracle/grb_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, scalar, one-point-at-a-time CPU restatement of GRANIITTI's event-weight path
 * (reference seam A: MFactorized::EventWeight, src/MFactorized.cc:142-184).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this file's
 * library; the product (graniitti_b200/) never does and has no CPU fallback.
 *
 * It follows the reference statement by statement, in the reference's operation order, and is compiled
 * with -ffp-contract=off so that on the same libm it reproduces the reference's rounding.  Each function
 * cites the reference file:line it restates.  It is pinned against the compiled reference itself
 * (oracle/_ref/gra/bin/ref_eval, built from the unmodified sources) through the golden vectors under
 * tests/golden/ (generator: tests/golden/make_golden.py); see tests/test_oracle_golden.py.
 *
 * Parity status: PINNED by reference-generated golden vectors (the reference's own test-suite holds no
 * vector for this path): bit-exact weights and flags on 75 golden sets -- Regge amplitudes, resonances,
 * ladders, the whole tensor-pomeron code incl. the resonance amplitude ME3 (grb_oracle_tres.inc) -- except the HELAS
 * yy -> l+ l- restatement (light-cone components: <= 1e-13).
 */
#include "grb_oracle.h"

#define _GNU_SOURCE
#include <complex.h>
#undef I /* the screening loop names its integral I */
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* Phase marks for the dynamic op count of this restatement (scripts/oracle_opcount.py builds an instrumented copy with
 * -DGRB_OPCOUNT and counts executed fp64 operations per phase; SURVEY.md section 8d convention).  No-ops in the normal build. */
enum { OPC_KIN = 0, OPC_SCAL_KIN, OPC_BORN, OPC_NODE_INV, OPC_NODE_KIN, OPC_NODE_SCAL, OPC_NODE_AMP, OPC_NODE_ACC, OPC_POST, OPC_NPHASE };
#ifdef GRB_OPCOUNT
int grb_opc_mark(int phase); /* returns the previous phase */
#define OPC(k) grb_opc_mark(k)
#else
#define OPC(k) (0)
#endif

/* Conditioning probe (tests/fullsize.py): the result of the k-th libm call of a point can be moved by one ulp.  Two
 * IEEE-754 platforms differ in exactly these results (glibc vs CUDA libm: <= 1-2 ulp; + - * / sqrt are correctly rounded
 * on both), so max_k |W(k-th result +- 1 ulp) - W| / |W| is the point's conditioning with respect to everything a
 * faithful implementation on another platform may legitimately compute differently.  Disabled (g_pert_call < 0) the
 * wrappers return their argument unchanged: the restatement's results do not depend on them. */
static int g_pert_call = -1, g_pert_sign = 1, g_call_idx = 0;
/* k >= 0: the k-th call of a point; k == -2: every call (a coherent shift); sign = +-n moves the result by n ulps (CUDA's
 * libm documents up to 2 ulp for sin / cos / sinh / tan / atan2, glibc up to 1) */
static int g_pert_start = 0, g_pert_stride = 0, g_pert_off = 0, g_pert_pz = 0;
static int g_pert_pz2 = 0;
void grb_oracle_perturb_pz(int n_ulps) { g_pert_pz = n_ulps; g_pert_pz2 = 0; }
void grb_oracle_perturb_pz2(int n_ulps) { g_pert_pz2 = n_ulps; g_pert_pz = 0; }
void grb_oracle_perturb_call(int k, int sign) { g_pert_call = k; g_pert_sign = sign; g_pert_stride = 0; }
/* k == -3: every call with index >= start and (index - start) % stride == off: the same call site in every node of the
 * screening loop (the loop re-derives y_X, sinh y_X, cosh y_X ... per node; an implementation that hoists them differs
 * coherently in all nodes) */
void grb_oracle_perturb_stride(int start, int stride, int off, int sign) {
  g_pert_call = -3; g_pert_sign = sign; g_pert_start = start; g_pert_stride = stride; g_pert_off = off;
}
int  grb_oracle_calls_of_last_point(void) { return g_call_idx; }
static inline double opert(double v) {
  const int idx = g_call_idx++;
  if (idx == g_pert_call || g_pert_call == -2 ||
      (g_pert_call == -3 && g_pert_stride > 0 && idx >= g_pert_start && (idx - g_pert_start) % g_pert_stride == g_pert_off)) {
    for (int n = abs(g_pert_sign); n > 0; --n) v = nextafter(v, g_pert_sign > 0 ? INFINITY : -INFINITY);
  }
  return v;
}
#ifndef GRB_OPCOUNT /* (the op-counting build lets the compiler merge identical libm calls) */
#define exp(x) opert(exp(x))
#define log(x) opert(log(x))
#define sin(x) opert(sin(x))
#define cos(x) opert(cos(x))
#define tan(x) opert(tan(x))
#define sinh(x) opert(sinh(x))
#define cosh(x) opert(cosh(x))
#define atan2(y, x) opert(atan2(y, x))
#define pow(x, y) opert(pow(x, y))
#endif

/* reference include/Graniitti/MMath.h:30 (long double literal rounded to double) */
#define O_PI 3.141592653589793238462643383279502884197169399375105820974944
/* reference include/Graniitti/MPDG.h:64-78 */
static const double O_GeV2barn = ((6.582119514E-25 * 2.99792458E8) * (6.582119514E-25 * 2.99792458E8)) / 1E-28;
/* reference include/Graniitti/MProcess.h:372 */
#define O_ZERO_EPS 1e-12

typedef struct {
  double x, y, z, e;
} v4;
typedef struct {
  double re, im;
} cx;

/* ---- M4Vec algebra, reference include/Graniitti/M4Vec.h:120-160,222-236 ----------------------------- */
static inline double msqrt(double x) { return sqrt(x > 0.0 ? x : 0.0); } /* MMath.h:250 */
static inline double pow2(double x) { return x * x; }
static inline v4     v4_add(v4 a, v4 b) { return (v4){a.x + b.x, a.y + b.y, a.z + b.z, a.e + b.e}; }
static inline v4     v4_sub(v4 a, v4 b) { return (v4){a.x - b.x, a.y - b.y, a.z - b.z, a.e - b.e}; }
static inline double v4_p3mod2(v4 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
static inline double v4_m2(v4 a) { return a.e * a.e - v4_p3mod2(a); }
static inline double v4_m(v4 a) {
  const double m2 = v4_m2(a);
  return (m2 > 0.0) ? msqrt(m2) : -msqrt(-m2);
}
static inline double v4_pt2(v4 a) { return a.x * a.x + a.y * a.y; }
static inline double v4_pt(v4 a) { return msqrt(v4_pt2(a)); }
static inline double v4_p3mod(v4 a) { return msqrt(v4_p3mod2(a)); }
static inline double v4_et(v4 a) { return msqrt(v4_m2(a) + v4_pt2(a)); }
static inline double v4_eta(v4 a) { return 0.5 * log((v4_p3mod(a) + a.z) / (v4_p3mod(a) - a.z)); }
static inline double v4_rap(v4 a) { return 0.5 * log((a.e + a.z) / (a.e - a.z)); }
static inline double v4_dot(v4 a, v4 b) { return a.e * b.e - (a.x * b.x + a.y * b.y + a.z * b.z); }

/* ---- std::complex<double> algebra as libstdc++ evaluates it ------------------------------------------ */
static inline cx cx_mul(cx a, cx b) { return (cx){a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
static inline cx cx_scale(cx a, double d) { return (cx){a.re * d, a.im * d}; }
static inline cx cx_add(cx a, cx b) { return (cx){a.re + b.re, a.im + b.im}; }
static inline double cx_abs2(cx a) { return pow2(hypot(a.re, a.im)); } /* MMath.h:253 abs2 = pow2(std::abs) */

/* ---- per-engine state = the pieces of gra::LORENTZSCALAR the path touches (MKinematics.h:1437-1551) -- */
typedef struct {
  const grb_process_desc *d;
  const grb_oracle_eikonal *eik;
  v4     pbeam1, pbeam2, beamsum;
  v4     pfinal[10], pfinal_orig[10];
  v4     node_p4[GRB_MAX_NODES];
  double node_moff[GRB_MAX_NODES];
  double ss[10][10];
  double s1, s2, t1, t2, t_hat, u_hat, m2, Y, Pt, x1, x2, qt1, qt2;
  v4     q1, q2;
  double M_MIN, M_MAX, DW;
  v4     pkt[7]; /* <C>: intermediate 'difference momenta' q_i (MContinuum::pkt_), Kf - 1 of them */
  double tt_1[10], tt_2[10], tt_xy[10][10]; /* t-type scalars of the ladders (MProcess.cc:1558-1564) */
  int    excite1, excite2; /* lts.excite1/2: which forward leg is an excited N* system (MProcess.cc:1317-1365) */
  double log_M2_f_min, log_M2_f_max;
  cx     hamp[GRB_MAX_HAMP];
  int    nhamp;
  /* tape of internal uniforms (reference: thread-local MRandom, src/MRandom.cc:62-65) */
  const double *tape;
  int           tape_len, tape_pos, tape_overrun;
  /* loop-invariant signature factors ReggeEta(a0,sgn), MForm.cc:382-385 */
  cx regge_eta0[3];
} ost;

/* test switch: evaluate ill-conditioned amplitude algebra in extended precision (noise measurement) */
static int g_oracle_extended = 0;
void grb_oracle_set_extended(int on) { g_oracle_extended = on; }

static double tape_U(ost *S, double a, double b) {
  double v = 0.5;
  if (S->tape_pos < S->tape_len)
    v = S->tape[S->tape_pos];
  else
    S->tape_overrun = 1;
  S->tape_pos++;
  return a + (b - a) * v;
}

/* reference src/MRandom.cc:82-107 RelativisticBWRandom */
static double bw_random(ost *S, double m0, double Gamma, double LIMIT, double M_MIN) {
  if (Gamma < 1e-40) return m0;
  const double m2    = pow2(m0);
  const double m2max = pow2(m0 + LIMIT * Gamma);
  const double m2min = fmax(fmax(0.0, pow2(M_MIN)), pow2(m0 - LIMIT * Gamma));
  double       m2val = 0.0;
  for (int it = 0; it < 100000; ++it) {
    const double R = tape_U(S, 0, 1);
    m2val = m2 + m0 * Gamma *
                     tan(atan2(m2min - m2, m0 * Gamma) +
                         R * (atan2(m2max - m2, m0 * Gamma) - atan2(m2min - m2, m0 * Gamma)));
    if (m2val >= 0) break;
  }
  return msqrt(m2val);
}

/* reference src/MProcess.cc:753-812 GetOffShellMass (FLATMASS2 == false branch) */
static double offshell_mass(ost *S, int node) {
  const grb_process_desc *d = S->d;
  const grb_node *        b = &d->nodes[node];
  if (b->width < 1e-40) return b->mass;
  double daughter_masses = 0;
  if (b->n_children != 0) {
    for (int i = 0; i < b->n_children; ++i) {
      const grb_node *l = &d->nodes[b->first_child + i];
      daughter_masses += fmax(0.0, bw_random(S, l->mass, l->width, d->offshell, 0.0));
    }
    daughter_masses += 1e-5;
  }
  /* inner loop returns at its first trial: max(dm, .) >= dm */
  return fmax(daughter_masses, bw_random(S, b->mass, b->width, d->offshell, daughter_masses));
}

/* reference include/Graniitti/MKinematics.h:33-64 SolvePz3_A (same symbol names) */
static double solve_pz3_A(double m3, double pt3, double pt4, double pz5, double E5, double s) {
  const double t2  = E5 * E5;
  const double t3  = pt3 * pt3;
  const double t4  = pt4 * pt4;
  const double t5  = pz5 * pz5;
  const double t6  = m3 * m3;
  const double t7  = sqrt(s);
  const double t8  = s * t2 * 6.0;
  const double t9  = pow(s, 3.0 / 2.0);
  const double t10 = t2 * t2;
  const double t11 = t3 * t3;
  const double t12 = t4 * t4;
  const double t13 = t5 * t5;
  const double t14 = s * s;
  const double t15 = t5 * t6 * 4.0;
  const double t16 = t3 * t5 * 2.0;
  const double t17 = t4 * t5 * 2.0;
  const double t18 = E5 * t6 * t7 * 8.0;
  const double t19 = E5 * t3 * t7 * 4.0;
  const double t20 = E5 * t4 * t7 * 4.0;
  const double t21 = E5 * t5 * t7 * 4.0;
  const double t22 = t8 + t10 + t11 + t12 + t13 + t14 + t15 + t16 + t17 + t18 + t19 + t20 + t21 -
                     E5 * t9 * 4.0 - s * t3 * 2.0 - s * t4 * 2.0 - s * t5 * 2.0 - s * t6 * 4.0 -
                     t2 * t3 * 2.0 - t2 * t4 * 2.0 - t2 * t5 * 2.0 - t3 * t4 * 2.0 - t2 * t6 * 4.0 -
                     E5 * t2 * t7 * 4.0;
  const double t23 = sqrt(t22);
  return pz5 * (-1.0 / 2.0) -
         (E5 * t23 * (1.0 / 2.0) - t7 * t23 * (1.0 / 2.0) + pz5 * (t3 * (1.0 / 2.0) - t4 * (1.0 / 2.0))) /
             (s + t2 - t5 - E5 * t7 * 2.0);
}

/* reference include/Graniitti/MKinematics.h:117-160 SolvePz3_B */
static double solve_pz3_B(double m3, double m4, double pt3, double pt4, double pz5, double E5, double s) {
  const double t2  = E5 * E5;
  const double t3  = m3 * m3;
  const double t4  = m4 * m4;
  const double t5  = pt3 * pt3;
  const double t6  = pt4 * pt4;
  const double t7  = pz5 * pz5;
  const double t8  = sqrt(s);
  const double t9  = s * t2 * 6.0;
  const double t10 = pow(s, 3.0 / 2.0);
  const double t11 = t2 * t2;
  const double t12 = t3 * t3;
  const double t13 = t4 * t4;
  const double t14 = t5 * t5;
  const double t15 = t6 * t6;
  const double t16 = t7 * t7;
  const double t17 = s * s;
  const double t18 = t3 * t5 * 2.0;
  const double t19 = t4 * t6 * 2.0;
  const double t20 = t3 * t7 * 2.0;
  const double t21 = t4 * t7 * 2.0;
  const double t22 = t5 * t7 * 2.0;
  const double t23 = t6 * t7 * 2.0;
  const double t24 = E5 * t3 * t8 * 4.0;
  const double t25 = E5 * t4 * t8 * 4.0;
  const double t26 = E5 * t5 * t8 * 4.0;
  const double t27 = E5 * t6 * t8 * 4.0;
  const double t28 = E5 * t7 * t8 * 4.0;
  const double t29 = t9 + t11 + t12 + t13 + t14 + t15 + t16 + t17 + t18 + t19 + t20 + t21 + t22 + t23 +
                     t24 + t25 + t26 + t27 + t28 - E5 * t10 * 4.0 - s * t3 * 2.0 - s * t4 * 2.0 -
                     s * t5 * 2.0 - s * t6 * 2.0 - s * t7 * 2.0 - t2 * t3 * 2.0 - t2 * t4 * 2.0 -
                     t2 * t5 * 2.0 - t3 * t4 * 2.0 - t2 * t6 * 2.0 - t2 * t7 * 2.0 - t3 * t6 * 2.0 -
                     t4 * t5 * 2.0 - t5 * t6 * 2.0 - E5 * t2 * t8 * 4.0;
  const double t30 = sqrt(t29);
  return pz5 * (-1.0 / 2.0) -
         (E5 * t30 * (1.0 / 2.0) - t8 * t30 * (1.0 / 2.0) +
          pz5 * (t3 * (1.0 / 2.0) - t4 * (1.0 / 2.0) + t5 * (1.0 / 2.0) - t6 * (1.0 / 2.0))) /
             (s + t2 - t7 - E5 * t8 * 2.0);
}

/* reference MKinematics.h:220-229 SolvePz */
static double solve_pz(double m3, double m4, double pt3, double pt4, double pz5, double E5, double s) {
  if (fabs(m3 - m4) < 1e-9) return solve_pz3_A(m3, pt3, pt4, pz5, E5, s);
  return solve_pz3_B(m3, m4, pt3, pt4, pz5, E5, s);
}

/* reference MMath.h:191-198 CheckEMC (note: energy difference is not abs()) */
static int check_emc(v4 diff) {
  const double eps = 1e-6;
  if (fabs(diff.x) > eps || fabs(diff.y) > eps || fabs(diff.z) > eps || diff.e > eps) return 0;
  return 1;
}

/* reference MKinematics.h:383-401 LorentzBoost; (beta*p3).sum() of libstdc++ sums from the last element */
static v4 lorentz_boost(v4 boost, double M0, v4 p, int sign) {
  const double b0 = sign * boost.x / boost.e, b1 = sign * boost.y / boost.e, b2 = sign * boost.z / boost.e;
  const double gamma  = boost.e / M0;
  double       kappa1 = b2 * p.z;
  kappa1 += b1 * p.y;
  kappa1 += b0 * p.x;
  const double kappa2 = gamma * (gamma * kappa1 / (1.0 + gamma) + p.e);
  return (v4){p.x + kappa2 * b0, p.y + kappa2 * b1, p.z + kappa2 * b2, gamma * (p.e + kappa1)};
}

/* reference MKinematics.h:294-298 DecayMomentum, :267-271 dPhi2 */
static double decay_momentum(double x, double y, double z) {
  return msqrt((x - y - z) * (x + y + z) * (x + y - z) * (x - y + z)) / (2.0 * x);
}
static double dphi2(double E, double p_norm) {
  const double V = 2.0 * 2.0 * O_PI;
  return V * p_norm / (16.0 * pow2(O_PI) * E);
}

/* reference MKinematics.h:403-459 FlatIsotropic + Isotropic + TwoBodyPhaseSpace */
static double two_body(ost *S, v4 mother, double M0, double m1, double m2, v4 *p1, v4 *p2) {
  const double pnorm    = decay_momentum(M0, m1, m2);
  const double costheta = tape_U(S, -1.0, 1.0);
  const double sintheta = msqrt(1.0 - pow2(costheta));
  const double phi      = tape_U(S, 0.0, 2.0 * O_PI);
  const double k0 = pnorm * sintheta * cos(phi), k1 = pnorm * sintheta * sin(phi), k2 = pnorm * costheta;
  const double e0 = msqrt(pow2(m1) + pow2(pnorm)), e1 = msqrt(pow2(m2) + pow2(pnorm));
  *p1 = lorentz_boost(mother, M0, (v4){k0, k1, k2, e0}, 1);
  *p2 = lorentz_boost(mother, M0, (v4){-k0, -k1, -k2, e1}, 1);
  return dphi2(M0, pnorm);
}

/* reference src/MProcess.cc:1231-1284 ConstructDecayKinematics (2-body branches) */
static int construct_decay(ost *S, int node) {
  const grb_node *b = &S->d->nodes[node];
  if (b->n_children == 0) return 1;
  if (b->n_children != 2) return 0;
  const int c0 = b->first_child, c1 = b->first_child + 1;
  double    m0v, m1v;
  for (int it = 0;; ++it) {
    m0v = offshell_mass(S, c0);
    m1v = offshell_mass(S, c1);
    S->node_moff[c0] = m0v;
    S->node_moff[c1] = m1v;
    /* std::accumulate(m.rbegin(), m.rend(), 0.0): 0 + m[last] + ... + m[0] */
    if (v4_m(S->node_p4[node]) < (0.0 + m1v) + m0v) {
      if (it > 100000 || S->tape_overrun) return 0;
      continue;
    }
    break;
  }
  const double w = two_body(S, S->node_p4[node], v4_m(S->node_p4[node]), m0v, m1v, &S->node_p4[c0], &S->node_p4[c1]);
  if (w < 0) return 0;
  if (!construct_decay(S, c0)) return 0;
  if (!construct_decay(S, c1)) return 0;
  return 1;
}

/* reference src/MProcess.cc:1499-1584 GetLorentzScalars */
static int lorentz_scalars_impl(ost *S, int Nf);
static int lorentz_scalars(ost *S, int Nf) {
#ifdef GRB_OPCOUNT
  const int prev = OPC(OPC_SCAL_KIN);
  if (prev == OPC_NODE_KIN) OPC(OPC_NODE_SCAL);
  const int ok = lorentz_scalars_impl(S, Nf);
  OPC(prev);
  return ok;
#else
  return lorentz_scalars_impl(S, Nf);
#endif
}
static int lorentz_scalars_impl(ost *S, int Nf) {
  for (int i = 0; i <= Nf; ++i)
    for (int j = 0; j <= Nf; ++j) {
      if (i < j) S->ss[i][j] = v4_m2(v4_add(S->pfinal[i], S->pfinal[j]));
      if (S->ss[i][j] < 0) return 0;
    }
  for (int i = 0; i <= Nf; ++i)
    for (int j = i; j <= Nf; ++j)
      if (i != j) S->ss[j][i] = S->ss[i][j];
  S->s1 = S->ss[1][0];
  S->s2 = S->ss[2][0];
  if (S->s1 < 0 || S->s2 < 0) return 0;
  S->q1 = v4_sub(S->pbeam1, S->pfinal[1]);
  S->q2 = v4_sub(S->pbeam2, S->pfinal[2]);
  S->t1 = v4_m2(S->q1);
  S->t2 = v4_m2(S->q2);
  if (S->t1 > 0 || S->t2 > 0) return 0;
  if (S->d->n_legs == 2) {
    S->t_hat = v4_m2(v4_sub(S->q1, S->node_p4[0]));
    S->u_hat = v4_m2(v4_sub(S->q1, S->node_p4[1]));
  }
  for (int i = 3; i <= Nf; ++i) S->tt_1[i] = v4_m2(v4_sub(S->q1, S->pfinal[i]));
  for (int i = 3; i <= Nf; ++i)
    for (int j = 3; j <= Nf; ++j) S->tt_xy[i][j] = v4_m2(v4_sub(v4_sub(S->q1, S->pfinal[i]), S->pfinal[j]));
  for (int i = 3; i <= Nf; ++i) S->tt_2[i] = v4_m2(v4_sub(S->q2, S->pfinal[i]));
  S->x1  = (1 - S->pfinal[1].z / S->pbeam1.z);
  S->x2  = (1 - S->pfinal[2].z / S->pbeam2.z);
  S->qt1 = v4_pt(S->q1);
  S->qt2 = v4_pt(S->q2);
  S->m2  = v4_m2(S->pfinal[0]);
  S->Y   = v4_rap(S->pfinal[0]);
  S->Pt  = v4_pt(S->pfinal[0]);
  return 1;
}

/* reference src/MProcess.cc:1317-1365 SampleForwardMasses: log-flat M^2 in [1.09^2, XI_max s]; single excitation picks
 * the side with one internal uniform */
static void sample_forward_masses(ost *S, const double *rvec, double *m1, double *m2) {
  const grb_process_desc *d = S->d;
  *m1 = d->beam1_mass;
  *m2 = d->beam2_mass;
  S->excite1 = 0;
  S->excite2 = 0;
  if (d->excitation != 0) {
    const double M2_f_min = pow2(1.09), M2_f_max = d->gcuts.XI_max * d->s;
    S->log_M2_f_min = log(M2_f_min);
    S->log_M2_f_max = log(M2_f_max);
    if (d->excitation == 1) {
      const double u = S->log_M2_f_min + (S->log_M2_f_max - S->log_M2_f_min) * rvec[0];
      const double r = exp(u);
      if (tape_U(S, 0, 1) < 0.5) {
        *m1 = msqrt(r);
        S->excite1 = 1;
      } else {
        *m2 = msqrt(r);
        S->excite2 = 1;
      }
    } else {
      const double u1 = S->log_M2_f_min + (S->log_M2_f_max - S->log_M2_f_min) * rvec[0];
      const double r1 = exp(u1);
      const double u2 = S->log_M2_f_min + (S->log_M2_f_max - S->log_M2_f_min) * rvec[1];
      const double r2 = exp(u2);
      *m1 = msqrt(r1);
      *m2 = msqrt(r2);
      S->excite1 = 1;
      S->excite2 = 1;
    }
  }
}

/* reference src/MFactorized.cc:307-412 B51BuildKin (symmetric beams: the lab boost of :333-340 is the
 * identity; asymmetric beams are rejected at creation) */
static int b51_build_kin(ost *S, double pt1, double pt2, double phi1, double phi2, double yX, double m2X,
                         double m1, double m2) {
  const grb_process_desc *d = S->d;
  v4 p1 = {pt1 * cos(phi1), pt1 * sin(phi1), 0, 0};
  v4 p2 = {pt2 * cos(phi2), pt2 * sin(phi2), 0, 0};
  v4 pX = {-(p1.x + p2.x), -(p1.y + p2.y), 0, 0};
  const double mtX = msqrt(m2X + v4_pt2(pX));
  pX.z = mtX * sinh(yX);
  pX.e = mtX * cosh(yX);
  if (pX.e > (d->sqrt_s - (m1 + m2))) return 0;
  double p1z = solve_pz(m1, m2, pt1, pt2, pX.z, pX.e, d->s);
  /* conditioning probe: the solved longitudinal momentum moved by n ulps.  A 1-ulp libm difference upstream changes the
   * exact solution by far less than one ulp of p1z ~ sqrt(s)/2, but flips its ROUNDING with probability ~1e-3; with an
   * excited forward system that flip moves M^2 = E^2 - p^2 by 2 p ulp(p) = 1.2e-8 GeV^2, which every node of the screening
   * loop inherits through m1 = pfinal_orig[1].M() (MFactorized.cc:97-98) into S3FINEL(t, M^2). */
  for (int n = abs(g_pert_pz); n > 0; --n) p1z = nextafter(p1z, g_pert_pz > 0 ? INFINITY : -INFINITY);
  double p2z = -(pX.z + p1z);
  for (int n = abs(g_pert_pz2); n > 0; --n) p2z = nextafter(p2z, g_pert_pz2 > 0 ? INFINITY : -INFINITY);
  if (p1z < 0 || p2z > 0) return 0;
  if (isnan(p1z)) return 0; /* NaN compares false above and would be caught by CheckEMC's comparisons? no:
                               NaN passes all '>' tests; the reference then fails GetLorentzScalars
                               only by luck -- see note in DESIGN.md; kept explicit here */
  p1.z = p1z;
  p1.e = msqrt(pow2(m1) + pow2(pt1) + pow2(p1z));
  p2.z = p2z;
  p2.e = msqrt(pow2(m2) + pow2(pt2) + pow2(p2z));
  S->pfinal[1] = p1;
  S->pfinal[2] = p2;
  S->pfinal[0] = pX;
  if (!check_emc(v4_sub(S->beamsum, v4_add(v4_add(S->pfinal[1], S->pfinal[2]), S->pfinal[0])))) return 0;

  if (d->n_legs == 2) {
    S->DW = two_body(S, S->pfinal[0], msqrt(m2X), S->node_moff[0], S->node_moff[1], &S->node_p4[0], &S->node_p4[1]);
  } else {
    return 0; /* 3-/N-body central systems are outside the device path */
  }
  if (S->DW < 0) return 0;
  for (int i = 0; i < d->n_legs; ++i) S->pfinal[i + 3] = S->node_p4[i];
  for (int i = 0; i < d->n_legs; ++i)
    if (!construct_decay(S, i)) return 0;
  return lorentz_scalars(S, d->n_legs + 2);
}

/* reference src/MFactorized.cc:245-304 B51RandomKin (EXCITATION == 0) */
static int b51_random_kin(ost *S, const double *r) {
  const grb_process_desc *d  = S->d;
  const grb_gencut *      gc = &d->gcuts;
  const double u1 = log(gc->forward_pt_min + O_ZERO_EPS) +
                    (log(gc->forward_pt_max) - log(gc->forward_pt_min + O_ZERO_EPS)) * r[0];
  const double u2 = log(gc->forward_pt_min + O_ZERO_EPS) +
                    (log(gc->forward_pt_max) - log(gc->forward_pt_min + O_ZERO_EPS)) * r[1];
  const double pt1  = exp(u1);
  const double pt2  = exp(u2);
  const double phi1 = 2.0 * O_PI * r[2];
  const double phi2 = 2.0 * O_PI * r[3];
  const double yX   = gc->Y_min + (gc->Y_max - gc->Y_min) * r[4];
  for (int trials = 0;; ++trials) {
    double M_sum = 0.0;
    for (int i = 0; i < d->n_legs; ++i) {
      S->node_moff[i] = offshell_mass(S, i);
      M_sum += S->node_moff[i];
    }
    S->M_MIN = M_sum + 1e-4;
    S->M_MAX = d->sqrt_s - (v4_m(S->pbeam1) + v4_m(S->pbeam2));
    S->M_MIN = fmax(S->M_MIN, gc->M_min);
    S->M_MAX = fmin(S->M_MAX, gc->M_max);
    if (S->M_MIN < S->M_MAX) break;
    if (trials > 100000 || S->tape_overrun) return 0;
  }
  const double m2X = pow2(S->M_MIN) + (pow2(S->M_MAX) - pow2(S->M_MIN)) * r[5];
  double m1, m2;
  sample_forward_masses(S, r + 6, &m1, &m2);
  return b51_build_kin(S, pt1, pt2, phi1, phi2, yX, m2X, m1, m2);
}

/* reference src/MFactorized.cc:84-139 LoopKinematics */
static int loop_kinematics(ost *S, double p1px, double p1py, double p2px, double p2py) {
  const grb_process_desc *d = S->d;
  S->pfinal[1].x = p1px;
  S->pfinal[1].y = p1py;
  S->pfinal[2].x = p2px;
  S->pfinal[2].y = p2py;
  const double pt1 = v4_pt(S->pfinal[1]);
  const double pt2 = v4_pt(S->pfinal[2]);
  const double m1  = v4_m(S->pfinal_orig[1]);
  const double m2  = v4_m(S->pfinal_orig[2]);
  const double mX  = v4_m(S->pfinal_orig[0]);
  const double yX  = v4_rap(S->pfinal_orig[0]);
  S->pfinal[0].x = -(S->pfinal[1].x + S->pfinal[2].x);
  S->pfinal[0].y = -(S->pfinal[1].y + S->pfinal[2].y);
  const double mtX = msqrt(pow2(mX) + v4_pt2(S->pfinal[0]));
  S->pfinal[0].z = mtX * sinh(yX);
  S->pfinal[0].e = mtX * cosh(yX);
  if (S->pfinal[0].e > (d->sqrt_s - (m1 + m2))) return 0;
  const double p1z = solve_pz(m1, m2, pt1, pt2, S->pfinal[0].z, S->pfinal[0].e, d->s);
  const double p2z = -(S->pfinal[0].z + p1z);
  if (isnan(p1z)) return 0;
  if (p1z < 0 || p2z > 0) return 0;
  S->pfinal[1].z = p1z;
  S->pfinal[1].e = msqrt(pow2(m1) + pow2(pt1) + pow2(p1z));
  S->pfinal[2].z = p2z;
  S->pfinal[2].e = msqrt(pow2(m2) + pow2(pt2) + pow2(p2z));
  if (!check_emc(v4_sub(S->beamsum, v4_add(v4_add(S->pfinal[1], S->pfinal[2]), S->pfinal[0])))) return 0;
  return lorentz_scalars(S, d->n_legs + 2);
}

/* ---- <C> phase-space class: MContinuum, 2 -> N direct (restated for N = 4, two central legs) ---------------------- */

/* reference src/MContinuum.cc:438-538 BLinearSystem: p = A b with b0 = q0 - (p1+p2), b_i = -q_{i-1} - (p1+p2); A is the
 * inverse of the ladder's linear system (:480-516, tabulated there for Kf = 2..8; Kf = 2, 4 and 6 restated).  The loop adds
 * b[j] * A[i][j] for every j (zero entries add signed zeros). */
static const double o_blin2[2][2] = {{1.0 / 2.0, 0.0}, {0.0, 1.0 / 2.0}};
static const double o_blin4[4][4] = {{7.0 / 8.0, 1.0 / 8.0, -1.0 / 2.0, -1.0 / 4.0},
                                     {3.0 / 8.0, 5.0 / 8.0, -1.0 / 2.0, -1.0 / 4.0},
                                     {-1.0 / 8.0, 1.0 / 8.0, 1.0 / 2.0, -1.0 / 4.0},
                                     {-5.0 / 8.0, -3.0 / 8.0, 1.0 / 2.0, 3.0 / 4.0}};
static const double o_blin6[6][6] = {{4.0 / 3.0, 1.0 / 2.0, -2.0 / 3.0, -1.0 / 2.0, -1.0 / 3.0, -1.0 / 6.0},
                                     {5.0 / 6.0, 1.0 / 1.0, -2.0 / 3.0, -1.0 / 2.0, -1.0 / 3.0, -1.0 / 6.0},
                                     {1.0 / 3.0, 1.0 / 2.0, 1.0 / 3.0, -1.0 / 2.0, -1.0 / 3.0, -1.0 / 6.0},
                                     {-1.0 / 6.0, 0.0, 1.0 / 3.0, 1.0 / 2.0, -1.0 / 3.0, -1.0 / 6.0},
                                     {-2.0 / 3.0, -1.0 / 2.0, 1.0 / 3.0, 1.0 / 2.0, 2.0 / 3.0, -1.0 / 6.0},
                                     {-7.0 / 6.0, -1.0 / 1.0, 1.0 / 3.0, 1.0 / 2.0, 2.0 / 3.0, 5.0 / 6.0}};
static void blinear(int Kf, const v4 *q, v4 p1f, v4 p2f, v4 *p) {
  const v4 w = v4_add(p1f, p2f);
  v4 b[8];
  for (int i = 0; i < Kf; ++i)
    b[i] = (i == 0) ? v4_sub(q[0], w) : v4_sub((v4){q[i - 1].x * (-1.0), q[i - 1].y * (-1.0), q[i - 1].z * (-1.0), q[i - 1].e * (-1.0)}, w);
  for (int i = 0; i < Kf; ++i) {
    v4 acc = {0, 0, 0, 0};
    for (int j = 0; j < Kf; ++j) {
      const double a = (Kf == 2) ? o_blin2[i][j] : (Kf == 4 ? o_blin4[i][j] : o_blin6[i][j]);
      acc = v4_add(acc, (v4){b[j].x * a, b[j].y * a, b[j].z * a, b[j].e * a});
    }
    p[i] = acc;
  }
}

/* reference src/MContinuum.cc:334-435 BNBuildKin (symmetric beams, no forward excitation, veto cuts inactive) */
static int bn_build_kin(ost *S, double pt1, double pt2, double phi1, double phi2, const double *kt, const double *phik,
                        const double *y, double m1, double m2) {
  const grb_process_desc *d = S->d;
  const int Kf = d->n_legs;
  v4 p1 = {pt1 * cos(phi1), pt1 * sin(phi1), 0, 0};
  v4 p2 = {pt2 * cos(phi2), pt2 * sin(phi2), 0, 0};
  for (int i = 0; i < Kf - 1; ++i) S->pkt[i] = (v4){kt[i] * cos(phik[i]), kt[i] * sin(phik[i]), 0, 0};
  v4 p[8];
  blinear(Kf, S->pkt, p1, p2, p);
  v4 sumP = {0, 0, 0, 0};
  for (int i = 0; i < Kf; ++i) {
    const double m  = S->node_moff[i];
    const double mt = msqrt(pow2(m) + v4_pt2(p[i]));
    p[i].z = mt * sinh(y[i]);
    p[i].e = mt * cosh(y[i]);
    sumP = v4_add(sumP, p[i]);
  }
  if (sumP.e > d->sqrt_s) return 0;
  double p1z = solve_pz(m1, m2, pt1, pt2, sumP.z, sumP.e, d->s);
  for (int n = abs(g_pert_pz); n > 0; --n) p1z = nextafter(p1z, g_pert_pz > 0 ? INFINITY : -INFINITY); /* conditioning probe, see b51_build_kin */
  double p2z = -(sumP.z + p1z);
  for (int n = abs(g_pert_pz2); n > 0; --n) p2z = nextafter(p2z, g_pert_pz2 > 0 ? INFINITY : -INFINITY);
  if (p1z < 0 || p2z > 0) return 0;
  if (isnan(p1z)) return 0; /* see b51_build_kin */
  p1.z = p1z;
  p1.e = msqrt(pow2(m1) + pow2(pt1) + pow2(p1z));
  p2.z = p2z;
  p2.e = msqrt(pow2(m2) + pow2(pt2) + pow2(p2z));
  S->pfinal[1] = p1;
  S->pfinal[2] = p2;
  S->pfinal[0] = sumP;
  double sumM = 0;
  for (int i = 0; i < Kf; ++i) {
    S->node_p4[i]    = p[i];
    S->pfinal[i + 3] = p[i];
    sumM += v4_m(p[i]);
  }
  if (v4_m(sumP) < sumM) return 0;
  if (!check_emc(v4_sub(S->beamsum, v4_add(v4_add(S->pfinal[1], S->pfinal[2]), S->pfinal[0])))) return 0;
  for (int i = 0; i < d->n_legs; ++i)
    if (!construct_decay(S, i)) return 0;
  return lorentz_scalars(S, d->n_legs + 2);
}

/* reference src/MContinuum.cc:276-331 BNRandomKin: r = [pt1, pt2, phi1, phi2, kt_1..kt_{Kf-1}, phi_1..phi_{Kf-1}, y_1..y_Kf] */
static int bn_random_kin(ost *S, const double *r) {
  const grb_process_desc *d  = S->d;
  const grb_gencut *      gc = &d->gcuts;
  const int               Kf = d->n_legs;
  const double u1 = log(gc->forward_pt_min + O_ZERO_EPS) +
                    (log(gc->forward_pt_max) - log(gc->forward_pt_min + O_ZERO_EPS)) * r[0];
  const double u2 = log(gc->forward_pt_min + O_ZERO_EPS) +
                    (log(gc->forward_pt_max) - log(gc->forward_pt_min + O_ZERO_EPS)) * r[1];
  const double pt1 = exp(u1), pt2 = exp(u2);
  const double phi1 = 2.0 * O_PI * r[2], phi2 = 2.0 * O_PI * r[3];
  for (int i = 0; i < d->n_legs; ++i) S->node_moff[i] = offshell_mass(S, i);
  double kt[7], phik[7], y[8];
  int    ind = 4;
  for (int i = 0; i < Kf - 1; ++i) kt[i] = gc->kt_min + (gc->kt_max - gc->kt_min) * r[ind++];
  for (int i = 0; i < Kf - 1; ++i) phik[i] = 2.0 * O_PI * r[ind++];
  for (int i = 0; i < Kf; ++i) y[i] = gc->rap_min + (gc->rap_max - gc->rap_min) * r[ind++];
  double m1, m2; /* forward N* system masses from the coordinates after the rapidities (MContinuum.cc:322-328) */
  sample_forward_masses(S, r + ind, &m1, &m2);
  return bn_build_kin(S, pt1, pt2, phi1, phi2, kt, phik, y, m1, m2);
}

/* reference src/MContinuum.cc:94-161 LoopKinematics: the central legs ARE updated here (lts.pfinal[3..]) -- but not
 * lts.decaytree[i].p4, which t_hat / u_hat and the tensor-pomeron vertices keep reading */
static int loop_kinematics_C(ost *S, double p1px, double p1py, double p2px, double p2py) {
  const grb_process_desc *d = S->d;
  const int Kf = d->n_legs;
  S->pfinal[1].x = p1px;
  S->pfinal[1].y = p1py;
  S->pfinal[2].x = p2px;
  S->pfinal[2].y = p2py;
  v4 p[8];
  blinear(Kf, S->pkt, S->pfinal[1], S->pfinal[2], p);
  v4 sumP = {0, 0, 0, 0};
  for (int i = 0; i < Kf; ++i) {
    const double y = v4_rap(S->pfinal_orig[i + 3]);
    const double m = v4_m(S->pfinal_orig[i + 3]);
    S->pfinal[i + 3].x = p[i].x;
    S->pfinal[i + 3].y = p[i].y;
    const double mt = msqrt(pow2(m) + v4_pt2(S->pfinal[i + 3]));
    S->pfinal[i + 3].z = mt * sinh(y);
    S->pfinal[i + 3].e = mt * cosh(y);
    sumP = v4_add(sumP, S->pfinal[i + 3]);
  }
  const double pt1 = v4_pt(S->pfinal[1]);
  const double pt2 = v4_pt(S->pfinal[2]);
  const double m1  = v4_m(S->pfinal_orig[1]);
  const double m2  = v4_m(S->pfinal_orig[2]);
  const double p1z = solve_pz(m1, m2, pt1, pt2, sumP.z, sumP.e, d->s);
  const double p2z = -(sumP.z + p1z);
  if (isnan(p1z)) return 0;
  if (p1z < 0 || p2z > 0) return 0;
  S->pfinal[1].z = p1z;
  S->pfinal[1].e = msqrt(pow2(m1) + pow2(pt1) + pow2(p1z));
  S->pfinal[2].z = p2z;
  S->pfinal[2].e = msqrt(pow2(m2) + pow2(pt2) + pow2(p2z));
  S->pfinal[0] = sumP;
  if (!check_emc(v4_sub(S->beamsum, v4_add(v4_add(S->pfinal[1], S->pfinal[2]), S->pfinal[0])))) return 0;
  return lorentz_scalars(S, d->n_legs + 2);
}

/* ---- cuts ------------------------------------------------------------------------------------------- */

/* reference src/MUserCuts.cc:21-75 (ids 0, +-3, +-111) */
static int user_cut(const ost *S, int id) {
  if (id == 0) return 1;
  if (id == -3) return v4_pt(v4_sub(S->pfinal[1], S->pfinal[2])) < 0.3;
  if (id == 3) return v4_pt(v4_sub(S->pfinal[1], S->pfinal[2])) > 0.3;
  const double dot = S->pfinal[1].x * S->pfinal[2].x + S->pfinal[1].y * S->pfinal[2].y;
  if (id == -111) return dot < 0;
  if (id == 111) return dot > 0;
  return 1;
}

/* reference src/MProcess.cc:1138-1155 FindDecayCuts */
static void find_decay_cuts(const ost *S, int node, int *ok) {
  const grb_node *  b = &S->d->nodes[node];
  const grb_fidcut *f = &S->d->fcuts;
  if (b->n_children == 0) {
    const v4 p = S->node_p4[node];
    if (v4_pt(p) >= f->pt_min && v4_pt(p) <= f->pt_max && v4_et(p) >= f->Et_min && v4_et(p) <= f->Et_max &&
        v4_eta(p) >= f->eta_min && v4_eta(p) <= f->eta_max && v4_rap(p) >= f->rap_min &&
        v4_rap(p) <= f->rap_max) {
    } else {
      *ok = 0;
    }
  }
  for (int i = 0; i < b->n_children; ++i) find_decay_cuts(S, b->first_child + i, ok);
}

/* reference src/MProcess.cc:1085-1135 CommonCuts (no N* excitation) */
static int common_cuts(const ost *S) {
  const grb_fidcut *f  = &S->d->fcuts;
  int               ok = 1;
  if (f->active) {
    if (!user_cut(S, f->usercuts)) return 0;
    if (fabs(S->t1) >= f->forward_t_min && fabs(S->t1) <= f->forward_t_max && fabs(S->t2) >= f->forward_t_min &&
        fabs(S->t2) <= f->forward_t_max) {
    } else {
      return 0;
    }
    if (S->excite1 && !(v4_m(S->pfinal[1]) >= f->forward_M_min && v4_m(S->pfinal[1]) <= f->forward_M_max)) return 0;
    if (S->excite2 && !(v4_m(S->pfinal[2]) >= f->forward_M_min && v4_m(S->pfinal[2]) <= f->forward_M_max)) return 0;
    if (msqrt(S->m2) >= f->M_min && msqrt(S->m2) <= f->M_max && S->Y >= f->Y_min && S->Y <= f->Y_max &&
        S->Pt >= f->Pt_min && S->Pt <= f->Pt_max) {
    } else {
      return 0;
    }
    for (int i = 0; i < S->d->n_legs; ++i) find_decay_cuts(S, i, &ok);
  }
  return ok;
}

/* reference src/MProcess.cc:1157-1175 FindVetoCuts, :1069-1083 VetoCuts (central system only: forward
 * legs are final protons with NSTARS=0 and decayforward{1,2} hold no fragments) */
static void find_veto_cuts(const ost *S, int node, int *ok) {
  const grb_node *b = &S->d->nodes[node];
  if (b->n_children == 0) {
    const v4 p = S->node_p4[node];
    for (int i = 0; i < S->d->n_veto; ++i) {
      const grb_vetodomain *v = &S->d->veto[i];
      if (v4_pt(p) >= v->pt_min && v4_pt(p) <= v->pt_max && v4_eta(p) >= v->eta_min && v4_eta(p) <= v->eta_max)
        *ok = 0;
    }
  }
  for (int i = 0; i < b->n_children; ++i) find_veto_cuts(S, b->first_child + i, ok);
}
static int veto_cuts(const ost *S) {
  int ok = 1;
  if (S->d->veto_active)
    for (int i = 0; i < S->d->n_legs; ++i) find_veto_cuts(S, i, &ok);
  return ok;
}

/* ---- Regge continuum amplitude ---------------------------------------------------------------------- */

/* reference src/MForm.cc:382-385 ReggeEta */
static cx regge_eta(double alpha_t, double sigma) {
  const double denom = sin(O_PI * alpha_t);
  /* exp(-zi*PI*alpha_t): real part of the argument is (-0.0*PI)*alpha_t = +-0 -> exp = 1 */
  const double y = (-1.0 * O_PI) * alpha_t;
  const cx     e = {cos(y), sin(y)};
  const cx     v = {1.0 + sigma * e.re, sigma * e.im};
  return (cx){-v.re / denom, -v.im / denom};
}

/* reference src/MRegge.cc:888-906 PropOnly with MForm.cc:394-396 ReggeEtaLinear */
static cx prop_only(const ost *S, double s, double t) {
  const grb_param_regge *R = &S->d->regge;
  cx                     M = {0, 0};
  for (int k = 0; k < 3; ++k) {
    if (R->n[k]) {
      const double alpha = R->a0[k] + R->ap[k] * t;
      const double y     = ((-1.0 * O_PI) / 2.0 * R->ap[k]) * t;
      const cx     ph    = {cos(y), sin(y)};
      const cx     eta   = cx_mul(S->regge_eta0[k], ph);
      M = cx_add(M, cx_scale(eta, pow(s / R->s0, alpha)));
    }
  }
  return M;
}

/* reference src/MForm.cc:433-435 S3F */
static double s3f(const ost *S, double t) {
  return (1.0 / (1 - t / S->d->soft.fc1)) * (1.0 / (1 - t / S->d->soft.fc2));
}

/* reference src/MRegge.cc:1155-1218 Meson_FF (== Baryon_FF), Meson_prop */
static double meson_ff(const ost *S, double t_hat, double M2) {
  const grb_param_regge *R      = &S->d->regge;
  const double           tprime = t_hat - M2;
  switch (R->offshellFF) {
    case 0: return exp(-R->b_EXP * fabs(tprime));
    case 1: return exp(-R->b_OREAR * msqrt(fabs(tprime) + pow2(R->a_OREAR)) + R->a_OREAR * R->b_OREAR);
    default: return 1.0 / (1.0 - tprime / R->b_POW);
  }
}
static double meson_prop(double t_hat, double M2) { return 1.0 / (t_hat - M2); }

/* reference src/MRegge.cc:394-438 ME4 (+ :637-644 PomPomProtonVertex, elastic legs) */
/* reference src/MForm.cc:443-449 S3FINEL, src/MRegge.cc:637-644 PomPomProtonVertex */
static double s3finel(double t, double M2) {
  const double DELTA_P = 0.0808, a = 0.5616;
  return pow(fabs(t) / (M2 * (fabs(t) + a)), 0.5 * (1 + DELTA_P));
}
static void pom_pom_proton_vertex(const ost *S, double *FF_A, double *FF_B) {
  const grb_param_soft *p = &S->d->soft;
  *FF_A = S->excite1 ? msqrt(p->g3P * p->gN_P) * s3finel(S->t1, v4_m2(S->pfinal[1])) : p->gN_P * s3f(S, S->t1);
  *FF_B = S->excite2 ? msqrt(p->g3P * p->gN_P) * s3finel(S->t2, v4_m2(S->pfinal[2])) : p->gN_P * s3f(S, S->t2);
}
static cx regge_me4(ost *S, double sign) {
  const grb_process_desc *d   = S->d;
  const double            M2_ = pow2(d->nodes[0].mass);
  double                  FF_A, FF_B;
  pom_pom_proton_vertex(S, &FF_A, &FF_B);
  const double            gpp_P = d->regge.c[0] / d->soft.gN_P;
  cx A_t = prop_only(S, S->ss[1][3], S->t1);
  A_t    = cx_scale(A_t, FF_A);
  A_t    = cx_scale(A_t, meson_ff(S, S->t_hat, M2_));
  A_t    = cx_scale(A_t, gpp_P);
  A_t    = cx_scale(A_t, meson_prop(S->t_hat, M2_));
  A_t    = cx_scale(A_t, meson_ff(S, S->t_hat, M2_));
  A_t    = cx_scale(A_t, gpp_P);
  A_t    = cx_mul(A_t, prop_only(S, S->ss[2][4], S->t2));
  A_t    = cx_scale(A_t, FF_B);
  cx A_u = cx_scale(prop_only(S, S->ss[1][4], S->t1), sign);
  A_u    = cx_scale(A_u, FF_A);
  A_u    = cx_scale(A_u, meson_ff(S, S->u_hat, M2_));
  A_u    = cx_scale(A_u, gpp_P);
  A_u    = cx_scale(A_u, meson_prop(S->u_hat, M2_));
  A_u    = cx_scale(A_u, meson_ff(S, S->u_hat, M2_));
  A_u    = cx_scale(A_u, gpp_P);
  A_u    = cx_mul(A_u, prop_only(S, S->ss[2][3], S->t2));
  A_u    = cx_scale(A_u, FF_B);
  const cx A = cx_add(A_t, A_u);
  S->hamp[0] = A;
  S->nhamp   = 1;
  return A;
}

/* ---- tensor pomeron continuum (Ewerz-Maniatis-Nachtmann), 2 x pseudoscalar ------------------------- */
/* Restated with explicit rank-2/rank-4 Lorentz tensors and index loops exactly as the reference builds
 * them with FTensor (src/MTensorPomeron.cc:572-763); the device code uses the contracted closed form, so
 * the two are independent derivations.  Index 0 = E, metric g = diag(1,-1,-1,-1) (MTensorPomeron.h). */
static const double o_g[4][4] = {{1, 0, 0, 0}, {0, -1, 0, 0}, {0, 0, -1, 0}, {0, 0, 0, -1}};
static inline double v4_cov(v4 a, int mu) { /* M4Vec::operator%, M4Vec.h:230-237 */
  return mu == 0 ? a.e : (mu == 1 ? -a.x : (mu == 2 ? -a.y : -a.z));
}
static inline cx cx_mulr(cx a, double d) { return (cx){a.re * d, a.im * d}; }

/* reference src/MForm.cc:542-645 F1 with Kelly / dipole Sachs form factors */
static double o_mu_ratio(void) { return 2.792847337; }
static double o_GM(const ost *S, double Q2);
static double o_GE(const ost *S, double Q2) {
  const double mp = 0.938272081;
  Q2 = fabs(Q2);
  if (S->d->em_ff == 1) {
    const double tau = Q2 / pow2(2 * mp);
    double num = 0.0;
    num += 1;
    num += -0.24 * tau;
    double den = 1.0;
    den += 10.98 * tau;
    den += 12.82 * pow2(tau);
    den += 21.97 * (tau * pow2(tau));
    return num / den;
  }
  return o_GM(S, Q2) / o_mu_ratio();
}
static double o_GM(const ost *S, double Q2) {
  const double mp = 0.938272081;
  Q2 = fabs(Q2);
  if (S->d->em_ff == 1) {
    const double tau = Q2 / pow2(2 * mp);
    double num = 0.0;
    num += 1;
    num += 0.12 * tau;
    double den = 1.0;
    den += 10.97 * tau;
    den += 18.86 * pow2(tau);
    den += 6.55 * (tau * pow2(tau));
    return o_mu_ratio() * num / den;
  }
  return o_mu_ratio() / pow2(1.0 + Q2 / 0.71);
}
static double o_F1(const ost *S, double Q2) {
  const double mp = 0.938272081;
  Q2 = fabs(Q2);
  const double tau = Q2 / pow2(2 * mp);
  return 1.0 / (tau + 1) * o_GE(S, Q2) + tau / (tau + 1) * o_GM(S, Q2);
}

/* std::pow(complex z, double a) as libstdc++ evaluates it: real pow for positive real z, else
 * polar(exp(a * Re clog z), a * Im clog z) with glibc's clog */
static cx cx_powr(cx z, double a) {
  if (z.im == 0.0 && z.re > 0.0) return (cx){pow(z.re, a), 0.0};
  const double complex t = clog(__builtin_complex(z.re, z.im));
  const double rho = exp(a * creal(t)), th = a * cimag(t);
  return (cx){rho * cos(th), rho * sin(th)};
}

/* reference src/MTensorPomeron.cc:2288-2298 iD_P: FACTOR * (g g + g g - 1/2 g g) */
static cx tp_iDP_factor(const ost *S, double s, double t) {
  const grb_param_tensorpom *T = &S->d->tpom;
  const double alpha = 1.0 + T->delta_P + T->ap_P * t;
  /* -zi * s * ap_P = ((-0,-1)*s)*ap */
  const cx z = {-0.0 * s * T->ap_P, -1.0 * s * T->ap_P};
  return cx_mulr(cx_powr(z, alpha - 1.0), 1.0 / (4.0 * s));
}
static void tp_iDP(const ost *S, double s, double t, cx D[4][4][4][4]) {
  const cx F = tp_iDP_factor(S, s, t);
  for (int u = 0; u < 4; ++u)
    for (int v = 0; v < 4; ++v)
      for (int k = 0; k < 4; ++k)
        for (int l = 0; l < 4; ++l)
          D[u][v][k][l] = cx_mulr(F, o_g[u][k] * o_g[v][l] + o_g[u][l] * o_g[v][k] - 0.5 * o_g[u][v] * o_g[k][l]);
}
/* :1388-1399 iG_PppHE */
static void tp_iG_PppHE(const ost *S, v4 prime, v4 p, cx T[4][4]) {
  const v4 psum = v4_add(prime, p);
  const double f = 3.0 * S->d->tpom.gPNN * o_F1(S, v4_m2(v4_sub(prime, p)));
  const cx FACTOR = {-0.0 * f, -1.0 * f};
  for (int u = 0; u < 4; ++u)
    for (int v = 0; v < 4; ++v) T[u][v] = cx_mulr(cx_mulr(FACTOR, v4_cov(psum, u)), v4_cov(psum, v)); /* FACTOR * (psum % u) * (psum % v): left to right */
}
/* :2203-2215 iG_Ppsps */
static void tp_iG_Ppsps(v4 prime, v4 p, double g1, cx T[4][4]) {
  const v4 psum = v4_add(prime, p);
  const double M2 = v4_m2(psum);
  const double f = 2.0 * g1 * (1.0 / (1.0 - v4_m2(v4_sub(prime, p)) / 0.5));
  const cx FACTOR = {-0.0 * f, -1.0 * f};
  for (int u = 0; u < 4; ++u)
    for (int v = 0; v < 4; ++v)
      T[u][v] = cx_mulr(FACTOR, v4_cov(psum, u) * v4_cov(psum, v) - 0.25 * o_g[u][v] * M2);
}
/* A(mu,nu) D(mu,nu,a,b) Ga(a,b) Gb(c,d) D2(c,d,k,l) B(k,l), contracted left to right */
static cx tp_chain(cx A[4][4], cx D1[4][4][4][4], cx Ga[4][4], cx Gb[4][4], cx D2[4][4][4][4], cx B[4][4]) {
  /* FTensor: T4*T2 double contractions loop the contracted pair in natural order from result(0)
   * (Tensor4_times_Tensor2_double.hpp:17-27); T2*T2 double contractions recurse with the FIRST index fastest
   * (Tensor2_times_Tensor2.hpp:15-46) */
  cx s1 = {0, 0};
  for (int b = 0; b < 4; ++b)
    for (int a = 0; a < 4; ++a) {
      cx X = {0, 0};
      for (int m = 0; m < 4; ++m)
        for (int n = 0; n < 4; ++n) X = cx_add(X, cx_mul(D1[m][n][a][b], A[m][n]));
      s1 = (a == 0 && b == 0) ? cx_mul(X, Ga[a][b]) : cx_add(cx_mul(X, Ga[a][b]), s1);
    }
  cx r = {0, 0};
  for (int l = 0; l < 4; ++l)
    for (int k = 0; k < 4; ++k) {
      cx Z = {0, 0};
      for (int c = 0; c < 4; ++c)
        for (int e = 0; e < 4; ++e) Z = cx_add(Z, cx_mul(D2[c][e][k][l], cx_mul(s1, Gb[c][e])));
      r = (k == 0 && l == 0) ? cx_mul(Z, B[k][l]) : cx_add(cx_mul(Z, B[k][l]), r);
    }
  return r;
}

/* reference src/MTensorPomeron.cc:572-763,892-898 ME4, SPINMODE "2xS" */
static double tensor_me4(ost *S) {
  const grb_process_desc *d = S->d;
  if (!(d->nodes[0].spinX2 == 0 && d->nodes[1].spinX2 == 0)) {
    S->nhamp = 0;
    return NAN; /* 2xF / 2xV not restated yet */
  }
  const v4 pa = S->pbeam1, pb = S->pbeam2, p1 = S->pfinal[1], p2 = S->pfinal[2];
  const v4 p3 = S->node_p4[0], p4 = S->node_p4[1];
  const double M_ = d->nodes[0].mass;
  const v4 pt = v4_sub(v4_sub(pa, p1), p3);
  const v4 pu = v4_add(v4_sub(p4, pa), p1);
  const v4 mp3 = {-p3.x, -p3.y, -p3.z, -p3.e};
  static cx D13[4][4][4][4], D24[4][4][4][4], D14[4][4][4][4], D23[4][4][4][4];
  cx G1a[4][4], G2b[4][4], Gta[4][4], Gtb[4][4], Gua[4][4], Gub[4][4];
  tp_iDP(S, S->ss[1][3], S->t1, D13);
  tp_iDP(S, S->ss[2][4], S->t2, D24);
  tp_iDP(S, S->ss[1][4], S->t1, D14);
  tp_iDP(S, S->ss[2][3], S->t2, D23);
  tp_iG_PppHE(S, p1, pa, G1a);
  tp_iG_PppHE(S, p2, pb, G2b);
  const double gP = (d->nodes[0].mass < 0.2) ? d->tpom.gPpipi : d->tpom.gPKK;
  tp_iG_Ppsps(pt, mp3, gP, Gta);
  tp_iG_Ppsps(p4, pt, gP, Gtb);
  tp_iG_Ppsps(p4, pu, gP, Gua);
  tp_iG_Ppsps(pu, mp3, gP, Gub);
  const cx iDMES_t = {0.0 / (v4_m2(pt) - pow2(M_)), 1.0 / (v4_m2(pt) - pow2(M_))};
  const cx iDMES_u = {0.0 / (v4_m2(pu) - pow2(M_)), 1.0 / (v4_m2(pu) - pow2(M_))};
  cx M_t = tp_chain(G1a, D13, Gta, Gtb, D24, G2b);
  M_t    = cx_mul(M_t, cx_mulr(iDMES_t, pow2(meson_ff(S, v4_m2(pt), pow2(M_)))));
  cx M_u = tp_chain(G1a, D14, Gua, Gub, D23, G2b);
  M_u    = cx_mul(M_u, cx_mulr(iDMES_u, pow2(meson_ff(S, v4_m2(pu), pow2(M_)))));
  const cx sum = cx_add(M_t, M_u);
  const cx M   = cx_mul((cx){-0.0, -1.0}, sum);
  /* four helicity-conserving proton combinations push the same amplitude (:658-662,762) */
  S->nhamp = 4;
  double SumAmp2 = 0.0;
  for (int h = 0; h < 4; ++h) {
    S->hamp[h] = M;
    SumAmp2 += cx_abs2(M);
  }
  return SumAmp2 / 4;
}

/* The same 2xS amplitude in x87 extended precision (64-bit mantissa) from the same double inputs, through
 * the contracted closed form: used by the tests to MEASURE the rounding noise nu of the reference's
 * double-precision FTensor contraction (terms of size (E_beam K)^2 cancel down to (P.K)^2: factor 1e3-1e5),
 * so that device-vs-reference differences can be judged against max(1e-10, 8 kappa, 8 nu). */
static double tensor_me4_extended(ost *S, double *re, double *im) {
  typedef long double R;
  const grb_process_desc *d = S->d;
  const v4 pa = S->pbeam1, pb = S->pbeam2, p1 = S->pfinal[1], p2 = S->pfinal[2];
  const v4 p3 = S->node_p4[0], p4 = S->node_p4[1];
#define LV(n, X, Y, Z, E) const R n##x = (X), n##y = (Y), n##z = (Z), n##e = (E)
#define LDOT(a, b) (a##e * b##e - (a##x * b##x + a##y * b##y + a##z * b##z))
  LV(P1, (R)p1.x + pa.x, (R)p1.y + pa.y, (R)p1.z + pa.z, (R)p1.e + pa.e);
  LV(P2, (R)p2.x + pb.x, (R)p2.y + pb.y, (R)p2.z + pb.z, (R)p2.e + pb.e);
  LV(q1, (R)pa.x - p1.x, (R)pa.y - p1.y, (R)pa.z - p1.z, (R)pa.e - p1.e);
  LV(q2, (R)pb.x - p2.x, (R)pb.y - p2.y, (R)pb.z - p2.z, (R)pb.e - p2.e);
  LV(pt, q1x - p3.x, q1y - p3.y, q1z - p3.z, q1e - p3.e);
  LV(pu, p4.x - q1x, p4.y - q1y, p4.z - q1z, p4.e - q1e);
  LV(Kta, ptx - p3.x, pty - p3.y, ptz - p3.z, pte - p3.e);
  LV(Ktb, p4.x + ptx, p4.y + pty, p4.z + ptz, p4.e + pte);
  LV(Kua, p4.x + pux, p4.y + puy, p4.z + puz, p4.e + pue);
  LV(Kub, pux - p3.x, puy - p3.y, puz - p3.z, pue - p3.e);
  (void)q2x; (void)q2y; (void)q2z; (void)q2e;
  /* invariants the reference hands to the amplitude as doubles (lts.t1, lts.ss[][], pt.M2()) stay doubles:
   * only the tensor contraction itself is lifted to extended precision */
  const R t1 = S->t1, t2 = S->t2;
  const R P1sq = LDOT(P1, P1), P2sq = LDOT(P2, P2);
  const R d1t = LDOT(P1, Kta), d2t = LDOT(P2, Ktb), d1u = LDOT(P1, Kua), d2u = LDOT(P2, Kub);
  const R Rt = (2 * d1t * d1t - 0.5L * LDOT(Kta, Kta) * P1sq) * (2 * d2t * d2t - 0.5L * LDOT(Ktb, Ktb) * P2sq);
  const R Ru = (2 * d1u * d1u - 0.5L * LDOT(Kua, Kua) * P1sq) * (2 * d2u * d2u - 0.5L * LDOT(Kub, Kub) * P2sq);
  const R pt2 = v4_m2(v4_sub(v4_sub(pa, p1), p3)), pu2 = v4_m2(v4_add(v4_sub(p4, pa), p1));
  const R s13 = S->ss[1][3], s24 = S->ss[2][4], s14 = S->ss[1][4], s23 = S->ss[2][3];
  const R ap = d->tpom.ap_P, a1 = d->tpom.delta_P + ap * t1, a2 = d->tpom.delta_P + ap * t2;
  const R Et = expl(a1 * logl(s13 * ap) + a2 * logl(s24 * ap)) / (16 * s13 * s24);
  const R Eu = expl(a1 * logl(s14 * ap) + a2 * logl(s23 * ap)) / (16 * s14 * s23);
  const R M2 = (R)d->nodes[0].mass * d->nodes[0].mass;
  R Gt, Gu;
  {
    const grb_param_regge *Rg = &d->regge;
    R tp = pt2 - M2, tu = pu2 - M2, ft, fu;
    if (Rg->offshellFF == 2) { ft = 1 / (1 - tp / Rg->b_POW); fu = 1 / (1 - tu / Rg->b_POW); }
    else if (Rg->offshellFF == 0) { ft = expl(-Rg->b_EXP * fabsl(tp)); fu = expl(-Rg->b_EXP * fabsl(tu)); }
    else { ft = expl(-Rg->b_OREAR * sqrtl(fabsl(tp) + (R)Rg->a_OREAR * Rg->a_OREAR) + (R)Rg->a_OREAR * Rg->b_OREAR);
           fu = expl(-Rg->b_OREAR * sqrtl(fabsl(tu) + (R)Rg->a_OREAR * Rg->a_OREAR) + (R)Rg->a_OREAR * Rg->b_OREAR); }
    Gt = ft * ft / tp; Gu = fu * fu / tu;
  }
  const R gP = (d->nodes[0].mass < 0.2) ? d->tpom.gPpipi : d->tpom.gPKK;
  /* F_M arguments as the reference forms them per vertex, iG_Ppsps(k',k): (k'-k)^2 (doubles) */
  const v4 dpt = v4_sub(v4_sub(pa, p1), p3), dpu = v4_add(v4_sub(p4, pa), p1);
  const R FMt = (1 / (1 - (R)v4_m2(v4_add(dpt, p3)) / 0.5L)) * (1 / (1 - (R)v4_m2(v4_sub(p4, dpt)) / 0.5L));
  const R FMu = (1 / (1 - (R)v4_m2(v4_sub(p4, dpu)) / 0.5L)) * (1 / (1 - (R)v4_m2(v4_add(dpu, p3)) / 0.5L));
  const R pre = 36 * (R)d->tpom.gPNN * d->tpom.gPNN * gP * gP * (R)o_F1(S, (double)t1) * (R)o_F1(S, (double)t2);
  const R Rr = pre * (Et * Rt * FMt * Gt + Eu * Ru * FMu * Gu);
  const R ph = -(R)3.141592653589793238462643383279502884L / 2 * (a1 + a2);
  *re = (double)(cosl(ph) * Rr);
  *im = (double)(sinl(ph) * Rr);
#undef LV
#undef LDOT
  return (double)(Rr * Rr);
}

/* ---- resonances: MRegge::ME3 / PhotoME3 with Jacob-Wick spin weights --------------------------------- */
static inline cx cx_conj(cx a) { return (cx){a.re, -a.im}; }
/* complex / complex as libgcc __divdc3 evaluates it (what std::complex<double>::operator/= compiles to) */
static cx cx_div(cx x, cx y) {
  const double a = x.re, b = x.im, c = y.re, d = y.im;
  double ratio, denom, re, im;
  if (fabs(c) < fabs(d)) {
    ratio = c / d; denom = (c * ratio) + d;
    re = ((a * ratio) + b) / denom; im = ((b * ratio) - a) / denom;
  } else {
    ratio = d / c; denom = (d * ratio) + c;
    re = ((b * ratio) + a) / denom; im = (b - (a * ratio)) / denom;
  }
  return (cx){re, im};
}
static double v4_theta(v4 a) { return (a.x == 0.0 && a.y == 0.0 && a.z == 0.0) ? 0.0 : atan2(v4_pt(a), a.z); }
static double v4_phi(v4 a) { return (a.x == 0.0 && a.y == 0.0) ? 0.0 : atan2(a.y, a.x); }
/* ---- tensor pomeron continuum, 2 x vector meson (reference src/MTensorPomeron.cc:797-888) -------------------- */
/* :2500-2536 Gamma0 / Gamma2 rank-4 tensor functions (covariant components, M4Vec::operator%) */
static void tp_gamma0(v4 k1, v4 k2, double T[4][4][4][4]) {
  const double k1k2 = v4_dot(k1, k2);
  for (int u = 0; u < 4; ++u)
    for (int v = 0; v < 4; ++v)
      for (int k = 0; k < 4; ++k)
        for (int l = 0; l < 4; ++l)
          T[u][v][k][l] = (k1k2 * o_g[u][v] - v4_cov(k2, u) * v4_cov(k1, v)) *
                          (v4_cov(k1, k) * v4_cov(k2, l) + v4_cov(k2, k) * v4_cov(k1, l) - 0.5 * k1k2 * o_g[k][l]);
}
static void tp_gamma2(v4 k1, v4 k2, double T[4][4][4][4]) {
  const double k1k2 = v4_dot(k1, k2);
  for (int u = 0; u < 4; ++u)
    for (int v = 0; v < 4; ++v)
      for (int k = 0; k < 4; ++k)
        for (int l = 0; l < 4; ++l)
          T[u][v][k][l] = k1k2 * (o_g[u][k] * o_g[v][l] + o_g[u][l] * o_g[v][k]) +
                          o_g[u][v] * (v4_cov(k1, k) * v4_cov(k2, l) + v4_cov(k2, k) * v4_cov(k1, l)) -
                          v4_cov(k1, v) * v4_cov(k2, l) * o_g[u][k] - v4_cov(k1, v) * v4_cov(k2, k) * o_g[u][l] -
                          v4_cov(k2, u) * v4_cov(k1, l) * o_g[v][k] - v4_cov(k2, u) * v4_cov(k1, k) * o_g[v][l] -
                          (k1k2 * o_g[u][v] - v4_cov(k2, u) * v4_cov(k1, v)) * o_g[k][l];
}
/* :2224-2238 iG_Pvv(prime, p, g1, g2) */
static void tp_iG_Pvv(v4 prime, v4 p, double g1, double g2, cx T[4][4][4][4]) {
  static double G0[4][4][4][4], G2[4][4][4][4];
  const v4 mp = {-p.x, -p.y, -p.z, -p.e};
  tp_gamma0(prime, mp, G0);
  tp_gamma2(prime, mp, G2);
  const double fm = 1.0 / (1.0 - v4_m2(v4_sub(prime, p)) / 0.5); /* FM(q2, 0.5), :2486 */
  const cx FACTOR = {0.0 * fm, 1.0 * fm};
  for (int u = 0; u < 4; ++u)
    for (int v = 0; v < 4; ++v)
      for (int k = 0; k < 4; ++k)
        for (int l = 0; l < 4; ++l) T[u][v][k][l] = cx_mulr(FACTOR, 2.0 * g1 * G0[u][v][k][l] - g2 * G2[u][v][k][l]);
}
/* :2353-2368 iD_V: Reggeized vector propagator, T(u,v) = -g[u][v] / (m2 - M0^2) * (e^{i phi} s34 / s_thresh)^(alpha - 1) */
static void tp_iD_V(v4 p, double M0, double s34, cx T[4][4]) {
  const double m2 = v4_m2(p);
  const double alpha = 0.5 + 0.9 * m2;
  const double s_thresh = 4 * pow2(M0);
  const double phi = O_PI / 2.0 * exp((s_thresh - s34) / s_thresh) - O_PI / 2.0;
  const cx e = {exp(0.0 * phi) * cos(phi), exp(0.0 * phi) * sin(phi)}; /* std::exp(zi * phi) */
  const cx z = {e.re * s34 / s_thresh, e.im * s34 / s_thresh};
  const cx R = cx_powr(z, alpha - 1.0);
  for (int u = 0; u < 4; ++u)
    for (int v = 0; v < 4; ++v) T[u][v] = cx_mulr(R, -o_g[u][v] / (m2 - pow2(M0)));
}
/* src/MDirac.cc:393-470,661-679 massive spin-1 polarisation vectors, complex conjugated, contravariant [t,x,y,z] */
static void o_eps_massive_conj(v4 k, int m, cx e[4]) {
  const double theta = v4_theta(k), phi = v4_phi(k);
  if (m != 0) {
    const double N = 1.0 / sqrt(2.0), neg = -(double)m, pos = (double)m;
    e[0] = (cx){0.0, 0.0};
    e[1] = (cx){(neg * cos(theta) * cos(phi)) * N, -(sin(phi) * N)};
    e[2] = (cx){(neg * cos(theta) * sin(phi)) * N, (cos(phi) * N)};
    e[3] = (cx){pos * sin(theta) * N, -0.0};
    return;
  }
  const double E = k.e, M = v4_m(k);
  e[0] = (cx){v4_p3mod(k) / M, -0.0};
  e[1] = (cx){E * sin(theta) * cos(phi) / M, -0.0};
  e[2] = (cx){E * sin(theta) * sin(phi) / M, -0.0};
  e[3] = (cx){E * cos(theta) / M, -0.0};
}
/* Upper/lower block: R(a,b) = G(mu,nu) D(mu,nu,al,be) V(a,b,al,be) */
static void tp_block(cx G[4][4], cx D[4][4][4][4], cx V[4][4][4][4], cx R[4][4]) {
  cx X[4][4];
  for (int a = 0; a < 4; ++a)
    for (int b = 0; b < 4; ++b) {
      cx x = {0, 0};
      for (int m = 0; m < 4; ++m)
        for (int n = 0; n < 4; ++n) x = cx_add(x, cx_mul(D[m][n][a][b], G[m][n]));
      X[a][b] = x;
    }
  for (int r = 0; r < 4; ++r)
    for (int q = 0; q < 4; ++q) {
      cx x = {0, 0};
      for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) x = cx_add(x, cx_mul(X[a][b], V[r][q][a][b]));
      R[r][q] = x;
    }
}
static double tensor_me4_2xV(ost *S) {
  const grb_process_desc *d = S->d;
  const v4 pa = S->pbeam1, pb = S->pbeam2, p1 = S->pfinal[1], p2 = S->pfinal[2];
  const v4 p3 = S->node_p4[0], p4 = S->node_p4[1];
  const double M_ = d->nodes[0].mass;
  const v4 pt = v4_sub(v4_sub(pa, p1), p3);
  const v4 pu = v4_add(v4_sub(p4, pa), p1);
  const v4 mp3 = {-p3.x, -p3.y, -p3.z, -p3.e};
  double g1 = 0.0, g2 = 0.0;
  if (d->nodes[0].pdg == 113) { g1 = d->tpom.gPrhorho[0]; g2 = d->tpom.gPrhorho[1]; }
  else if (d->nodes[0].pdg == 333) { g1 = d->tpom.gPphiphi[0]; g2 = d->tpom.gPphiphi[1]; }
  static cx D13[4][4][4][4], D24[4][4][4][4], D14[4][4][4][4], D23[4][4][4][4];
  static cx GtA[4][4][4][4], GtB[4][4][4][4], GuA[4][4][4][4], GuB[4][4][4][4];
  cx G1a[4][4], G2b[4][4];
  tp_iDP(S, S->ss[1][3], S->t1, D13);
  tp_iDP(S, S->ss[2][4], S->t2, D24);
  tp_iDP(S, S->ss[1][4], S->t1, D14);
  tp_iDP(S, S->ss[2][3], S->t2, D23);
  tp_iG_PppHE(S, p1, pa, G1a);
  tp_iG_PppHE(S, p2, pb, G2b);
  tp_iG_Pvv(pt, mp3, g1, g2, GtA);
  tp_iG_Pvv(p4, pt, g1, g2, GtB);
  tp_iG_Pvv(p4, pu, g1, g2, GuA);
  tp_iG_Pvv(pu, mp3, g1, g2, GuB);
  const double s34 = v4_m2(S->pfinal[0]);
  cx DVt[4][4], DVu[4][4];
  tp_iD_V(pt, M_, s34, DVt);
  tp_iD_V(pu, M_, s34, DVu);
  const double fft = pow2(meson_ff(S, v4_m2(pt), pow2(M_))), ffu = pow2(meson_ff(S, v4_m2(pu), pow2(M_)));
  cx A[4][4], B[4][4], Y[4][4], Mt[4][4], Mu[4][4], M[4][4];
  /* single contractions recurse from the top index: t3 + (t2 + (t1 + t0)) (Tensor2_times_Tensor2.hpp:105-290) */
#define SUM4(expr) ({ cx acc_; { const int q = 0; acc_ = (expr); } for (int q = 1; q < 4; ++q) acc_ = cx_add((expr), acc_); acc_; })
  /* t-channel: M_t(r3,r4) = [A(r1,r3) iDV(r1,r2)] B(r4,r2) * FF^2 */
  tp_block(G1a, D13, GtA, A);
  tp_block(G2b, D24, GtB, B);
  for (int r3 = 0; r3 < 4; ++r3)
    for (int r2 = 0; r2 < 4; ++r2) Y[r3][r2] = SUM4(cx_mul(A[q][r3], DVt[q][r2]));
  for (int r3 = 0; r3 < 4; ++r3)
    for (int r4 = 0; r4 < 4; ++r4) Mt[r3][r4] = cx_mulr(SUM4(cx_mul(Y[r3][q], B[r4][q])), fft);
  /* u-channel: M_u(r3,r4) = [A(r4,r1) iDV(r1,r2)] B(r2,r3) * FF^2 */
  tp_block(G1a, D14, GuA, A);
  tp_block(G2b, D23, GuB, B);
  for (int r4 = 0; r4 < 4; ++r4)
    for (int r2 = 0; r2 < 4; ++r2) Y[r4][r2] = SUM4(cx_mul(A[r4][q], DVu[q][r2]));
  for (int r3 = 0; r3 < 4; ++r3)
    for (int r4 = 0; r4 < 4; ++r4) Mu[r3][r4] = cx_mulr(SUM4(cx_mul(Y[r4][q], B[q][r3])), ffu);
#undef SUM4
  for (int u = 0; u < 4; ++u)
    for (int v = 0; v < 4; ++v) M[u][v] = cx_mul((cx){-0.0, -1.0}, cx_add(Mt[u][v], Mu[u][v]));
  /* MassiveSpin1PolSum (:1227-1247): [eps3*(r3) eps4*(r4)] M(r3,r4), double contraction with r3 fastest */
  cx e3[3][4], e4[3][4], h[9];
  for (int m = 0; m < 3; ++m) { o_eps_massive_conj(p3, m - 1, e3[m]); o_eps_massive_conj(p4, m - 1, e4[m]); }
  for (int h3 = 0; h3 < 3; ++h3)
    for (int h4 = 0; h4 < 3; ++h4) {
      cx x = {0, 0};
      for (int r4 = 0; r4 < 4; ++r4)
        for (int r3 = 0; r3 < 4; ++r3) {
          const cx t = cx_mul(cx_mul(e3[h3][r3], e4[h4][r4]), M[r3][r4]);
          x = (r3 == 0 && r4 == 0) ? t : cx_add(t, x);
        }
      h[3 * h3 + h4] = x;
    }
  /* the four helicity-conserving proton combinations push the same nine amplitudes */
  S->nhamp = 36;
  double SumAmp2 = 0.0;
  for (int k = 0; k < 36; ++k) {
    S->hamp[k] = h[k % 9];
    SumAmp2 += cx_abs2(S->hamp[k]);
  }
  return SumAmp2 / 4;
}

/* ---- tensor pomeron continuum, 2 x fermion (p pbar, Lambda Lambdabar) and the yy[QED] lepton branch ---------- */
/* reference src/MTensorPomeron.cc:664-713 (QED), :765-804 (2xF); Dirac algebra of src/MDirac.cc in the DIRAC basis.
 * Explicit 4x4 complex matrices and rank-4 tensors, as the reference builds them. */
typedef struct { cx m[4][4]; } m44;
static m44 og_up[4], og_lo[4], og_sig_lo[4][4];
static int og_ready = 0;
static m44 m44_zero(void) { m44 r; memset(&r, 0, sizeof r); return r; }
static m44 m44_mul(const m44 *a, const m44 *b) { /* MMatrix operator*: C[i][j] += A[i][k] B[k][j] */
  m44 r = m44_zero();
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      for (int k = 0; k < 4; ++k) r.m[i][j] = cx_add(r.m[i][j], cx_mul(a->m[i][k], b->m[k][j]));
  return r;
}
static m44 m44_add(const m44 *a, const m44 *b) {
  m44 r;
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r.m[i][j] = cx_add(a->m[i][j], b->m[i][j]);
  return r;
}
static m44 m44_sub(const m44 *a, const m44 *b) {
  m44 r;
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r.m[i][j] = (cx){a->m[i][j].re - b->m[i][j].re, a->m[i][j].im - b->m[i][j].im};
  return r;
}
static m44 m44_scale(const m44 *a, cx s) {
  m44 r;
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r.m[i][j] = cx_mul(a->m[i][j], s);
  return r;
}
static m44 m44_scaler(const m44 *a, double s) { return m44_scale(a, (cx){s, 0.0}); }
static void og_init(void) { /* MDirac::InitGammaMatrices("DIRAC"), src/MDirac.cc:47-140 */
  if (og_ready) return;
  static const double y0[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, -1, 0}, {0, 0, 0, -1}};
  static const double y1[4][4] = {{0, 0, 0, 1}, {0, 0, 1, 0}, {0, -1, 0, 0}, {-1, 0, 0, 0}};
  static const double y2i[4][4] = {{0, 0, 0, -1}, {0, 0, 1, 0}, {0, 1, 0, 0}, {-1, 0, 0, 0}}; /* times i */
  static const double y3[4][4] = {{0, 0, 1, 0}, {0, 0, 0, -1}, {-1, 0, 0, 0}, {0, 1, 0, 0}};
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      og_up[0].m[i][j] = (cx){y0[i][j], 0}; og_lo[0].m[i][j] = (cx){y0[i][j], 0};
      og_up[1].m[i][j] = (cx){y1[i][j], 0}; og_lo[1].m[i][j] = (cx){-y1[i][j], 0};
      og_up[2].m[i][j] = (cx){0, y2i[i][j]}; og_lo[2].m[i][j] = (cx){0, -y2i[i][j]};
      og_up[3].m[i][j] = (cx){y3[i][j], 0}; og_lo[3].m[i][j] = (cx){-y3[i][j], 0};
    }
  for (int mu = 0; mu < 4; ++mu)
    for (int nu = 0; nu < 4; ++nu) {
      const m44 ab = m44_mul(&og_lo[mu], &og_lo[nu]), ba = m44_mul(&og_lo[nu], &og_lo[mu]);
      const m44 c = m44_sub(&ab, &ba);
      og_sig_lo[mu][nu] = m44_scale(&c, (cx){0.0, 0.5}); /* (zi / 2.0) */
    }
  og_ready = 1;
}
static inline double v4_con(v4 a, int mu) { return mu == 0 ? a.e : (mu == 1 ? a.x : (mu == 2 ? a.y : a.z)); }
static m44 o_fslash(v4 a) { /* MDirac::FSlash, :506-510: sum_mu gamma_up[mu] * (a % mu) */
  m44 r = m44_zero();
  for (int mu = 0; mu < 4; ++mu) {
    const m44 t = m44_scaler(&og_up[mu], v4_cov(a, mu));
    r = m44_add(&r, &t);
  }
  return r;
}
static m44 o_iD_F(v4 q, double m) { /* MDirac::iD_F, :382-384 */
  m44 s = o_fslash(q);
  for (int i = 0; i < 4; ++i) s.m[i][i] = cx_add(s.m[i][i], (cx){m, 0.0});
  const cx f = cx_div((cx){0.0, 1.0}, (cx){v4_m2(q) - pow2(m), 0.0});
  return m44_scale(&s, f);
}
static void o_uHelDirac(v4 p, int hel, cx u[4]) { /* MDirac::uHelDirac, :224-246 */
  const double E = p.e, p3 = v4_p3mod(p), m = v4_m(p), N = msqrt(E + m);
  const double th2 = v4_theta(p) / 2.0, phi = v4_phi(p), c = cos(th2), s = sin(th2);
  const cx ph = {cos(phi), sin(phi)};
  if (hel == 1) {
    u[0] = (cx){N * c, 0}; u[1] = cx_mulr(cx_mulr(ph, N), s);
    u[2] = (cx){N * p3 / (E + m) * c, 0}; u[3] = cx_mulr(cx_mulr(ph, N * p3 / (E + m)), s);
  } else {
    u[0] = (cx){-N * s, 0}; u[1] = cx_mulr(cx_mulr(ph, N), c);
    u[2] = (cx){N * p3 / (E + m) * s, 0}; u[3] = cx_mulr(cx_mulr(ph, -N * p3 / (E + m)), c);
  }
}
static void o_vHelDirac(v4 p, int hel, cx v[4]) { /* :248-257 */
  cx u[4];
  o_uHelDirac(p, -hel, u);
  v[0] = u[2]; v[1] = u[3]; v[2] = u[0]; v[3] = u[1];
}
static void o_bar(const cx s[4], cx b[4]) { /* MDirac::Bar: conj(s) . gamma^0 */
  b[0] = cx_conj(s[0]); b[1] = cx_conj(s[1]);
  b[2] = (cx){-s[2].re, s[2].im}; b[3] = (cx){-s[3].re, s[3].im};
}
static cx o_vmv(const cx a[4], const m44 *B, const cx c[4]) { /* matoper::VecMatVecMultiply */
  cx v = {0, 0};
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) v = cx_add(v, cx_mul(cx_mul(a[i], B->m[i][j]), c[j]));
  return v;
}
static void o_vm(const cx a[4], const m44 *B, cx c[4]) { /* matoper::VecMatMultiply */
  for (int j = 0; j < 4; ++j) {
    c[j] = (cx){0, 0};
    for (int k = 0; k < 4; ++k) c[j] = cx_add(c[j], cx_mul(a[k], B->m[k][j]));
  }
}
static double o_F2(const ost *S, double Q2) { /* form::F2, src/MForm.cc:550-555 */
  const double mp = 0.938272081;
  Q2 = fabs(Q2);
  const double tau = Q2 / pow2(2 * mp);
  return -1.0 / (tau + 1) * o_GE(S, Q2) + 1.0 / (tau + 1) * o_GM(S, Q2);
}
/* pomeron-fermion vertex matrix of iG_Ppp / iG_PppbarP (:1421-1432,:1465-1481) */
static m44 o_pff_A(v4 psum, const m44 *slash, int mu, int nu) {
  const m44 a = m44_scaler(&og_lo[mu], v4_cov(psum, nu)), b = m44_scaler(&og_lo[nu], v4_cov(psum, mu));
  m44 A = m44_add(&a, &b);
  A = m44_scaler(&A, 0.5);
  if (mu == nu) {
    m44 s = m44_scaler(slash, 0.25);
    s = m44_scaler(&s, o_g[mu][nu]);
    A = m44_sub(&A, &s);
  }
  return A;
}
/* :1441-1488 iG_PppbarP */
static void tp_iG_PppbarP(const ost *S, v4 prime, const cx ubar[4], v4 pt, const m44 *iSF, const cx v[4], v4 p,
                          cx T[4][4][4][4]) {
  const double fl = 3.0 * S->d->tpom.gPNN * o_F1(S, v4_m2(v4_sub(prime, pt)));
  const double fr = 3.0 * S->d->tpom.gPNN * o_F1(S, v4_m2(v4_sub(pt, p)));
  const cx lhsF = {-0.0 * fl, -1.0 * fl}, rhsF = {-0.0 * fr, -1.0 * fr};
  const v4 lsum = v4_add(prime, pt), rsum = v4_add(pt, p);
  const m44 lsl = o_fslash(lsum), rsl = o_fslash(rsum);
  m44 RA[4][4];
  for (int m1 = 0; m1 < 4; ++m1)
    for (int n1 = 0; n1 < 4; ++n1) RA[m1][n1] = o_pff_A(rsum, &rsl, m1, n1);
  for (int m2 = 0; m2 < 4; ++m2)
    for (int n2 = 0; n2 < 4; ++n2) {
      const m44 lA = o_pff_A(lsum, &lsl, m2, n2);
      const m44 lAS = m44_mul(&lA, iSF);
      cx ubarM[4];
      o_vm(ubar, &lAS, ubarM);
      for (int m1 = 0; m1 < 4; ++m1)
        for (int n1 = 0; n1 < 4; ++n1) T[m2][n2][m1][n1] = cx_mul(cx_mul(lhsF, o_vmv(ubarM, &RA[m1][n1], v)), rhsF);
    }
}
static cx tp_chain3x(cx G1[4][4], cx D1[4][4][4][4], cx C[4][4][4][4], int has_iD, cx iD, cx D2[4][4][4][4], cx G2[4][4], int lead_mzi);
static double tensor_me4_2xF(ost *S) {
  og_init();
  const grb_process_desc *d = S->d;
  const v4 pa = S->pbeam1, pb = S->pbeam2, p1 = S->pfinal[1], p2 = S->pfinal[2];
  const v4 p3 = S->node_p4[0], p4 = S->node_p4[1];
  const double M_ = d->nodes[0].mass;
  const v4 pt = v4_sub(v4_sub(pa, p1), p3);
  const v4 pu = v4_add(v4_sub(p4, pa), p1);
  const v4 mp3 = {-p3.x, -p3.y, -p3.z, -p3.e};
  static cx D13[4][4][4][4], D24[4][4][4][4], D14[4][4][4][4], D23[4][4][4][4], Gt[4][4][4][4], Gu[4][4][4][4];
  cx G1a[4][4], G2b[4][4];
  tp_iDP(S, S->ss[1][3], S->t1, D13);
  tp_iDP(S, S->ss[2][4], S->t2, D24);
  tp_iDP(S, S->ss[1][4], S->t1, D14);
  tp_iDP(S, S->ss[2][3], S->t2, D23);
  tp_iG_PppHE(S, p1, pa, G1a);
  tp_iG_PppHE(S, p2, pb, G2b);
  const m44 iSFt = o_iD_F(pt, M_), iSFu = o_iD_F(pu, M_);
  const double fft = pow2(meson_ff(S, v4_m2(pt), pow2(M_))), ffu = pow2(meson_ff(S, v4_m2(pu), pow2(M_))); /* Baryon_FF == Meson_FF */
  /* pomeron legs: X13(a,b) = G1a(m,n) D13(m,n,a,b) etc. */
  cx X13[4][4], X24[4][4], X14[4][4], X23[4][4];
  for (int a = 0; a < 4; ++a)
    for (int b = 0; b < 4; ++b) {
      cx x13 = {0, 0}, x24 = {0, 0}, x14 = {0, 0}, x23 = {0, 0};
      for (int m = 0; m < 4; ++m)
        for (int n = 0; n < 4; ++n) {
          x13 = cx_add(x13, cx_mul(G1a[m][n], D13[m][n][a][b]));
          x14 = cx_add(x14, cx_mul(G1a[m][n], D14[m][n][a][b]));
          x24 = cx_add(x24, cx_mul(D24[a][b][m][n], G2b[m][n]));
          x23 = cx_add(x23, cx_mul(D23[a][b][m][n], G2b[m][n]));
        }
      X13[a][b] = x13; X14[a][b] = x14; X24[a][b] = x24; X23[a][b] = x23;
    }
  cx h[4];
  for (int h3 = 0; h3 < 2; ++h3)
    for (int h4 = 0; h4 < 2; ++h4) {
      cx v3[4], u4[4], ubar4[4];
      o_vHelDirac(p3, h3 == 0 ? -1 : 1, v3);
      o_uHelDirac(p4, h4 == 0 ? -1 : 1, u4);
      o_bar(u4, ubar4);
      tp_iG_PppbarP(S, p4, ubar4, pt, &iSFt, v3, mp3, Gt);
      tp_iG_PppbarP(S, p4, ubar4, pu, &iSFu, v3, mp3, Gu);
      /* FTensor, left to right: [G1a(m,n) D13(m,n,a1,b1)] Gt(a2,b2,a1,b1) -> (a2,b2); D24(a2,b2,m2,n2) -> (m2,n2); G2b(m2,n2).
       * u channel: Gu(a1,b1,a2,b2).  (tp_chain3x of grb_oracle_tres.inc with the middle tensor indexed [a1][b1][a2][b2]) */
      static cx Ct[4][4][4][4];
      for (int a1 = 0; a1 < 4; ++a1)
        for (int b1 = 0; b1 < 4; ++b1)
          for (int a2 = 0; a2 < 4; ++a2)
            for (int b2 = 0; b2 < 4; ++b2) Ct[a1][b1][a2][b2] = Gt[a2][b2][a1][b1];
      cx Mt = tp_chain3x(G1a, D13, Ct, 0, (cx){0, 0}, D24, G2b, 0);
      cx Mu = tp_chain3x(G1a, D14, Gu, 0, (cx){0, 0}, D23, G2b, 0);
      Mt = cx_mulr(Mt, fft);
      Mu = cx_mulr(Mu, ffu);
      h[2 * h3 + h4] = cx_mul((cx){-0.0, -1.0}, cx_add(Mt, Mu));
    }
  S->nhamp = 16;
  double SumAmp2 = 0.0;
  for (int k = 0; k < 16; ++k) {
    S->hamp[k] = h[k % 4];
    SumAmp2 += cx_abs2(S->hamp[k]);
  }
  return SumAmp2 / 4;
}

/* :1342-1363 iG_ypp: full proton-photon-proton vertex between spinors */
static void tp_iG_ypp(const ost *S, v4 prime, v4 p, const cx ubar[4], const cx u[4], cx T[4]) {
  const double t = v4_m2(v4_sub(prime, p));
  const double e = msqrt(0.0072973525693 * 4.0 * O_PI);
  const v4 psum = v4_sub(prime, p);
  const double F1 = o_F1(S, t), F2 = o_F2(S, t);
  for (int mu = 0; mu < 4; ++mu) {
    m44 SUM = m44_zero();
    for (int nu = 0; nu < 4; ++nu) {
      const m44 t1 = m44_scaler(&og_sig_lo[mu][nu], v4_con(psum, nu));
      SUM = m44_add(&SUM, &t1);
    }
    const m44 A = m44_scaler(&og_lo[mu], F1);
    m44 B = m44_scale(&SUM, (cx){0.0, 1.0});
    for (int a = 0; a < 4; ++a) /* MMatrix::operator/(T): element / complex(2 m_p, 0) = component-wise division */
      for (int b = 0; b < 4; ++b) { B.m[a][b].re /= (2 * 0.938272081); B.m[a][b].im /= (2 * 0.938272081); }
    B = m44_scaler(&B, F2);
    m44 G = m44_add(&A, &B);
    G = m44_scale(&G, (cx){-0.0 * e, -1.0 * e});
    T[mu] = o_vmv(ubar, &G, u);
  }
}
/* :1367-1383 iG_yeebary */
static void tp_iG_yeebary(const cx ubar[4], const m44 *iSF, const cx v[4], cx T[4][4]) {
  const double e = msqrt(0.0072973525693 * 4.0 * O_PI);
  const cx vertex = {0.0, e};
  for (int mu = 0; mu < 4; ++mu) {
    const m44 gS = m44_mul(&og_lo[mu], iSF);
    cx lhs[4];
    o_vm(ubar, &gS, lhs);
    for (int nu = 0; nu < 4; ++nu) T[mu][nu] = cx_mul(cx_mul(vertex, o_vmv(lhs, &og_lo[nu], v)), vertex);
  }
}
/* :664-713: lepton pair via two-photon fusion, full QED spinor structure on the proton lines */
static double tensor_me4_yyQED(ost *S) {
  og_init();
  const grb_process_desc *d = S->d;
  const v4 pa = S->pbeam1, pb = S->pbeam2, p1 = S->pfinal[1], p2 = S->pfinal[2];
  const v4 p3 = S->node_p4[0], p4 = S->node_p4[1];
  const double M_ = d->nodes[0].mass;
  const v4 pt = v4_sub(v4_sub(pa, p1), p3);
  const v4 pu = v4_add(v4_sub(p4, pa), p1);
  const m44 iSFt = o_iD_F(pt, M_), iSFu = o_iD_F(pu, M_);
  int k = 0;
  for (int ha = 0; ha < 2; ++ha)
    for (int hb = 0; hb < 2; ++hb) {
      cx ua[4], ub[4], u1[4], u2[4], ubar1[4], ubar2[4], G1[4], G2[4];
      o_uHelDirac(pa, ha == 0 ? -1 : 1, ua);
      o_uHelDirac(pb, hb == 0 ? -1 : 1, ub);
      o_uHelDirac(p1, ha == 0 ? -1 : 1, u1);
      o_uHelDirac(p2, hb == 0 ? -1 : 1, u2);
      o_bar(u1, ubar1);
      o_bar(u2, ubar2);
      tp_iG_ypp(S, p1, pa, ubar1, ua, G1);
      tp_iG_ypp(S, p2, pb, ubar2, ub, G2);
      /* FTensor, left to right (:698-700): [(-zi) iG_1(m1)] iD_1(m1,n1) -> E2(n1); [iyFy_t(n2,n1) + iyFy_u(n1,n2)] E2(n1) -> E3(n2);
       * iD_2(n2,m2) -> E4(m2); iG_2(m2).  iD_y(q2)(u,v) = -zi g[u][v] / q2 (complex / double, MDirac.cc:372-378): diagonal */
      cx E2[4], D2y[4];
      for (int n = 0; n < 4; ++n) {
        const cx d1 = {(-0.0 * o_g[n][n]) / S->t1, (-1.0 * o_g[n][n]) / S->t1};
        D2y[n] = (cx){(-0.0 * o_g[n][n]) / S->t2, (-1.0 * o_g[n][n]) / S->t2};
        E2[n] = cx_mul(cx_mul(G1[n], (cx){-0.0, -1.0}), d1);
      }
      for (int h3 = 0; h3 < 2; ++h3)
        for (int h4 = 0; h4 < 2; ++h4) {
          cx v3[4], u4[4], ubar4[4], Ft[4][4], Fu[4][4];
          o_vHelDirac(p3, h3 == 0 ? -1 : 1, v3);
          o_uHelDirac(p4, h4 == 0 ? -1 : 1, u4);
          o_bar(u4, ubar4);
          tp_iG_yeebary(ubar4, &iSFt, v3, Ft);
          tp_iG_yeebary(ubar4, &iSFu, v3, Fu);
          cx M = {0, 0};
          for (int n2 = 0; n2 < 4; ++n2) {
            cx E3 = {0, 0};
            for (int n1 = 0; n1 < 4; ++n1) E3 = cx_add(E3, cx_mul(cx_add(Ft[n2][n1], Fu[n1][n2]), E2[n1]));
            M = cx_add(M, cx_mul(cx_mul(E3, D2y[n2]), G2[n2]));
          }
          S->hamp[k++] = M;
        }
    }
  S->nhamp = 16;
  double SumAmp2 = 0.0;
  for (int i = 0; i < 16; ++i) SumAmp2 += cx_abs2(S->hamp[i]);
  return SumAmp2 / 4;
}

static double o_factorial(int n) {
  double x = 1;
  for (int i = 1; i <= n; ++i) x *= i;
  return x;
}
/* reference src/MForm.cc:743-827 CBW, types 1-4 */
static cx o_cbw(const ost *S, const grb_resonance *r) {
  const double m2 = S->m2, M0 = r->mass, G = r->width, J = r->spinX2 / 2.0;
  const cx mone = {-1.0, 0.0};
  switch (r->BW) {
    case 1: return cx_div(mone, (cx){m2 - M0 * M0 + 0.0 * M0 * G, 1.0 * M0 * G});
    case 2: return cx_div(mone, (cx){m2 - M0 * M0 + 0.0 * sqrt(m2) * G, 1.0 * sqrt(m2) * G});
    case 3: {
      const double mA = v4_m(S->node_p4[0]), mB = v4_m(S->node_p4[1]);
      const double u = msqrt((m2 - pow2(mA + mB)) * (m2 - pow2(mA - mB))) / (2 * msqrt(m2));
      const double d = msqrt((M0 * M0 - pow2(mA + mB)) * (M0 * M0 - pow2(mA - mB))) / (2 * M0);
      const double B = pow(u / d, 2 * (int)J + 1);
      return cx_div(mone, (cx){m2 - M0 * M0, G * M0 * M0 / msqrt(m2) * B});
    }
    default: {
      const cx den = {m2 - M0 * M0, M0 * G};
      double num;
      if ((int)J == 0) num = -1.0;
      else if ((int)(J * 2) == 1) num = -2.0 * msqrt(m2);
      else if ((int)J == 1) num = -(1.0 - m2 / (M0 * M0));
      else if ((int)(J * 2) == 3) num = -(2.0 / 3.0) * msqrt(m2) * (1.0 - m2 / (M0 * M0));
      else if ((int)J == 2) num = -(7.0 / 6.0 - (4.0 / 3.0) * (m2 / (M0 * M0)) + (2.0 / 3.0) * (m2 * m2) / pow2(pow2(M0)));
      else num = -1.0;
      return cx_div((cx){num, 0.0}, den);
    }
  }
}
/* reference src/MSpin.cc:927-970 WignerSmalld, :978-980 WignerD */
static double o_wigner_d(double theta, double m, double mp, double J) {
  const double factor = msqrt(o_factorial((int)(J + m)) * o_factorial((int)(J - m)) * o_factorial((int)(J + mp)) * o_factorial((int)(J - mp)));
  double s = 0.0;
  for (int k = 0; k <= 2 * J; ++k) {
    double value = 0;
    if (J - mp - k >= 0 && J + m - k >= 0 && k + mp - m >= 0 && k >= 0) {
      const double val1 = pow(-1, mp - m + k) / (o_factorial((int)(J - mp - k)) * o_factorial((int)(J + m - k)) *
                                                  o_factorial((int)(k + mp - m)) * o_factorial(k));
      const double val2 = pow(cos(theta / 2.0), 2 * J + m - mp - 2 * k) * pow(sin(theta / 2.0), mp - m + 2 * k);
      value = val1 * val2;
    }
    s += value;
  }
  return factor * s;
}
static cx o_wigner_D(double theta, double phi, double m, double mp, double J) {
  const double d = o_wigner_d(theta, m, mp, J);
  const double y = (1.0 * mp) * phi;
  return (cx){d * cos(y), d * sin(y)};
}
#define OMAXJ 5 /* 2J+1 <= 5 */
/* reference src/MSpin.cc:631-690 fMatrix: rows = final helicity pairs, cols = initial projections */
static int o_fmatrix(const cx *T, int n1, int n2, double J, double s1, double s2, double theta, double phi, cx f[9][OMAXJ]) {
  const int N = n1 * n2, nJ = (int)(2 * J + 1);
  int i = 0;
  for (int mu = 0; mu < n1; ++mu)
    for (int nu = 0; nu < n2; ++nu) {
      const double lambda = (-s1 + mu) - (-s2 + nu);
      for (int j = 0; j < nJ; ++j) f[i][j] = cx_mul(o_wigner_D(theta, phi, lambda, -J + j, J), T[mu * n2 + nu]);
      ++i;
    }
  return N;
}
/* reference src/MSpin.cc:544-601 GetRhoRotation */
static void o_rho_rotation(const ost *S, int frame, double *th, double *ph) {
  const v4 X = S->pfinal[0];
  const double MX = v4_m(X);
  if (frame == 0) { *th = 0; *ph = 0; }
  else if (frame == 1) { *th = v4_theta(X); *ph = v4_phi(X); }
  else if (frame == 2) {
    const v4 a = lorentz_boost(X, MX, S->pbeam1, -1), b = lorentz_boost(X, MX, S->pbeam2, -1);
    double na = 0.0, nb = 0.0;
    na += pow2(hypot(a.x, 0)); na += pow2(hypot(a.y, 0)); na += pow2(hypot(a.z, 0)); na = msqrt(na);
    nb += pow2(hypot(b.x, 0)); nb += pow2(hypot(b.y, 0)); nb += pow2(hypot(b.z, 0)); nb = msqrt(nb);
    const double zx = a.x / na - b.x / nb, zy = a.y / na - b.y / nb, zz = a.z / na - b.z / nb;
    double nz = 0.0;
    nz += pow2(hypot(zx, 0)); nz += pow2(hypot(zy, 0)); nz += pow2(hypot(zz, 0)); nz = msqrt(nz);
    const v4 bij = {zx / nz, zy / nz, zz / nz, 0.0};
    *th = v4_theta(bij); *ph = v4_phi(bij);
  } else {
    const v4 q = lorentz_boost(X, MX, S->q1, -1);
    *th = v4_theta(q); *ph = v4_phi(q);
  }
}
/* NORM * sqrt(Re Tr[f (D^+ rho D) f^+]) with MMatrix's k-sequential products (MMatrix.h:234-254) */
static double o_spin_trace(const ost *S, const grb_resonance *r, cx f[9][OMAXJ], int N) {
  const double J = r->spinX2 / 2.0;
  const int n = r->spinX2 + 1;
  double th, ph;
  o_rho_rotation(S, r->frame, &th, &ph);
  cx D[OMAXJ][OMAXJ], rho[OMAXJ][OMAXJ], t1[OMAXJ][OMAXJ], rr[OMAXJ][OMAXJ];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      D[i][j] = o_wigner_D(th, ph, -J + i, -J + j, J);
      rho[i][j] = (cx){r->rho_re[i * GRB_RHO_DIM + j], r->rho_im[i * GRB_RHO_DIM + j]};
    }
  for (int i = 0; i < n; ++i)      /* D^+ * rho */
    for (int j = 0; j < n; ++j) {
      cx c = {0, 0};
      for (int k = 0; k < n; ++k) c = cx_add(c, cx_mul(cx_conj(D[k][i]), rho[k][j]));
      t1[i][j] = c;
    }
  for (int i = 0; i < n; ++i)      /* (D^+ rho) * D */
    for (int j = 0; j < n; ++j) {
      cx c = {0, 0};
      for (int k = 0; k < n; ++k) c = cx_add(c, cx_mul(t1[i][k], D[k][j]));
      rr[i][j] = c;
    }
  cx fr[9][OMAXJ];                 /* f * rho_ROT */
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < n; ++j) {
      cx c = {0, 0};
      for (int k = 0; k < n; ++k) c = cx_add(c, cx_mul(f[i][k], rr[k][j]));
      fr[i][j] = c;
    }
  cx tr = {0, 0};                  /* Tr[(f rho) f^+] */
  for (int i = 0; i < N; ++i) {
    cx c = {0, 0};
    for (int k = 0; k < n; ++k) c = cx_add(c, cx_mul(fr[i][k], cx_conj(f[i][k])));
    tr = cx_add(tr, c);
  }
  return msqrt(r->spinX2 + 1.0) * msqrt(tr.re);
}
/* reference src/MSpin.cc:280-400 ProdAmp */
static cx o_prod_amp(const ost *S, const grb_resonance *r) {
  cx A0 = {r->g_re, r->g_im};
  if (!r->spingen) return A0;
  if (r->spinX2 == 0 && r->P == 1) return A0;
  if (r->spinX2 == 0 && r->P == -1) {
    double D = v4_phi(S->pfinal[1]) - v4_phi(S->pfinal[2]);
    while (D >= O_PI) D -= 2.0 * O_PI;
    while (D < -O_PI) D += 2.0 * O_PI;
    const double dphi = fabs(D);
    const double rt = pow(fabs(S->t1), fabs(1.0) / 2.0) * pow(fabs(S->t2), fabs(1.0) / 2.0);
    return cx_scale(cx_scale(A0, rt), sin(dphi));
  }
  double s1 = 0, s2 = 0;
  cx T0[9] = {{1.0, 0.0}};
  int n1 = 1, n2 = 1;
  if (r->spinX2 == 2 && r->P == 1) {
    s1 = 1; s2 = 1; n1 = 3; n2 = 3;
    const double t[9] = {0.000, -0.500, 0.000, 0.500, 0.000, -0.500, 0.000, 0.500, 0.000};
    for (int i = 0; i < 9; ++i) T0[i] = (cx){t[i], 0.0};
  }
  if (r->spinX2 == 2 && r->P == -1) {
    s1 = 0; s2 = 1; n1 = 1; n2 = 3;
    T0[0] = (cx){-0.697, 0}; T0[1] = (cx){-0.169, 0}; T0[2] = (cx){0.697, 0};
  }
  const v4 q1 = lorentz_boost(S->pfinal[0], v4_m(S->pfinal[0]), S->q1, -1);
  cx f0[9][OMAXJ];
  const int N = o_fmatrix(T0, n1, n2, r->spinX2 / 2.0, s1, s2, v4_theta(q1), v4_phi(q1), f0);
  return cx_scale(A0, o_spin_trace(S, r, f0, N));
}
/* reference src/MSpin.cc:406-483 DecayAmp (two final-state daughters, no cascades) */
static cx o_decay_amp(const ost *S, const grb_resonance *r) {
  const double A0 = r->g_decay;
  if (!r->spindec) return (cx){A0, 0.0};
  if (S->d->n_legs != 2) return (cx){A0, 0.0};
  if (r->spinX2 == 0) return cx_scale((cx){r->T_re[0], r->T_im[0]}, A0);
  const double s1 = S->d->nodes[0].spinX2 / 2.0, s2 = S->d->nodes[1].spinX2 / 2.0;
  const v4 bd = lorentz_boost(S->pfinal[0], v4_m(S->pfinal[0]), S->node_p4[0], -1);
  cx T[9], f[9][OMAXJ];
  for (int i = 0; i < r->T_rows * r->T_cols; ++i) T[i] = (cx){r->T_re[i], r->T_im[i]};
  const int N = o_fmatrix(T, r->T_rows, r->T_cols, r->spinX2 / 2.0, s1, s2, v4_theta(bd), v4_phi(bd), f);
  return (cx){A0 * o_spin_trace(S, r, f, N), 0.0};
}
/* reference src/MForm.cc:664-684 CohFlux (alpha_QED() = alpha_0, MQED.h:18-21) */
static double o_coh_flux(const ost *S, double xi, double t, double pt) {
  const double mp = 0.938272081;
  const double pt2 = pow2(pt), xi2 = pow2(xi), mp2 = pow2(mp), Q2 = fabs(t);
  const double PART1 = (4.0 * pow2(mp) * pow2(o_GE(S, Q2)) + Q2 * pow2(o_GM(S, Q2))) / (4.0 * pow2(mp) + Q2);
  const double PART2 = pow2(o_GM(S, Q2));
  const double DELTA = pt2 / (pt2 + xi2 * mp2);
  double f = 0.0072973525693 / O_PI * ((1.0 - xi) * pow2(DELTA) * PART1 + (xi2 / 4.0) * DELTA * PART2);
  f /= xi;
  f /= pt2;
  f *= 16.0 * (O_PI * O_PI);
  return f;
}
/* reference src/MForm.cc:463-518 F2xQ2 (Donnachie-Landshoff or CKMT), F1xQ2 (Callan-Gross) */
static double o_f2xq2(const ost *S, double xbj, double Q2) {
  if (S->d->soft.f2_model == 1) {
    const double A = 0.324, B = 0.098, DELTA_P = 0.0808, DELTA_R = 0.5475, a = 0.561991692786383, b = 0.011133;
    return A * pow(xbj, -DELTA_P) * pow(Q2 / (Q2 + a), 1 + DELTA_P) + B * pow(xbj, 1 - DELTA_R) * pow(Q2 / (Q2 + b), DELTA_R);
  }
  const double A = 0.1502, B_u = 1.2064, B_d = 0.1798, alpha_R = 0.4150, DELTA_0 = 0.0800;
  const double a = 0.2631, b = 0.6452, c = 3.5489, d = 1.1170;
  const double n_Q2     = (3.0 / 2.0) * (1 + Q2 / (Q2 + c));
  const double DELTA_Q2 = DELTA_0 * (1 + (2 * Q2) / (Q2 + d));
  const double C1 = pow(Q2 / (Q2 + a), 1.0 + DELTA_Q2);
  const double C2 = pow(Q2 / (Q2 + b), alpha_R);
  return A * pow(xbj, -DELTA_Q2) * pow(1 - xbj, n_Q2 + 4.0) * C1 +
         pow(xbj, 1.0 - alpha_R) * (B_u * pow(1 - xbj, n_Q2) + B_d * pow(1 - xbj, n_Q2 + 1.0)) * C2;
}
static double o_f1xq2(const ost *S, double xbj, double Q2) { return o_f2xq2(S, xbj, Q2) / (2.0 * xbj); }
/* reference src/MForm.cc:700-719 IncohFlux: kt-unintegrated photon flux of a dissociating proton */
static double o_incoh_flux(const ost *S, double xi, double t, double pt, double M2) {
  const double mp = 0.938272081, mp2 = pow2(mp);
  const double pt2 = pow2(pt), xi2 = pow2(xi), Q2 = fabs(t);
  const double xbj   = Q2 / (Q2 + M2 - mp2);
  const double DELTA = pt2 / (pt2 + xi * (M2 - mp2) + xi2 * mp2);
  double f = 0.0072973525693 / O_PI *
             ((1.0 - xi) * pow2(DELTA) * o_f2xq2(S, xbj, Q2) / (Q2 + M2 - mp2) +
              (xi2 / (4.0 * pow2(xbj))) * DELTA * 2.0 * xbj * o_f1xq2(S, xbj, Q2) / (Q2 + M2 - mp2));
  f /= xi;
  f /= pt2;
  f *= 16.0 * (O_PI * O_PI);
  return f;
}
/* reference src/MRegge.cc:852-882 PhotoProp (elastic leg) */
static cx o_photo_prop(const ost *S, double s, double t, double m2, int excite, double M2_forward) {
  const grb_param_regge *R = &S->d->regge;
  const double B0 = pow(fmax(pow2(0.7), m2) / 2.0, -O_PI / 2.0) + 3 * O_PI / 2.0;
  const double W02 = pow2(1.0);
  const double alpha = R->a0[0] + R->ap[0] * t;
  const double y = ((-1.0 * O_PI) / 2.0 * R->ap[0]) * t;
  const cx eta = cx_mul(S->regge_eta0[0], (cx){cos(y), sin(y)});
  /* proton / dissociative system vertex, MRegge.cc:876-880 */
  const double FF = excite ? msqrt(S->d->soft.g3P * S->d->soft.gN_P) * s3finel(t, M2_forward) : S->d->soft.gN_P * exp(B0 * t / 2.0);
  return cx_scale(cx_scale(eta, pow(s / W02, alpha)), FF);
}
/* reference src/MRegge.cc:677-706 ME3 */
/* reference src/MRegge.cc:470-540 ME6: the 2 -> 6 ladder summed over the 24 orderings (a, b, c, d) of the four central
 * legs, in the order of math::Permutations (MMath.h:662-720: lexicographic successor) */
static cx regge_me6(ost *S) {
  const grb_process_desc *d = S->d;
  const double M2_A = pow2(d->nodes[0].mass), M2_B = pow2(d->nodes[2].mass);
  double FF_A, FF_B;
  pom_pom_proton_vertex(S, &FF_A, &FF_B);
  const double gpp_P = d->regge.c[0] / d->soft.gN_P;
  cx  A = {0, 0};
  int set[4] = {3, 4, 5, 6};
  for (;;) {
    const int a = set[0], b = set[1], c = set[2], dd = set[3];
    const double tt_ab = S->tt_1[a], tt_bc = S->tt_xy[a][b], tt_cd = S->tt_2[dd];
    cx sub = prop_only(S, S->ss[1][a], S->t1);
    sub = cx_scale(sub, FF_A);
    sub = cx_scale(sub, meson_ff(S, tt_ab, M2_A));
    sub = cx_scale(sub, gpp_P);
    sub = cx_scale(sub, meson_prop(tt_ab, M2_A));
    sub = cx_scale(sub, meson_ff(S, tt_bc, M2_A));
    sub = cx_scale(sub, gpp_P);
    sub = cx_mul(sub, prop_only(S, S->ss[b][c], tt_bc));
    sub = cx_scale(sub, meson_ff(S, tt_bc, M2_B));
    sub = cx_scale(sub, gpp_P);
    sub = cx_scale(sub, meson_prop(tt_cd, M2_B));
    sub = cx_scale(sub, meson_ff(S, tt_cd, M2_B));
    sub = cx_scale(sub, gpp_P);
    sub = cx_mul(sub, prop_only(S, S->ss[2][dd], S->t2));
    sub = cx_scale(sub, FF_B);
    A = cx_add(A, sub);
    /* next permutation in lexicographic order */
    int j = -1;
    for (int i = 0; i < 3; ++i)
      if (set[i] < set[i + 1]) j = i;
    if (j < 0) break;
    int l = j + 1;
    for (int i = j + 1; i < 4; ++i)
      if (set[j] < set[i]) l = i;
    int t = set[j]; set[j] = set[l]; set[l] = t;
    for (int lo = j + 1, hi = 3; lo < hi; ++lo, --hi) { t = set[lo]; set[lo] = set[hi]; set[hi] = t; }
  }
  S->hamp[0] = A;
  S->nhamp   = 1;
  return A;
}
/* reference src/MRegge.cc:560-633 ME8: the 2 -> 8 ladder summed over the 720 orderings (a, b, c, d, e, f) of the six central
 * legs in the order of math::Permutations; the scalars are the ones the reference reads (tt_cd = tt_xy[b][c] and
 * tt_de = tt_xy[c][d] are two-leg sums, and there is no form factor on tt_cd: restated as written) */
static cx regge_me8(ost *S) {
  const grb_process_desc *d = S->d;
  const double M2_A = pow2(d->nodes[0].mass), M2_B = pow2(d->nodes[2].mass), M2_C = pow2(d->nodes[4].mass);
  double FF_A, FF_B;
  pom_pom_proton_vertex(S, &FF_A, &FF_B);
  const double gpp_P = d->regge.c[0] / d->soft.gN_P;
  cx  A = {0, 0};
  int set[6] = {3, 4, 5, 6, 7, 8};
  for (;;) {
    const int a = set[0], b = set[1], c = set[2], dd = set[3], e = set[4], f = set[5];
    const double tt_ab = S->tt_1[a], tt_bc = S->tt_xy[a][b], tt_cd = S->tt_xy[b][c], tt_de = S->tt_xy[c][dd], tt_ef = S->tt_2[f];
    cx sub = prop_only(S, S->ss[1][a], S->t1);
    sub = cx_scale(sub, FF_A);
    sub = cx_scale(sub, meson_ff(S, tt_ab, M2_A));
    sub = cx_scale(sub, gpp_P);
    sub = cx_scale(sub, meson_prop(tt_ab, M2_A));
    sub = cx_scale(sub, meson_ff(S, tt_bc, M2_A));
    sub = cx_scale(sub, gpp_P);
    sub = cx_mul(sub, prop_only(S, S->ss[b][c], tt_bc));
    sub = cx_scale(sub, meson_ff(S, tt_bc, M2_B));
    sub = cx_scale(sub, gpp_P);
    sub = cx_scale(sub, meson_prop(tt_cd, M2_B));
    sub = cx_scale(sub, meson_ff(S, tt_de, M2_B));
    sub = cx_scale(sub, gpp_P);
    sub = cx_mul(sub, prop_only(S, S->ss[dd][e], tt_de));
    sub = cx_scale(sub, meson_ff(S, tt_de, M2_C));
    sub = cx_scale(sub, gpp_P);
    sub = cx_scale(sub, meson_prop(tt_ef, M2_C));
    sub = cx_scale(sub, meson_ff(S, tt_ef, M2_C));
    sub = cx_scale(sub, gpp_P);
    sub = cx_mul(sub, prop_only(S, S->ss[2][f], S->t2));
    sub = cx_scale(sub, FF_B);
    A = cx_add(A, sub);
    int j = -1;
    for (int i = 0; i < 5; ++i)
      if (set[i] < set[i + 1]) j = i;
    if (j < 0) break;
    int l = j + 1;
    for (int i = j + 1; i < 6; ++i)
      if (set[j] < set[i]) l = i;
    int t = set[j]; set[j] = set[l]; set[l] = t;
    for (int lo = j + 1, hi = 5; lo < hi; ++lo, --hi) { t = set[lo]; set[lo] = set[hi]; set[hi] = t; }
  }
  S->hamp[0] = A;
  S->nhamp   = 1;
  return A;
}
static cx regge_me3(ost *S, const grb_resonance *r) {
  double FF_A, FF_B;
  pom_pom_proton_vertex(S, &FF_A, &FF_B);
  cx A = cx_scale(prop_only(S, S->s1, S->t1), 2.0);
  A = cx_scale(A, FF_A);
  A = cx_mul(A, o_cbw(S, r));
  A = cx_scale(A, exp(-pow2(S->m2 - pow2(r->mass)) / pow2(pow2(r->g_FF))));
  A = cx_mul(A, prop_only(S, S->s2, S->t2));
  A = cx_scale(A, FF_B);
  const cx A_spin = cx_mul(o_prod_amp(S, r), o_decay_amp(S, r));
  const double V = pow(1.0 / S->m2, S->d->regge.omega);
  return cx_scale(cx_mul(A, A_spin), V);
}
/* reference src/MRegge.cc:781-828 PhotoME3 (pp initial state, elastic legs) */
static cx regge_photo_me3(ost *S, const grb_resonance *r) {
  const cx common = cx_scale(o_cbw(S, r), exp(-pow2(S->m2 - pow2(r->mass)) / pow2(pow2(r->g_FF))));
  const double fl1 = S->excite1 ? o_incoh_flux(S, S->x1, S->t1, S->qt1, v4_m2(S->pfinal[1])) : o_coh_flux(S, S->x1, S->t1, S->qt1);
  const double fl2 = S->excite2 ? o_incoh_flux(S, S->x2, S->t2, S->qt2, v4_m2(S->pfinal[2])) : o_coh_flux(S, S->x2, S->t2, S->qt2);
  const double gf1 = msqrt(fl1 / S->x1);
  const double gf2 = msqrt(fl2 / S->x2);
  const cx A_1 = cx_mul(cx_scale(common, gf2), o_photo_prop(S, S->s1, S->t1, pow2(r->mass), S->excite1, v4_m2(S->pfinal[1])));
  const cx A_2 = cx_mul(cx_scale(common, gf1), o_photo_prop(S, S->s2, S->t2, pow2(r->mass), S->excite2, v4_m2(S->pfinal[2])));
  const cx A_prod = cx_add(A_1, A_2);
  const cx A_spin = cx_mul(o_prod_amp(S, r), o_decay_amp(S, r));
  return cx_mul(A_prod, A_spin);
}
/* reference PROC_9 / PROC_17 (include/Graniitti/MSubProc.h:233-262,399-436): coherent sum in std::map order */
static double res_amp2(ost *S, int with_continuum) {
  cx A = {0.0, 0.0};
  if (with_continuum) A = regge_me4(S, 1.0);
  for (int i = 0; i < S->d->n_res; ++i) {
    const grb_resonance *r = &S->d->res[i];
    const int J = (int)(r->spinX2 / 2.0);
    A = cx_add(A, (J == 1 && r->P == -1) ? regge_photo_me3(S, r) : regge_me3(S, r));
  }
  S->hamp[0] = A;
  S->nhamp = 1;
  return cx_abs2(A);
}

/* ---- gamma gamma -> l+ l- : MadGraph5/HELAS helicity amplitudes + kt-EPA fluxes ------------------------ */
/* reference src/Amplitude/AMP_MG5_yy_ll.cc:41-255 (CalcAmp2, calculate_wavefunctions),
 * src/Amplitude/HelAmps_sm.cc (ixxxxx :265, oxxxxx :140, vxxxxx :207, FFV1_0 :678, FFV1_1 :930, FFV1_2 :435),
 * include/Graniitti/MKinematics.h:799-907 (OffShell2LightCone), src/MFlux.cc:22-44 (ApplyktEPAfluxes).
 * The ALOHA vertex routines are restated in light-cone combinations V+- = V0 +- V3, Vc = V1 + i V2,
 * Vd = V1 - i V2 (same for the real propagator momentum P), which is what the generated expressions reduce to. */
static inline cx cx_sub(cx a, cx b) { return (cx){a.re - b.re, a.im - b.im}; }
static inline cx cx_i(cx a) { return (cx){-a.im, a.re}; }          /* i * a */
static inline cx cx_r(double r) { return (cx){r, 0.0}; }
static double o_sgn(double a, double b) { return (b < 0) ? -fabs(a) : fabs(a); }

/* p = (E,px,py,pz) MadGraph convention; w[0..1] momentum, w[2..5] wavefunction */
static void o_vxxxxx(const double p[4], int nhel, int nsv, cx vc[6]) { /* massless vector */
  const double sqh = sqrt(0.5), hel = (double)nhel;
  vc[0] = (cx){p[0] * nsv, p[3] * nsv};
  vc[1] = (cx){p[1] * nsv, p[2] * nsv};
  const double pp = p[0], pt = sqrt((p[1] * p[1]) + (p[2] * p[2]));
  vc[2] = (cx){0.0, 0.0};
  vc[5] = (cx){hel * pt / pp * sqh, 0.0};
  if (pt != 0.0) {
    const double pzpt = p[3] / (pp * pt) * sqh * hel;
    vc[3] = (cx){-p[1] * pzpt, -nsv * p[2] / pt * sqh};
    vc[4] = (cx){-p[2] * pzpt, nsv * p[1] / pt * sqh};
  } else {
    vc[3] = (cx){-hel * sqh, 0.0};
    vc[4] = (cx){0.0, nsv * o_sgn(sqh, p[3])};
  }
}
static void o_ixxxxx(const double p[4], double fmass, int nhel, int nsf, cx fi[6]) {
  cx chi[2];
  fi[0] = (cx){-p[0] * nsf, -p[3] * nsf};
  fi[1] = (cx){-p[1] * nsf, -p[2] * nsf};
  const int nh = nhel * nsf;
  if (fmass != 0.0) {
    const double pp = fmin(p[0], sqrt(p[1] * p[1] + p[2] * p[2] + p[3] * p[3]));
    if (pp == 0.0) {
      double sqm[2];
      sqm[0] = sqrt(fabs(fmass)); sqm[1] = o_sgn(sqm[0], fmass);
      const int ip = (1 + nh) / 2, im = (1 - nh) / 2;
      fi[2] = cx_r(ip * sqm[ip]); fi[3] = cx_r(im * nsf * sqm[ip]); fi[4] = cx_r(ip * nsf * sqm[im]); fi[5] = cx_r(im * sqm[im]);
    } else {
      double sf[2], omega[2], sfomega[2];
      sf[0] = (1 + nsf + (1 - nsf) * nh) * 0.5; sf[1] = (1 + nsf - (1 - nsf) * nh) * 0.5;
      omega[0] = sqrt(p[0] + pp); omega[1] = fmass / omega[0];
      const int ip = (1 + nh) / 2, im = (1 - nh) / 2;
      sfomega[0] = sf[0] * omega[ip]; sfomega[1] = sf[1] * omega[im];
      const double pp3 = fmax(pp + p[3], 0.0);
      chi[0] = (cx){sqrt(pp3 * 0.5 / pp), 0};
      if (pp3 == 0.0) chi[1] = (cx){-nh, 0};
      else { const double dn = sqrt(2.0 * pp * pp3); chi[1] = (cx){nh * p[1] / dn, p[2] / dn}; }
      fi[2] = cx_scale(chi[im], sfomega[0]); fi[3] = cx_scale(chi[ip], sfomega[0]);
      fi[4] = cx_scale(chi[im], sfomega[1]); fi[5] = cx_scale(chi[ip], sfomega[1]);
    }
  } else {
    double sqp0p3;
    if (p[1] == 0.0 && p[2] == 0.0 && p[3] < 0.0) sqp0p3 = 0.0; else sqp0p3 = sqrt(fmax(p[0] + p[3], 0.0)) * nsf;
    chi[0] = (cx){sqp0p3, 0.0};
    if (sqp0p3 == 0.0) chi[1] = (cx){-nhel * sqrt(2.0 * p[0]), 0.0};
    else chi[1] = (cx){nh * p[1] / sqp0p3, p[2] / sqp0p3};
    if (nh == 1) { fi[2] = cx_r(0); fi[3] = cx_r(0); fi[4] = chi[0]; fi[5] = chi[1]; }
    else { fi[2] = chi[1]; fi[3] = chi[0]; fi[4] = cx_r(0); fi[5] = cx_r(0); }
  }
}
static void o_oxxxxx(const double p[4], double fmass, int nhel, int nsf, cx fo[6]) {
  cx chi[2];
  fo[0] = (cx){p[0] * nsf, p[3] * nsf};
  fo[1] = (cx){p[1] * nsf, p[2] * nsf};
  const int nh = nhel * nsf;
  if (fmass != 0.000) {
    double pp = fmin(p[0], sqrt((p[1] * p[1]) + (p[2] * p[2]) + (p[3] * p[3])));
    if (pp == 0.000) {
      double sqm[2];
      sqm[0] = sqrt(fabs(fmass)); sqm[1] = o_sgn(sqm[0], fmass);
      const int ip = -((1 - nh) / 2) * nhel, im = (1 + nh) / 2 * nhel;
      fo[2] = cx_r(im * sqm[abs(ip)]); fo[3] = cx_r(ip * nsf * sqm[abs(ip)]);
      fo[4] = cx_r(im * nsf * sqm[abs(im)]); fo[5] = cx_r(ip * sqm[abs(im)]);
    } else {
      double sf[2], omega[2], sfomeg[2];
      sf[0] = (double)(1 + nsf + (1 - nsf) * nh) * 0.5; sf[1] = (double)(1 + nsf - (1 - nsf) * nh) * 0.5;
      omega[0] = sqrt(p[0] + pp); omega[1] = fmass / omega[0];
      const int ip = (1 + nh) / 2, im = (1 - nh) / 2;
      sfomeg[0] = sf[0] * omega[ip]; sfomeg[1] = sf[1] * omega[im];
      const double pp3 = fmax(pp + p[3], 0.00);
      chi[0] = (cx){sqrt(pp3 * 0.5 / pp), 0.00};
      if (pp3 == 0.00) chi[1] = (cx){-nh, 0.00};
      else { const double dn = sqrt(2.0 * pp * pp3); chi[1] = (cx){nh * p[1] / dn, -p[2] / dn}; }
      fo[2] = cx_scale(chi[im], sfomeg[1]); fo[3] = cx_scale(chi[ip], sfomeg[1]);
      fo[4] = cx_scale(chi[im], sfomeg[0]); fo[5] = cx_scale(chi[ip], sfomeg[0]);
    }
  } else {
    double sqp0p3;
    if ((p[1] == 0.00) && (p[2] == 0.00) && (p[3] < 0.00)) sqp0p3 = 0.00; else sqp0p3 = sqrt(fmax(p[0] + p[3], 0.00)) * nsf;
    chi[0] = (cx){sqp0p3, 0.00};
    if (sqp0p3 == 0.000) chi[1] = cx_scale((cx){-nhel, 0.00}, sqrt(2.0 * p[0]));
    else chi[1] = (cx){nh * p[1] / sqp0p3, -p[2] / sqp0p3};
    if (nh == 1) { fo[2] = chi[0]; fo[3] = chi[1]; fo[4] = cx_r(0); fo[5] = cx_r(0); }
    else { fo[2] = cx_r(0); fo[3] = cx_r(0); fo[4] = chi[1]; fo[5] = chi[0]; }
  }
}
typedef struct { cx p, m, c, d; } lc4; /* V+ , V- , V1+iV2 , V1-iV2 */
static lc4 o_lc(const cx V[6]) {
  lc4 r;
  r.p = cx_add(V[2], V[5]); r.m = cx_sub(V[2], V[5]);
  r.c = cx_add(V[3], cx_i(V[4])); r.d = cx_sub(V[3], cx_i(V[4]));
  return r;
}
/* FFV1_0: COUP * (-i) * [ F1 (gamma.V) F2 ] */
static cx o_ffv1_0(const cx F1[6], const cx F2[6], const cx V3[6], cx COUP) {
  const lc4 v = o_lc(V3);
  cx t = cx_mul(F1[2], cx_add(cx_mul(F2[4], v.p), cx_mul(F2[5], v.c)));
  t = cx_add(t, cx_mul(F1[3], cx_add(cx_mul(F2[4], v.d), cx_mul(F2[5], v.m))));
  t = cx_add(t, cx_mul(F1[4], cx_sub(cx_mul(F2[2], v.m), cx_mul(F2[3], v.c))));
  t = cx_add(t, cx_mul(F1[5], cx_sub(cx_mul(F2[3], v.p), cx_mul(F2[2], v.d))));
  const cx mi = {0.0, -1.0};
  return cx_mul(cx_mul(COUP, mi), t);
}
/* FFV1_2 with M = W = 0: off-shell incoming-flow spinor F2 from (F1, V3) */
static void o_ffv1_2(const cx F1[6], const cx V3[6], cx COUP, cx F2[6]) {
  F2[0] = cx_add(F1[0], V3[0]); F2[1] = cx_add(F1[1], V3[1]);
  const double P0 = -F2[0].re, P1 = -F2[1].re, P2 = -F2[1].im, P3 = -F2[0].im;
  const cx denom = cx_div(COUP, cx_r((P0 * P0) - (P1 * P1) - (P2 * P2) - (P3 * P3)));
  const lc4 v = o_lc(V3);
  const double Pp = P0 + P3, Pm = P0 - P3;
  const cx Pc = {P1, P2}, Pd = {P1, -P2};
  const cx a = cx_add(cx_mul(F1[2], v.p), cx_mul(F1[3], v.d)), b = cx_add(cx_mul(F1[2], v.c), cx_mul(F1[3], v.m));
  const cx c = cx_sub(cx_mul(F1[4], v.m), cx_mul(F1[5], v.d)), e = cx_sub(cx_mul(F1[4], v.c), cx_mul(F1[5], v.p));
  const cx id = cx_i(denom);
  F2[2] = cx_mul(id, cx_sub(cx_scale(a, Pm), cx_mul(Pd, b)));
  F2[3] = cx_mul(id, cx_sub(cx_scale(b, Pp), cx_mul(Pc, a)));
  F2[4] = cx_mul(id, cx_sub(cx_scale(c, Pp), cx_mul(Pd, e)));
  F2[5] = cx_mul(id, cx_sub(cx_mul(Pc, c), cx_scale(e, Pm)));
}
/* FFV1_1 with M = W = 0: off-shell outgoing-flow spinor F1 from (F2, V3) */
static void o_ffv1_1(const cx F2[6], const cx V3[6], cx COUP, cx F1[6]) {
  F1[0] = cx_add(F2[0], V3[0]); F1[1] = cx_add(F2[1], V3[1]);
  const double P0 = -F1[0].re, P1 = -F1[1].re, P2 = -F1[1].im, P3 = -F1[0].im;
  const cx denom = cx_div(COUP, cx_r((P0 * P0) - (P1 * P1) - (P2 * P2) - (P3 * P3)));
  const lc4 v = o_lc(V3);
  const double Pp = P0 + P3, Pm = P0 - P3;
  const cx Pc = {P1, P2}, Pd = {P1, -P2};
  const cx a = cx_sub(cx_mul(F2[2], v.m), cx_mul(F2[3], v.c)), b = cx_sub(cx_mul(F2[2], v.d), cx_mul(F2[3], v.p));
  const cx c = cx_add(cx_mul(F2[4], v.p), cx_mul(F2[5], v.c)), e = cx_add(cx_mul(F2[4], v.d), cx_mul(F2[5], v.m));
  const cx id = cx_i(denom);
  F1[2] = cx_mul(id, cx_sub(cx_mul(Pc, b), cx_scale(a, Pp)));
  F1[3] = cx_mul(id, cx_sub(cx_scale(b, Pm), cx_mul(Pd, a)));
  F1[4] = cx_mul(id, cx_sub(cx_mul(Pc, e), cx_scale(c, Pm)));
  F1[5] = cx_mul(id, cx_sub(cx_mul(Pd, c), cx_scale(e, Pp)));
}
/* reference include/Graniitti/MKinematics.h:799-907 OffShell2LightCone, N = 2 final states */
static void o_offshell2lightcone(v4 *p1, v4 *p2, v4 p[2]) {
  const int N = 2;
  p1->e = v4_p3mod(*p1);
  p2->e = v4_p3mod(*p2);
  const v4 q = v4_add(*p1, *p2);
  double m[2];
  for (int i = 0; i < N; ++i) m[i] = v4_m(p[i]);
  int iter = 0;
  while (1) {
    double Esum = 0.0;
    for (int i = 0; i < N; ++i) Esum += p[i].e;
    const double dE = Esum - q.e;
    for (int i = 0; i < N; ++i) {
      const double E = p[i].e - dE / N;
      const double a = msqrt(E * E - m[i] * m[i]) / v4_p3mod(p[i]);
      p[i] = (v4){a * p[i].x, a * p[i].y, a * p[i].z, E};
    }
    v4 sum = {0, 0, 0, 0};
    for (int i = 0; i < N; ++i) sum = v4_add(sum, p[i]);
    const v4 D = v4_sub(sum, q);
    for (int i = 0; i < N; ++i) {
      const double dx = D.x * (1.0 / N), dy = D.y * (1.0 / N), dz = D.z * (1.0 / N);
      p[i].x = p[i].x - dx; p[i].y = p[i].y - dy; p[i].z = p[i].z - dz;
      p[i].e = msqrt(v4_p3mod2(p[i]) + m[i] * m[i]);
    }
    ++iter;
    sum = (v4){0, 0, 0, 0};
    for (int i = 0; i < N; ++i) sum = v4_add(sum, p[i]);
    const double IVM = fabs(v4_m2(v4_sub(q, sum)));
    if (iter > 15 || IVM < 1e-10) break;
  }
}
/* reference PROC_3: MGamma::yyffbar (lepton pair) -> AMP_MG5_yy_ll::CalcAmp2, then flux::ApplyktEPAfluxes */
static double yy_con_amp2(ost *S) {
  static const int HEL[16][4] = {{-1, -1, -1, -1}, {-1, -1, -1, 1}, {-1, -1, 1, -1}, {-1, -1, 1, 1}, {-1, 1, -1, -1},
                                 {-1, 1, -1, 1},  {-1, 1, 1, -1},  {-1, 1, 1, 1},  {1, -1, -1, -1}, {1, -1, -1, 1},
                                 {1, -1, 1, -1},  {1, -1, 1, 1},   {1, 1, -1, -1}, {1, 1, -1, 1},   {1, 1, 1, -1}, {1, 1, 1, 1}};
  v4 pf[2] = {S->node_p4[0], S->node_p4[1]};
  const double m3 = v4_m(pf[0]), m4 = v4_m(pf[1]);
  v4 q1 = S->q1, q2 = S->q2;
  o_offshell2lightcone(&q1, &q2, pf);
  const double P1[4] = {q1.e, q1.x, q1.y, q1.z}, P2[4] = {q2.e, q2.x, q2.y, q2.z};
  const double P3[4] = {pf[0].e, pf[0].x, pf[0].y, pf[0].z}, P4[4] = {pf[1].e, pf[1].x, pf[1].y, pf[1].z};
  /* GC_3 = -(ee * i), ee = 2 sqrt(1/137.03599908) sqrt(pi)  (Parameters_sm.cc:93-101, QED_alpha = ZERO); otherwise the
   * MG5 model value mdl_ee = 2 sqrt(aEW) sqrt(pi), aEW = 1 / aEWM1, aEWM1 = 1.325070e+02 (Parameters_sm.cc:31,45-49,
   * MG5cards/param_card_yy_ll.dat) */
  const double aEWM1 = 1.325070e+02, aEW = 1. / aEWM1;
  const double ee = S->d->qed_alpha == 0 ? 2. * sqrt(1.0 / 137.03599908) * sqrt(M_PI) : 2 * sqrt(aEW) * sqrt(M_PI);
  const cx GC3 = {-(ee * 0.0), -(ee * 1.0)};
  double amp2 = 0.0;
  for (int ih = 0; ih < 16; ++ih) {
    cx w0[6], w1[6], w2[6], w3[6], w4[6], w5[6];
    o_vxxxxx(P1, HEL[ih][0], -1, w0);
    o_vxxxxx(P2, HEL[ih][1], -1, w1);
    o_ixxxxx(P3, m3, HEL[ih][2], -1, w2);
    o_oxxxxx(P4, m4, HEL[ih][3], +1, w3);
    o_ffv1_2(w2, w0, GC3, w4);
    o_ffv1_1(w3, w0, GC3, w5);
    const cx a0 = o_ffv1_0(w4, w3, w1, GC3), a1 = o_ffv1_0(w2, w5, w1, GC3);
    S->hamp[ih] = cx_add(cx_add((cx){0, 0}, a0), a1);
  }
  for (int ih = 0; ih < 16; ++ih) amp2 += cx_abs2(S->hamp[ih]);
  amp2 /= 4;
  S->nhamp = 16;
  /* kt-EPA fluxes, src/MFlux.cc:22-44 */
  const double f1 = S->excite1 ? o_incoh_flux(S, S->x1, S->t1, S->qt1, v4_m2(S->pfinal[1])) : o_coh_flux(S, S->x1, S->t1, S->qt1);
  const double f2 = S->excite2 ? o_incoh_flux(S, S->x2, S->t2, S->qt2, v4_m2(S->pfinal[2])) : o_coh_flux(S, S->x2, S->t2, S->qt2);
  const double fluxes = f1 * f2 * (1.0 / (S->x1 * S->x2));
  const double tot = fluxes * amp2;
  const double sf = msqrt(fluxes);
  for (int ih = 0; ih < 16; ++ih) S->hamp[ih] = cx_scale(S->hamp[ih], sf);
  return tot;
}

/* reference seam B: MSubProc::GetBareAmplitude2 -> PROC_n::Amp2 (include/Graniitti/MSubProc.h) */
#include "grb_oracle_tres.inc"

static double bare_amp2(ost *S) {
  switch (S->d->proc) {
    case GRB_PROC_PP_RESTENSOR: return tensor_me3(S);
    case GRB_PROC_PP_RES_CONTENSOR: return tensor_res_con(S);
    case GRB_PROC_PP_CON: return S->d->n_legs == 6 ? cx_abs2(regge_me8(S)) : (S->d->n_legs == 4 ? cx_abs2(regge_me6(S)) : cx_abs2(regge_me4(S, 1.0)));
    case GRB_PROC_PP_CON_MINUS: return cx_abs2(regge_me4(S, -1.0));
    case GRB_PROC_YY_CON: {
      const int apdg = abs(S->d->nodes[0].pdg);
      if (!(apdg == 11 || apdg == 13 || apdg == 15) || S->d->nodes[1].pdg != -S->d->nodes[0].pdg) {
        S->nhamp = 0;
        return NAN; /* only charged lepton pairs are restated */
      }
      return yy_con_amp2(S);
    }
    case GRB_PROC_PP_RES: return res_amp2(S, 0);
    case GRB_PROC_PP_RES_CON: return res_amp2(S, 1);
    case GRB_PROC_YY_QED: {
      const int apdg = abs(S->d->nodes[0].pdg);
      if (!(apdg == 11 || apdg == 13 || apdg == 15) || S->d->nodes[1].pdg != -S->d->nodes[0].pdg) {
        S->nhamp = 0;
        return NAN; /* quark pairs need fragmentation downstream: not restated */
      }
      return tensor_me4_yyQED(S);
    }
    case GRB_PROC_PP_CONTENSOR: {
      if (S->d->nodes[0].spinX2 == 2 && S->d->nodes[1].spinX2 == 2) return tensor_me4_2xV(S);
      if (S->d->nodes[0].spinX2 == 1 && S->d->nodes[1].spinX2 == 1) return tensor_me4_2xF(S);
      if (g_oracle_extended) {
        double re, im;
        const double a2 = tensor_me4_extended(S, &re, &im);
        for (int h = 0; h < 4; ++h) S->hamp[h] = (cx){re, im};
        S->nhamp = 4;
        return a2;
      }
      return tensor_me4(S);
    }
    default: S->nhamp = 0; return NAN;
  }
}

/* ---- screening -------------------------------------------------------------------------------------- */

/* reference src/MEikonal.cc:489-512 IArray1D::Interpolate1D (linear grid) */
static cx interpolate1d(const grb_oracle_eikonal *E, double a) {
  const double MIN = E->min, MAX = E->max;
  const double STEP = (MAX - MIN) / E->n; /* IArray1D::Set, MEikonal.h:134-138 */
  if (a < MIN) a = MIN;
  int i = (int)floor((a - MIN) / STEP);
  if (i < 0) i = 0;
  if (i >= E->n) i = E->n - 1;
  const double dx = (a - E->x[i]);
  const double h  = (E->x[i + 1] - E->x[i]);
  /* F[i] + (a - x_i) * ((F[i+1] - F[i]) / (x_{i+1} - x_i)), complex/real then real*complex */
  const cx slope = {(E->re[i + 1] - E->re[i]) / h, (E->im[i + 1] - E->im[i]) / h};
  return (cx){E->re[i] + dx * slope.re, E->im[i] + dx * slope.im};
}

/* reference include/Graniitti/MMath.h:502-516 Simpson38Weight */
static double simpson38_w(int i, int N) {
  if (i == 0 || i == N) return 1.0;
  return (i % 3 == 0) ? 2.0 : 3.0;
}

/* reference src/MProcess.cc:157-288 S3ScreenedAmp2 (non-gg initial states) */
static double screened_amp2(ost *S) {
  const grb_process_desc *d = S->d;
  if (!d->screening) {
    OPC(OPC_BORN);
    const double a2 = bare_amp2(S);
    OPC(OPC_KIN);
    return a2;
  }
  OPC(OPC_BORN);
  bare_amp2(S);
  OPC(OPC_KIN);
  if (S->nhamp == 0) return NAN;
  const int nh = S->nhamp;
  cx        hamp_0[GRB_MAX_HAMP];
  memcpy(hamp_0, S->hamp, sizeof(cx) * nh);
  const double p1T[2] = {S->pfinal[1].x, S->pfinal[1].y};
  const double p2T[2] = {S->pfinal[2].x, S->pfinal[2].y};
  memcpy(S->pfinal_orig, S->pfinal, sizeof(S->pfinal));
  const int    NKT = d->loop_nkt, NPHI = d->loop_nphi;
  const double StepKT  = (d->loop_max_kt - d->loop_min_kt) / NKT;
  const double MinPhi  = 0.0;
  const double MaxPhi  = 2.0 * O_PI;
  const double StepPhi = (MaxPhi - MinPhi) / NPHI;
  cx           I[GRB_MAX_HAMP];
  for (int h = 0; h < nh; ++h) I[h] = (cx){0, 0};
  /* Simpson38Integral2D (MMath.h:456-472) sums W[i][j]*f[i][j] row-major; nodes whose kinematics fail
   * keep f = 0 and add (signed) zero, so accumulating on the fly visits the same sequence of adds */
  for (int i = 0; i < NPHI + 1; ++i) {
    const double phi = MinPhi + i * StepPhi;
    for (int j = 0; j < NKT + 1; ++j) {
      OPC(OPC_NODE_INV); /* what depends on (i, j) only: hoisted to table-set time by any batched implementation */
      const double kt    = d->loop_min_kt + j * StepKT;
      const double kt2   = pow2(kt);
      const cx     A_eik = interpolate1d(S->eik, kt2);
      const double ktx = kt * cos(phi), kty = kt * sin(phi);
      const double Wij = simpson38_w(i, NPHI) * simpson38_w(j, NKT);
      OPC(OPC_NODE_KIN);
      if (!(d->ps_class == GRB_CLASS_C ? loop_kinematics_C : loop_kinematics)(S, p1T[0] - ktx, p1T[1] - kty, p2T[0] + ktx, p2T[1] + kty)) {
        for (int h = 0; h < nh; ++h) { I[h].re += Wij * 0.0; I[h].im += Wij * 0.0; }
        continue;
      }
      OPC(OPC_NODE_AMP);
      bare_amp2(S);
      OPC(OPC_NODE_ACC);
      for (int h = 0; h < nh; ++h) {
        const cx f = cx_scale(cx_mul(A_eik, S->hamp[h]), kt);
        I[h].re += Wij * f.re; /* double * complex */
        I[h].im += Wij * f.im;
      }
    }
  }
  OPC(OPC_POST);
  const double scale = 9 * (StepPhi * StepKT) / 64;
  /* norm = zi / (8 pi^2 s): complex / double */
  const cx norm = {0.0 / (8.0 * (O_PI * O_PI) * d->s), 1.0 / (8.0 * (O_PI * O_PI) * d->s)};
  cx       hamp_loop[GRB_MAX_HAMP];
  for (int h = 0; h < nh; ++h) hamp_loop[h] = cx_mul(norm, cx_scale(I[h], scale));
  (d->ps_class == GRB_CLASS_C ? loop_kinematics_C : loop_kinematics)(S, p1T[0], p1T[1], p2T[0], p2T[1]);
  double amp2 = 0.0;
  for (int h = 0; h < nh; ++h) amp2 += cx_abs2(cx_add(hamp_0[h], hamp_loop[h]));
  if (nh != 1) amp2 /= 4;
  memcpy(S->hamp, hamp_0, sizeof(cx) * nh); /* for the dump: born amplitudes */
  OPC(OPC_KIN);
  return amp2;
}

/* ---- weight ----------------------------------------------------------------------------------------- */

/* reference src/MFactorized.cc:440-450 B51PhaseSpaceWeight */
static double b51_ps_weight(const ost *S) {
  const double J = 1.0 / fabs(S->pfinal[1].z / S->pfinal[1].e - S->pfinal[2].z / S->pfinal[2].e);
  const double twopi = 2.0 * O_PI;
  const double pow5  = twopi * ((twopi * twopi) * (twopi * twopi));
  return (1.0 / 2.0) * (1.0 / pow5) * (v4_pt(S->pfinal[1]) / (2.0 * S->pfinal[1].e)) *
         (v4_pt(S->pfinal[2]) / (2.0 * S->pfinal[2].e)) * J;
}

/* reference src/MFactorized.cc:432-437 B51IntegralVolume + src/MProcess.cc:1390-1400 ForwardVolume */
static double b51_integral_volume(const ost *S) {
  const grb_gencut *gc      = &S->d->gcuts;
  const double      PHI_vol = pow2(2.0 * O_PI);
  const double      J_pt    = v4_pt(S->pfinal[1]) * v4_pt(S->pfinal[2]);
  const double      PT_vol  = J_pt * pow2(log(gc->forward_pt_max) - log(gc->forward_pt_min + O_ZERO_EPS));
  double            fv      = PHI_vol * PT_vol;
  if (S->d->excitation == 1) { /* MProcess.cc:1402-1411: log-change of variable jacobian */
    const double J_M2 = S->excite1 ? v4_m2(S->pfinal[1]) : v4_m2(S->pfinal[2]);
    fv = PHI_vol * PT_vol * J_M2 * (S->log_M2_f_max - S->log_M2_f_min);
  } else if (S->d->excitation == 2) {
    const double J_M2 = v4_m2(S->pfinal[1]) * v4_m2(S->pfinal[2]);
    fv = PHI_vol * PT_vol * J_M2 * pow2(S->log_M2_f_max - S->log_M2_f_min);
  }
  return (pow2(S->M_MAX) - pow2(S->M_MIN)) * (gc->Y_max - gc->Y_min) * fv;
}

/* reference src/MContinuum.cc:639-658 BNPhaseSpaceWeight, :619-637 BNIntegralVolume (Kf = 2, Nf = 4) */
static double bn_ps_weight(const ost *S) {
  const double J = 1.0 / fabs(S->pfinal[1].z / S->pfinal[1].e - S->pfinal[2].z / S->pfinal[2].e);
  const int    Nf   = S->d->n_legs + 2;
  double       PROD = 1.0;
  for (int i = 0; i < S->d->n_legs - 1; ++i) PROD *= v4_pt(S->pkt[i]);
  return (1.0 / pow(2, 2 * (Nf - 2))) * (1.0 / pow(2.0 * O_PI, 3 * Nf - 4)) * (v4_pt(S->pfinal[1]) / (2.0 * S->pfinal[1].e)) *
         (v4_pt(S->pfinal[2]) / (2.0 * S->pfinal[2].e)) * J * PROD;
}
static double bn_integral_volume(const ost *S) {
  const grb_gencut *gc      = &S->d->gcuts;
  const int         Kf      = S->d->n_legs;
  const double      PHI_vol = pow2(2.0 * O_PI);
  const double      J_pt    = v4_pt(S->pfinal[1]) * v4_pt(S->pfinal[2]);
  const double      PT_vol  = J_pt * pow2(log(gc->forward_pt_max) - log(gc->forward_pt_min + O_ZERO_EPS));
  double            fv      = PHI_vol * PT_vol;
  if (S->d->excitation == 1) { /* ForwardVolume, MProcess.cc:1402-1411 (as in b51_integral_volume) */
    const double J_M2 = S->excite1 ? v4_m2(S->pfinal[1]) : v4_m2(S->pfinal[2]);
    fv = PHI_vol * PT_vol * J_M2 * (S->log_M2_f_max - S->log_M2_f_min);
  } else if (S->d->excitation == 2) {
    const double J_M2 = v4_m2(S->pfinal[1]) * v4_m2(S->pfinal[2]);
    fv = PHI_vol * PT_vol * J_M2 * pow2(S->log_M2_f_max - S->log_M2_f_min);
  }
  double            KT_vol  = pow(gc->kt_max - gc->kt_min, Kf - 1);
  KT_vol *= pow(2.0 * O_PI, Kf - 1);
  KT_vol *= pow(gc->rap_max - gc->rap_min, Kf);
  return KT_vol * fv;
}

/* reference src/MFactorized.cc:142-184 EventWeight; src/MContinuum.cc:164-193 for the <C> class */
static double event_weight(ost *S, const double *u, uint8_t *flags) {
  const grb_process_desc *d = S->d;
  double                  W = 0.0;
  OPC(OPC_KIN);
  const int kin  = d->ps_class == GRB_CLASS_C ? bn_random_kin(S, u) : b51_random_kin(S, u);
  const int fid  = common_cuts(S);
  const int veto = veto_cuts(S);
  if (kin && fid && veto && d->ps_class == GRB_CLASS_C) {
    const double MatESQ = screened_amp2(S);
    /* CascadePS() == 1 (no amplitude on the path activates a cascade phase space) */
    W = 1.0 * (1.0 / d->S_factor) * bn_ps_weight(S) * bn_integral_volume(S) * MatESQ * O_GeV2barn / (2.0 * d->s);
  } else if (kin && fid && veto) {
    const double MatESQ  = screened_amp2(S);
    double       C_space = 1.0;
    if (d->n_legs != 0 && d->ps_active) {
      C_space = S->DW / (2 * O_PI);
      C_space *= 1.0; /* CascadePS() == 1: no amplitude on the device path sets branch.PS_active */
    }
    W = C_space * (1.0 / d->S_factor) * b51_ps_weight(S) * b51_integral_volume(S) * MatESQ * O_GeV2barn /
        (2.0 * d->s);
  }
  int amp = 1;
  if (isnan(W) || isinf(W)) {
    W   = 0;
    amp = 0;
  }
  *flags = (uint8_t)((kin ? GRB_FLAG_KIN : 0) | (fid ? GRB_FLAG_FID : 0) | (veto ? GRB_FLAG_VETO : 0) |
                     (amp ? GRB_FLAG_AMP : 0));
  return W;
}

/* ---- public entry ----------------------------------------------------------------------------------- */

/* test hook: per evaluated point, the final-state momenta the amplitude saw -- pa, pb, p1, p2, pX, p3, p4 as (x, y, z, e),
 * then s1, s2, t1, t2, Re / Im of the first helicity amplitude: GRB_ORACLE_CAPTURE_LEN doubles.  Born-level runs only (the
 * screening loop overwrites the state).  Used by tests/test_tensor_res.py to feed the host-compiled closed form. */
#define GRB_ORACLE_CAPTURE_LEN 34
static double *g_capture = NULL;
void grb_oracle_capture(double *buf) { g_capture = buf; }

int grb_oracle_eval(const grb_process_desc *d, const grb_oracle_eikonal *eik, int64_t n, const double *u,
                    int dim, const double *tape, int dtape, double *w, uint8_t *flags, double *dump,
                    int32_t *tape_used) {
  if (!d || !u || !w || !flags) return GRB_ERR_ARG;
  if ((d->ps_class != GRB_CLASS_F && d->ps_class != GRB_CLASS_C) || d->flatamp != 0) return GRB_ERR_UNSUPPORTED;
  /* forward N* excitation (NSTARS = 1, 2) is restated for the <F> class Regge continuum (PomPomProtonVertex), yy[CON] (IncohFlux) and the
   * tensor-pomeron continuum (MTensorPomeron::ME4 has no inelastic vertex: the elastic formulas are evaluated on the excited kinematics) */
  if (d->excitation != 0 && !(d->proc == GRB_PROC_PP_CON || d->proc == GRB_PROC_PP_CON_MINUS || d->proc == GRB_PROC_YY_CON ||
                              d->proc == GRB_PROC_PP_CONTENSOR || d->proc == GRB_PROC_PP_RES || d->proc == GRB_PROC_PP_RES_CON ||
                              d->proc == GRB_PROC_PP_RESTENSOR || d->proc == GRB_PROC_PP_RES_CONTENSOR))
    return GRB_ERR_UNSUPPORTED;
  if (d->ps_class == GRB_CLASS_C && !(d->n_legs == 2 || ((d->n_legs == 4 || d->n_legs == 6) && d->proc == GRB_PROC_PP_CON))) return GRB_ERR_UNSUPPORTED;
  if (d->screening && (!eik || eik->n < 1)) return GRB_ERR_STATE;
  if (dim < d->lipsdim) return GRB_ERR_ARG;
  ost *S = (ost *)calloc(1, sizeof(ost));
  S->d      = d;
  S->eik    = eik;
  S->pbeam1 = (v4){0, 0, d->pbeam1_pz, d->pbeam1_E};
  S->pbeam2 = (v4){0, 0, d->pbeam2_pz, d->pbeam2_E};
  S->beamsum = v4_add(S->pbeam1, S->pbeam2);
  if (fabs(S->beamsum.z) > 1e-6) {
    free(S);
    return GRB_ERR_UNSUPPORTED;
  }
  for (int k = 0; k < 3; ++k) S->regge_eta0[k] = regge_eta(d->regge.a0[k], d->regge.sgn[k]);
  int rc = GRB_OK;
  for (int64_t k = 0; k < n; ++k) {
    S->tape         = tape ? tape + k * dtape : NULL;
    S->tape_len     = tape ? dtape : 0;
    S->tape_pos     = 0;
    S->tape_overrun = 0;
    g_call_idx      = 0;
    w[k]            = event_weight(S, u + k * dim, &flags[k]);
    if (tape_used) tape_used[k] = S->tape_overrun ? -S->tape_pos : S->tape_pos;
    if (g_capture) {
      double *c = g_capture + k * GRB_ORACLE_CAPTURE_LEN;
      const v4 mv[7] = {S->pbeam1, S->pbeam2, S->pfinal[1], S->pfinal[2], S->pfinal[0], S->node_p4[0], S->node_p4[1]};
      for (int i = 0; i < 7; ++i) { c[4 * i] = mv[i].x; c[4 * i + 1] = mv[i].y; c[4 * i + 2] = mv[i].z; c[4 * i + 3] = mv[i].e; }
      c[28] = S->s1; c[29] = S->s2; c[30] = S->t1; c[31] = S->t2; c[32] = S->hamp[0].re; c[33] = S->hamp[0].im;
    }
    if (dump) {
      double *D = dump + k * GRB_DUMP_LEN;
      D[GRB_DUMP_T1] = S->t1; D[GRB_DUMP_T2] = S->t2; D[GRB_DUMP_S1] = S->s1; D[GRB_DUMP_S2] = S->s2;
      D[GRB_DUMP_M2] = S->m2; D[GRB_DUMP_Y] = S->Y; D[GRB_DUMP_PT] = S->Pt;
      D[GRB_DUMP_THAT] = S->t_hat; D[GRB_DUMP_UHAT] = S->u_hat; D[GRB_DUMP_AMP2] = 0.0;
      D[GRB_DUMP_HAMP0_RE] = S->hamp[0].re; D[GRB_DUMP_HAMP0_IM] = S->hamp[0].im;
    }
  }
  free(S);
  return rc;
}
