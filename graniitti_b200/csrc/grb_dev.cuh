// grb_dev.cuh -- device-side building blocks of the event-weight path (sm_100a, fp64 CUDA cores).
//
// This translation unit is compiled with -fmad=false: every a*b+c written below is a separately rounded
// multiply and add, exactly like the reference build (g++ -O3 without -march=native emits no FMA on
// x86-64, reference Makefile:195-205,232-239).  That matters for the cancellation-prone kinematics
// (t1 = (p_beam - p1)^2 formed from O(6500 GeV) components).  Where contraction is
// harmless and pays (amplitude algebra inside the screening loop) fma() is written explicitly.
//
// Nothing here is a translation of reference classes: momenta live in registers as plain doubles, the
// process is a POD in constant memory, decay trees are flattened, and every loop invariant of the
// screening integral is hoisted to table-set time.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/grb.h"
#include "grb_fastmath.cuh"
#include "grb_tres.cuh"

namespace grb {
constexpr int ITEM_ROWS = 40;  // rows of the screening work list (the layout is the IT_* enum of grb_accum.cuh)


// reference include/Graniitti/MMath.h:30
#define GRB_PI 3.141592653589793238462643383279502884197169399375105820974944

struct v4 {
  double x, y, z, e;
};
struct cx {
  double re, im;
};

// ---- flattened, pre-digested process constants (constant memory) -------------------------------------
struct DevNode {
  double mass, width;
  double m2, m0G;        // m0^2, m0*Gamma
  double m2max, m2lo;    // (m0 + L*G)^2, (m0 - L*G)^2
  double bwA0, bwBA0;    // atan2(m2min - m2, m0G) and (atan2(m2max - m2, m0G) - that) for M_MIN = 0
  int    first_child, n_children, bw_active, pad_;
};

struct DevProc {
  // process selection
  int proc, screening, n_nodes, n_legs, n_decays, lipsdim, fid_active, usercuts, veto_active, n_veto;
  int ps_active, offshellFF, regge_n[3], pad_;
  int decay_order[GRB_MAX_NODES];  // nodes with daughters in the reference's depth-first order
  // beams
  double s, sqrt_s, s_pow15;       // s^(3/2) as the reference's std::pow(s, 3.0/2.0)
  double pb1z, pb1e, pb2z, pb2e, bsum_z, bsum_e;
  double mb1, mb2;                 // PDG beam masses (SampleForwardMasses, MProcess.cc:1318)
  double M_MAX0;                   // sqrt_s - (pbeam1.M() + pbeam2.M()), MFactorized.cc:277
  // sampling map (MFactorized.cc:245-258)
  double log_pt0, dlog_pt, Y_min, dY, gM_min, gM_max;
  // <C> class sampling map (MContinuum.cc:276-331): kt = kt_min + dkt u, y_i = rap_min + drap u; weight constant
  double kt_min, dkt, rap_min, drap, w_const_C;
  int    ps_class, pad_c_;
  // forward N* excitation (NSTARS, MProcess.cc:1317-1365,1402-1411; MRegge.cc:637-644): log-flat M^2 sampling range,
  // sqrt(g3P gN_P) / gN_P (inelastic over elastic proton vertex coupling), forward-mass fiducial cut
  int    excitation, f2_model;  // NSTARS; PARAM_STRUCTURE.F2 (0 CKMT, 1 DL)
  double log_M2f_min, dlog_M2f, exc_vertex, f_fM_min, f_fM_max;
  double offshell;
  // cuts
  double f_eta_min, f_eta_max, f_rap_min, f_rap_max, f_pt_min, f_pt_max, f_Et_min, f_Et_max;
  double f_M_min, f_M_max, f_Y_min, f_Y_max, f_Pt_min, f_Pt_max, f_t_min, f_t_max;
  grb_vetodomain veto[GRB_MAX_VETO];
  // weight: (1/S) * GeV2barn / (2 s) * (2pi)^2 * dlog_pt^2 * (Ymax-Ymin) * 1/2 * (2pi)^-5 * 1/4
  double w_const;
  // Regge continuum (MRegge.cc:394-438, 888-906; MForm.cc:382-396,433-435)
  double a0[3], ap[3], ln_inv_s0, eta0_re[3], eta0_im[3], phase_c[3];  // phase_c = -pi/2 * ap
  double gN_P, fc1, fc2, gpp_P, M2_leg0, b_EXP, a_OREAR, b_OREAR, b_POW, me4_sign;
  double M2_leg2;  // PDG mass^2 of the third central leg (ME6: propagator mass of the lower half of the ladder)
  double M2_leg4;  // PDG mass^2 of the fifth central leg (ME8: propagator mass of the lowest third of the ladder)
  double inv_fc1, inv_fc2, inv_b_POW, g2_const, eta0sq_re, eta0sq_im;  // g2_const = (gpp_P gN_P)^2
  double omega;                                                       // PARAM_REGGE::omega (ME3 flux power)
  double yy_ee4;                                                      // ee^4, ee = 2 sqrt(pi alpha(0)) (MG5 GC_3)
  // tensor pomeron (MTensorPomeron.cc:572-763): 36 gPNN^2 g^2, delta_P, alpha'_P, ln alpha'_P, -pi/2, (2 mp)^2
  double tp_coupling, tp_delta, tp_ap, tp_ln_ap, tp_phase, four_mp2;
  double tp_g1, tp_g2, tp_vv_coupling;  // 2xV: Pomeron-V-V couplings (a, b of gPrhorho / gPphiphi), 9 gPNN^2
  // 2xF / yy[QED] (grb_ferm.cuh): 81 gPNN^4, PDG mass of the central fermion, e = sqrt(4 pi alpha(0)), 1/(2 m_p),
  // beam helicity spinors u(p_beam, -1/+1) (real: phi = 0), evaluated by the host libm as MDirac::uHelDirac does
  double tp_ff_coupling, mass_leg0, qed_e, inv_2mp;
  double bsp[2][2][4];
  int    em_kelly, amp_kind;
  DevNode nodes[GRB_MAX_NODES];
};

// resonances (constant memory): pre-digested grb_resonance + Wigner small-d coefficient tables
struct DevResonance {
  int    spinX2, P, BW, spingen, spindec, frame, photo, pad_;
  double mass, width, M2, g_re, g_im, inv_gFF4, g_decay, B0;  // B0: PhotoProp slope (MRegge.cc:852-853)
  double T_re[9], T_im[9];
  double rho_re[25], rho_im[25];
};
struct DevRes {
  int          n, any_photo, bw3, pad_;
  DevResonance r[GRB_MAX_RES];
  tres::ResPar tp[GRB_MAX_RES];  // the same resonances as MTensorPomeron::ME3 reads them (PP[RESTENSOR], PP[RES+CONTENSOR])
  // wd[J-1][(i*5+j)*5+k] = sqrt((J+m)!(J-m)!(J+mp)!(J-mp)!) (-1)^(mp-m+k) / [(J-mp-k)!(J+m-k)!(k+mp-m)! k!],
  // m = i-J, mp = j-J (reference src/MSpin.cc:927-950); 0 where a factorial argument is negative
  double wd[2][125];
};

// screening loop discretisation + per-node eikonal amplitude (constant memory)
struct DevLoop {
  int    nkt, nphi;                       // intervals: nodes are nkt+1, nphi+1
  int    nphi_nodes, pad_;                // phi nodes visited: nphi+1, or nphi when phi=0 and phi=2pi are folded
  double kt[GRB_MAX_LOOP_KT];             // kt_j = MinLoopKT + j*StepKT
  double aeik_re[GRB_MAX_LOOP_KT];        // A_eik(kt_j^2), Interpolate1D done at table-set time
  double aeik_im[GRB_MAX_LOOP_KT];
  double wkt[GRB_MAX_LOOP_KT];            // Simpson 3/8 weights
  double b_re[GRB_MAX_LOOP_KT];           // A_eik(kt_j^2) * kt_j * wkt_j
  double b_im[GRB_MAX_LOOP_KT];
  double cphi[GRB_MAX_LOOP_PHI], sphi[GRB_MAX_LOOP_PHI], wphi[GRB_MAX_LOOP_PHI];
  double norm;                            // 9*StepPhi*StepKT/64 / (8 pi^2 s); hamp_loop = i*norm*sum
};

// arguments of one k_points / k_screen launch (one chunk of points)
struct PointArgs {
  int64_t n;
  // parity-mode inputs: SoA u[dim][ld], tape[dtape][ld], pointers pre-offset to this chunk's first point
  const double *u, *tape;
  int64_t       ld;
  int           dim, dtape;
  int           tape_rng;  // parity mode with internal draws from the point's Philox stream (seed, iter, first_call + idx)
  // VEGAS-mode inputs
  int           vegas, bins, mode;
  const double *xmat;  // [bins][dim] device
  uint64_t      seed;
  uint32_t      iter;
  int64_t       first_call;
  double        maxf;
  // per-point outputs (parity mode; nullable)
  double * w;
  uint8_t *flags;
  double * dump;
  // screening work list
  double *      items;
  int32_t *     item_idx;
  uint64_t *    item_bins;
  uint64_t *    item_bins_hi;  // VEGAS bins of dimensions 8..15 (four-leg <C> class: 14 dimensions); six legs (20): a third word
                               // for dimensions 16.. at [item_cap + slot]
  unsigned int *item_count;
  unsigned int *tile_count;    // k_points' tile counter (one unsigned, zeroed per launch)
  int64_t       item_cap;
  // accumulators (device)
  grb_accum *acc;
  // event generation (mode == GRB_MODE_GENERATE): hit-or-miss accepted events, SoA [npart*4 + 2][ev_cap]
  // rows: px,py,pz,E of lts.pfinal[0..2] then of every decay-tree node; then W and f = W * vegasweight
  double *      ev_out;
  unsigned int *ev_count;    // [0] accepted events written, [1] overflows (f > maxf)
  int64_t       ev_cap;
  int           npart;       // 3 + n_nodes
  double *      pay;         // screened generation: per work-list slot payload [npart*4][item_cap]
  // launch shaping for the k_points / k_screen overlap (grb_api.cu): k_points beside a running k_screen takes one
  // 128-thread CTA per SM; k_screen then leaves a quarter of the register file free (3 of its 4 resident CTAs per SM)
  int side_launch, screen_ctas;
  int sm_count;   // (host side: launch sizing)
};

// launchers implemented in grb_kernels.cu
cudaError_t upload_proc(const DevProc &p, cudaStream_t st);
cudaError_t upload_loop(const DevLoop &l, cudaStream_t st);
cudaError_t upload_res(const DevRes &r, cudaStream_t st);
int         launch_points(const PointArgs &a, int n_nodes, int amp_kind, int sm_count, cudaStream_t st);
int         launch_screen(const PointArgs &a, int amp_kind, int regge_ff, int sm_count, cudaStream_t st);
cudaError_t run_dfma_peak(int sm_count, double *tflops);
cudaError_t run_fastmath(int kind, int64_t n, const double *x_host, double *y_host);

// ---- small math ----------------------------------------------------------------------------------------
__device__ __forceinline__ double msqrt(double x) { return sqrt(fmax(0.0, x)); }  // MMath.h:250
__device__ __forceinline__ double pow2(double x) { return x * x; }
__device__ __forceinline__ double m2_of(double x, double y, double z, double e) {   // M4Vec.h:123,139
  return e * e - (x * x + y * y + z * z);
}
__device__ __forceinline__ double m2_of(const v4 &a) { return m2_of(a.x, a.y, a.z, a.e); }
__device__ __forceinline__ double m_of(const v4 &a) {  // M4Vec.h:125
  const double m2 = m2_of(a);
  return (m2 > 0.0) ? msqrt(m2) : -msqrt(-m2);
}
__device__ __forceinline__ double pt2_of(const v4 &a) { return a.x * a.x + a.y * a.y; }
__device__ __forceinline__ double rap_of(const v4 &a) { return 0.5 * log((a.e + a.z) / (a.e - a.z)); }

// reference include/Graniitti/MKinematics.h:33-64 SolvePz3_A: the t22 chain is kept in the reference's
// association (it carries the rounding that decides the last bits of p1z); constants in s are hoisted.
// Provenance: the reference's formula is itself the output of a computer-algebra system (a closed-form root of the
// energy-momentum constraint, temporaries t2 .. t30); its term order IS the parity contract of the solved p_z, so the
// checked variants below restate it term for term -- the folded `FAST` variants next to them are this repo's own.
struct PzConst {
  double s, t7, t9, t14;  // s, sqrt(s), s^(3/2), s^2
};
// FAST = true: the branch-free form for the one-basic-block loop node -- the MUFU-seeded sqrt / division sequences run
// unconditionally and `dom` records whether their range guards held (the caller re-evaluates the node on the checked
// path otherwise); FAST = false: per-call guards with the libm fallback.
template <bool FAST>
__device__ __forceinline__ double solve_pz3_A_t(double m3, double pt3, double pt4, double pz5, double E5,
                                                const PzConst &c, bool &dom) {
  const double s   = c.s;
  const double t2  = E5 * E5;
  const double t3  = pt3 * pt3;
  const double t4  = pt4 * pt4;
  const double t5  = pz5 * pz5;
  const double t6  = m3 * m3;
  const double t7  = c.t7;
  const double t8  = s * t2 * 6.0;
  const double t9  = c.t9;
  const double t10 = t2 * t2;
  const double t11 = t3 * t3;
  const double t12 = t4 * t4;
  const double t13 = t5 * t5;
  const double t14 = c.t14;
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
  double t23;
  if (FAST) { dom = dom && sqrt_fast_ok(t22); t23 = sqrt_fast_core(t22); }
  else t23 = sqrt_fast_ok(t22) ? sqrt_fast_core(t22) : sqrt(t22);
  const double num = E5 * t23 * (1.0 / 2.0) - t7 * t23 * (1.0 / 2.0) + pz5 * (t3 * (1.0 / 2.0) - t4 * (1.0 / 2.0));
  const double den = s + t2 - t5 - E5 * t7 * 2.0;
  double       q;
  if (FAST) { q = div_fast_core(num, den); dom = dom && div_fast_ok(num, q); }
  else q = div_checked(num, den);
  return pz5 * (-1.0 / 2.0) - q;
}
__device__ __forceinline__ double solve_pz3_A(double m3, double pt3, double pt4, double pz5, double E5, const PzConst &c) {
  bool dom = true;
  return solve_pz3_A_t<false>(m3, pt3, pt4, pz5, E5, c, dom);
}

// reference MKinematics.h:117-160 SolvePz3_B (unequal forward masses)
template <bool FAST>
__device__ __forceinline__ double solve_pz3_B_t(double m3, double m4, double pt3, double pt4, double pz5,
                                                double E5, const PzConst &c, bool &dom) {
  const double s   = c.s;
  const double t2  = E5 * E5;
  const double t3  = m3 * m3;
  const double t4  = m4 * m4;
  const double t5  = pt3 * pt3;
  const double t6  = pt4 * pt4;
  const double t7  = pz5 * pz5;
  const double t8  = c.t7;
  const double t9  = s * t2 * 6.0;
  const double t10 = c.t9;
  const double t11 = t2 * t2;
  const double t12 = t3 * t3;
  const double t13 = t4 * t4;
  const double t14 = t5 * t5;
  const double t15 = t6 * t6;
  const double t16 = t7 * t7;
  const double t17 = c.t14;
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
  double t30;
  if (FAST) { dom = dom && sqrt_fast_ok(t29); t30 = sqrt_fast_core(t29); }
  else t30 = sqrt_fast_ok(t29) ? sqrt_fast_core(t29) : sqrt(t29);
  const double num = E5 * t30 * (1.0 / 2.0) - t8 * t30 * (1.0 / 2.0) +
                     pz5 * (t3 * (1.0 / 2.0) - t4 * (1.0 / 2.0) + t5 * (1.0 / 2.0) - t6 * (1.0 / 2.0));
  const double den = s + t2 - t7 - E5 * t8 * 2.0;
  double       q;
  if (FAST) { q = div_fast_core(num, den); dom = dom && div_fast_ok(num, q); }
  else q = div_checked(num, den);
  return pz5 * (-1.0 / 2.0) - q;
}
__device__ __forceinline__ double solve_pz3_B(double m3, double m4, double pt3, double pt4, double pz5, double E5,
                                              const PzConst &c) {
  bool dom = true;
  return solve_pz3_B_t<false>(m3, m4, pt3, pt4, pz5, E5, c, dom);
}

// ---- SolvePz for the one-basic-block loop node: the same t22 / t29 chains with every power-of-two factor folded --------
// Multiplying by 2, 4, 8 or 1/2 is exact in binary floating point (no overflow / underflow at these magnitudes), so
// fl(fl(a b) 2) = fl(a (2 b)): the factors 2.0 / 4.0 / 8.0 / (1/2) of the reference's terms ride on a loop invariant
// (2 m^2, 2 s, 2 sqrt(s) ...) or on one per-node value shared by several terms (4 E5, 2 E5^2, 2 pz5^2).  Every term and
// the summation order are unchanged, so p1z comes out bit-identical (pinned by tests/test_fastmath.py on 2^22 argument
// sets) with ~30 fewer fp64 instructions per node.  PzInv = the per-point invariants.
struct PzInv {
  double m3sq, m4sq, m3sq_x2, m4sq_x2, m3q, m4q;  // m^2, 2 m^2, m^4
  double k_s3, k_s4, k_34;                        // s m3^2 2, s m4^2 2 (A: s m3^2 4), m3^2 m4^2 2
  double s, s_x2, rs, rs_x2, s15, ssq;            // s, 2 s, sqrt s, 2 sqrt s, s^(3/2), s^2
};
__device__ __forceinline__ PzInv make_pz_inv(double m3, double m4, const PzConst &c, bool same) {
  PzInv I;
  I.s = c.s; I.s_x2 = 2.0 * c.s; I.rs = c.t7; I.rs_x2 = 2.0 * c.t7; I.s15 = c.t9; I.ssq = c.t14;
  I.m3sq = m3 * m3; I.m4sq = m4 * m4; I.m3sq_x2 = 2.0 * I.m3sq; I.m4sq_x2 = 2.0 * I.m4sq;
  I.m3q = I.m3sq * I.m3sq; I.m4q = I.m4sq * I.m4sq;
  I.k_s3 = same ? c.s * I.m3sq * 4.0 : c.s * I.m3sq * 2.0;
  I.k_s4 = c.s * I.m4sq * 2.0;
  I.k_34 = I.m3sq * I.m4sq * 2.0;
  return I;
}
// key: running minimum of the square-root range keys (fast_key(), grb_fastmath.cuh), so that all range guards of a loop
// node cost one integer min each and ONE compare at the end; dom: the division's guard
__device__ __forceinline__ double solve_pz3_A_fast(const PzInv &I, double pt3, double pt4, double pz5, double E5, unsigned &key, bool &dom) {
  const double t2 = E5 * E5, t3 = pt3 * pt3, t4 = pt4 * pt4, t5 = pz5 * pz5;  // t6 = m3^2 rides in I.m3sq_x2, I.k_s3
  const double E4 = 4.0 * E5, t2x2 = 2.0 * t2, t5x2 = 2.0 * t5, t6x4 = 2.0 * I.m3sq_x2;
  const double t8  = I.s * t2 * 6.0;
  const double t10 = t2 * t2, t11 = t3 * t3, t12 = t4 * t4, t13 = t5 * t5;
  const double t15 = t5 * t6x4;                 // t5 t6 4
  const double t16 = t3 * t5x2, t17 = t4 * t5x2;  // t3 t5 2, t4 t5 2
  const double t18 = E4 * I.m3sq_x2 * I.rs;     // E5 t6 t7 8
  const double t19 = E4 * t3 * I.rs, t20 = E4 * t4 * I.rs, t21 = E4 * t5 * I.rs;
  const double t22 = t8 + t10 + t11 + t12 + t13 + I.ssq + t15 + t16 + t17 + t18 + t19 + t20 + t21 -
                     E4 * I.s15 - I.s_x2 * t3 - I.s_x2 * t4 - I.s_x2 * t5 - I.k_s3 -
                     t2x2 * t3 - t2x2 * t4 - t2x2 * t5 - t3 * t4 * 2.0 - t2 * t6x4 -
                     E4 * t2 * I.rs;
  key = min(key, fast_key(t22));
  const double t23h = 0.5 * sqrt_fast_core(t22);
  const double pz5h = 0.5 * pz5;
  const double num = E5 * t23h - I.rs * t23h + pz5h * (t3 - t4);
  const double den = I.s + t2 - t5 - E5 * I.rs_x2;
  const double q   = div_fast_core(num, den);
  dom = dom & div_fast_ok(num, q);
  return -pz5h - q;
}
// key: running minimum of the square-root range keys (fast_key(), grb_fastmath.cuh), so that all range guards of a loop
// node cost one integer min each and ONE compare at the end; dom: the division's guard
__device__ __forceinline__ double solve_pz3_B_fast(const PzInv &I, double pt3, double pt4, double pz5, double E5, unsigned &key, bool &dom) {
  const double t2 = E5 * E5, t3 = I.m3sq, t4 = I.m4sq, t5 = pt3 * pt3, t6 = pt4 * pt4, t7 = pz5 * pz5;
  const double E4 = 4.0 * E5, t2x2 = 2.0 * t2, t7x2 = 2.0 * t7;
  const double t9  = I.s * t2 * 6.0;
  const double t11 = t2 * t2, t14 = t5 * t5, t15 = t6 * t6, t16 = t7 * t7;
  const double t18 = I.m3sq_x2 * t5, t19 = I.m4sq_x2 * t6, t20 = I.m3sq_x2 * t7, t21 = I.m4sq_x2 * t7;
  const double t22 = t5 * t7x2, t23 = t6 * t7x2;
  const double t24 = E4 * t3 * I.rs, t25 = E4 * t4 * I.rs, t26 = E4 * t5 * I.rs, t27 = E4 * t6 * I.rs, t28 = E4 * t7 * I.rs;
  const double t29 = t9 + t11 + I.m3q + I.m4q + t14 + t15 + t16 + I.ssq + t18 + t19 + t20 + t21 + t22 + t23 +
                     t24 + t25 + t26 + t27 + t28 - E4 * I.s15 - I.k_s3 - I.k_s4 -
                     I.s_x2 * t5 - I.s_x2 * t6 - I.s_x2 * t7 - t2 * I.m3sq_x2 - t2 * I.m4sq_x2 -
                     t2x2 * t5 - I.k_34 - t2x2 * t6 - t2x2 * t7 - I.m3sq_x2 * t6 -
                     I.m4sq_x2 * t5 - t5 * t6 * 2.0 - E4 * t2 * I.rs;
  key = min(key, fast_key(t29));
  const double t30h = 0.5 * sqrt_fast_core(t29);
  const double pz5h = 0.5 * pz5;
  const double num = E5 * t30h - I.rs * t30h + pz5h * (t3 - t4 + t5 - t6);
  const double den = I.s + t2 - t7 - E5 * I.rs_x2;
  const double q   = div_fast_core(num, den);
  dom = dom & div_fast_ok(num, q);
  return -pz5h - q;
}

__device__ __forceinline__ double solve_pz3_A_fast(const PzInv &I, double pt3, double pt4, double pz5, double E5, bool &dom) {
  unsigned key = 0xffffffffu;
  const double r = solve_pz3_A_fast(I, pt3, pt4, pz5, E5, key, dom);
  dom = dom && fast_key_ok(key);
  return r;
}
__device__ __forceinline__ double solve_pz3_B_fast(const PzInv &I, double pt3, double pt4, double pz5, double E5, bool &dom) {
  unsigned key = 0xffffffffu;
  const double r = solve_pz3_B_fast(I, pt3, pt4, pz5, E5, key, dom);
  dom = dom && fast_key_ok(key);
  return r;
}

__device__ __forceinline__ double solve_pz(double m3, double m4, double pt3, double pt4, double pz5,
                                           double E5, const PzConst &c) {  // MKinematics.h:220-229
  if (fabs(m3 - m4) < 1e-9) return solve_pz3_A(m3, pt3, pt4, pz5, E5, c);
  return solve_pz3_B(m3, m4, pt3, pt4, pz5, E5, c);
}

// reference MKinematics.h:383-401 LorentzBoost (sign = +1); libstdc++'s (beta*p3).sum() adds from the
// last element, kept so that daughters match the reference to the last bit on equal inputs.
__device__ __forceinline__ v4 boost_from_rest(const v4 &mother, double M0, const v4 &p) {
  const double b0 = mother.x / mother.e, b1 = mother.y / mother.e, b2 = mother.z / mother.e;
  const double gamma  = mother.e / M0;
  const double kappa1 = (b2 * p.z + b1 * p.y) + b0 * p.x;
  const double kappa2 = gamma * (gamma * kappa1 / (1.0 + gamma) + p.e);
  v4           r;
  r.x = p.x + kappa2 * b0;
  r.y = p.y + kappa2 * b1;
  r.z = p.z + kappa2 * b2;
  r.e = gamma * (p.e + kappa1);
  return r;
}

// reference MKinematics.h:294-298 DecayMomentum
__device__ __forceinline__ double decay_momentum(double x, double y, double z) {
  return msqrt((x - y - z) * (x + y + z) * (x + y - z) * (x - y + z)) / (2.0 * x);
}

// reference MKinematics.h:403-459: FlatIsotropic + Isotropic + boost; returns dPhi2 (:267-271)
__device__ __forceinline__ double two_body(const v4 &mother, double M0, double m1, double m2, double r_cos,
                                           double r_phi, v4 &p1, v4 &p2) {
  const double pnorm    = decay_momentum(M0, m1, m2);
  const double costheta = -1.0 + (1.0 - -1.0) * r_cos;          // MRandom::U(-1,1), MRandom.cc:62-65
  const double sintheta = msqrt(1.0 - pow2(costheta));
  const double phi      = 0.0 + (2.0 * GRB_PI - 0.0) * r_phi;   // MRandom::U(0,2pi)
  double       sp, cp;
  sincos(phi, &sp, &cp);
  v4 k;
  k.x = pnorm * sintheta * cp;
  k.y = pnorm * sintheta * sp;
  k.z = pnorm * costheta;
  k.e = msqrt(pow2(m1) + pow2(pnorm));
  p1  = boost_from_rest(mother, M0, k);
  k.x = -k.x;
  k.y = -k.y;
  k.z = -k.z;
  k.e = msqrt(pow2(m2) + pow2(pnorm));
  p2  = boost_from_rest(mother, M0, k);
  const double V = 2.0 * 2.0 * GRB_PI;
  return V * pnorm / (16.0 * pow2(GRB_PI) * M0);
}

}  // namespace grb
