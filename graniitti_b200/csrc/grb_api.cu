// grb_api.cu -- the C ABI of include/grb.h: engine lifecycle, constant digestion, batched evaluation and
// the device VEGAS iteration.  No CPU fallback: every compute entry point needs a CUDA device and returns
// GRB_ERR_CUDA otherwise.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "grb_dev.cuh"
#include "grb_eikonal.h"
#include "grb_host.h"

using namespace grb;

struct grb_engine {
  int              device   = 0;
  int              sm_count = 148;
  cudaStream_t     stream   = nullptr;
  cudaEvent_t      ev0 = nullptr, ev1 = nullptr;
  grb_process_desc desc;
  DevProc          dproc;
  DevLoop          dloop;
  DevRes           dres;
  bool             have_loop = false;
  bool             const_dirty = true;  // the engine's parameter blocks changed since they were last uploaded
  // screening work list (one chunk)
  int64_t       chunk      = 0;
  double *      d_items    = nullptr;
  int32_t *     d_item_idx = nullptr;
  uint64_t *    d_item_bins = nullptr, *d_item_bins_hi = nullptr;
  unsigned int *d_item_count = nullptr;
  // second work list + side stream: k_points of chunk c+1 runs beside k_screen of chunk c (grb_vegas_iterate_async)
  double *      d_items2 = nullptr;
  int32_t *     d_item_idx2 = nullptr;
  uint64_t *    d_item_bins2 = nullptr, *d_item_bins_hi2 = nullptr;
  unsigned int *d_item_count2 = nullptr;
  cudaStream_t  stream_pts = nullptr;
  cudaEvent_t   ev_pts[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr}, ev_start = nullptr;
  int           overlap = 0;   // GRB_OVERLAP: 0 off, n > 0: on, k_screen with n CTAs per SM
  // event generation buffers (allocated on first GRB_MODE_GENERATE call)
  double *      d_ev = nullptr, *d_pay = nullptr;
  unsigned int *d_ev_count = nullptr;
  int64_t       ev_cap = 0;
  int64_t       ev_dropped = 0;  // accepted events of the last GENERATE iteration that did not fit ev_cap
  // VEGAS
  grb_accum *d_acc  = nullptr;
  double *   d_xmat = nullptr;
  // scratch for host-pointer callers
  void * d_scratch     = nullptr;
  size_t scratch_bytes = 0;
  // grb_eval with host buffers: copy-in / compute / copy-out pipeline over slices of the batch
  cudaStream_t stream_in = nullptr, stream_out = nullptr;
  cudaEvent_t  ev_in = nullptr, ev_done = nullptr;
  int64_t      host_slice = 0;
  // NCCL communicator (optional, loaded at run time)
  void *nccl_comm = nullptr;
  int   comm_rank = 0, comm_size = 1;
  double *d_max2 = nullptr;
  // timing of the last call: ev0/ev1 bracket all launches, evk[] bracket each k_screen launch
  std::vector<cudaEvent_t> evk;
  int                      n_evk = 0;
  bool        timed = false;
  int         last_launches = 0;
  std::string err;
};

// The parameter blocks (DevProc / DevLoop / DevRes) live in __constant__ memory: one copy per device.  Several engines
// may share a device (and be driven from several host threads), so the copy has an owner per device, guarded by a mutex
// that is held while a call enqueues its kernels; a new owner first makes its stream wait for the previous owner's last
// launch (an event recorded after every enqueue), then uploads its own blocks on its own stream.
struct ConstSlot {
  std::mutex        mu;
  const grb_engine *owner    = nullptr;
  cudaEvent_t       last_use = nullptr;
};
static ConstSlot         g_const[64];
static inline ConstSlot &const_slot(int device) { return g_const[(unsigned)device % 64u]; }
static thread_local std::string g_last_global_error;

#define ITEM_NF grb::ITEM_ROWS  /* rows of the screening work list (grb_dev.cuh; = IT_NF of grb_accum.cuh, asserted there) */

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      e->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                 \
      return GRB_ERR_CUDA;                                                                         \
    }                                                                                              \
  } while (0)

static double m_of_beam(double pz, double E) {  // M4Vec::M() of (0,0,pz,E), reference M4Vec.h:123-125
  const double m2 = E * E - (0.0 * 0.0 + 0.0 * 0.0 + pz * pz);
  return (m2 > 0.0) ? std::sqrt(std::max(0.0, m2)) : -std::sqrt(std::max(0.0, -m2));
}

// one active Regge trajectory (TUNE0): the screening kernel's one-basic-block node applies (grb_point.cuh)
static int regge_fast_ff(const DevProc &P) {
  if (P.ps_class == GRB_CLASS_F && P.excitation == 0 && P.amp_kind == 1) return 0;  // tensor pomeron 2xS
  if (P.ps_class == GRB_CLASS_C && P.n_legs == 2 && P.excitation == 0 && P.amp_kind == 0 && P.regge_n[0] && !P.regge_n[1] && !P.regge_n[2] &&
      P.offshellFF >= 0 && P.offshellFF <= 2)
    return 10 + P.offshellFF;                                                       // <C> class Regge continuum
  if (P.ps_class == GRB_CLASS_F && P.excitation == 0 && P.amp_kind == 0 && P.regge_n[0] && !P.regge_n[1] && !P.regge_n[2] &&
      P.offshellFF >= 0 && P.offshellFF <= 2)
    return (P.ln_inv_s0 == 0.0 ? 20 : 0) + P.offshellFF;  // s0 = 1 GeV^2: ln(1/s0) = 0 exactly
  return -1;
}

// desc -> DevProc. Host libm (glibc, the reference's) evaluates the loop invariants.
static int digest(const grb_process_desc &d, DevProc &P, std::string &err) {
  memset(&P, 0, sizeof(P));
  if (d.abi_version != GRB_ABI_VERSION) { err = "abi_version mismatch"; return GRB_ERR_ARG; }
  if (d.ps_class != GRB_CLASS_F && d.ps_class != GRB_CLASS_C) { err = "only the <F> and <C> phase-space classes are on the device path"; return GRB_ERR_UNSUPPORTED; }
  if (d.ps_class == GRB_CLASS_C && d.veto_active) { err = "<C> with veto cuts needs forward fragmentation (outside the device path)"; return GRB_ERR_UNSUPPORTED; }
  // (forward N* excitation: both phase-space classes, every amplitude on the device path)
  if (d.excitation < 0 || d.excitation > 2) { err = "MProcess::SetExcitation: Not valid input (0,1,2)"; return GRB_ERR_ARG; }
  if (d.flatamp != 0) { err = "@FLATAMP is outside the device path"; return GRB_ERR_UNSUPPORTED; }
  if (d.flatmass2 != 0) { err = "FLATMASS2 sampling is outside the device path"; return GRB_ERR_UNSUPPORTED; }
  if (d.proc != GRB_PROC_PP_CON && d.proc != GRB_PROC_PP_CON_MINUS && d.proc != GRB_PROC_PP_CONTENSOR &&
      d.proc != GRB_PROC_PP_RES && d.proc != GRB_PROC_PP_RES_CON && d.proc != GRB_PROC_YY_CON && d.proc != GRB_PROC_YY_QED &&
      d.proc != GRB_PROC_PP_RESTENSOR && d.proc != GRB_PROC_PP_RES_CONTENSOR) {
    err = "process id " + std::to_string(d.proc) + " has no device amplitude yet";
    return GRB_ERR_UNSUPPORTED;
  }
  if (d.proc == GRB_PROC_YY_CON) {
    const int a = std::abs(d.nodes[0].pdg);
    if (!(a == 11 || a == 13 || a == 15) || d.nodes[1].pdg != -d.nodes[0].pdg || d.n_nodes != 2) {
      err = "yy[CON]: only charged lepton pairs (MG5 yy -> l+ l-) have a device amplitude";
      return GRB_ERR_UNSUPPORTED;
    }
  }
  if (d.proc == GRB_PROC_YY_QED) {
    const int a = std::abs(d.nodes[0].pdg);
    if (!(a == 11 || a == 13 || a == 15) || d.nodes[1].pdg != -d.nodes[0].pdg || d.n_nodes != 2) {
      err = "yy[QED]: charged lepton pairs have a device amplitude (quark pairs need fragmentation downstream)";
      return GRB_ERR_UNSUPPORTED;
    }
  }
  if (d.proc == GRB_PROC_PP_CONTENSOR) {
    const bool s2 = d.nodes[0].spinX2 == 0 && d.nodes[1].spinX2 == 0;
    const bool f2 = d.nodes[0].spinX2 == 1 && d.nodes[1].spinX2 == 1 && d.n_nodes == 2;
    const bool v2 = d.nodes[0].spinX2 == 2 && d.nodes[1].spinX2 == 2 && d.nodes[0].pdg == d.nodes[1].pdg &&
                    (d.nodes[0].pdg == 113 || d.nodes[0].pdg == 333);
    if (!s2 && !v2 && !f2) {
      err = "PP[CONTENSOR]: pseudoscalar (2xS), rho0 rho0 / phi phi (2xV) and stable baryon (2xF) pairs have a device amplitude";
      return GRB_ERR_UNSUPPORTED;
    }
  }
  const bool me6 = d.n_legs == 4 && d.ps_class == GRB_CLASS_C && d.proc == GRB_PROC_PP_CON && d.n_nodes == 4;
  const bool me8 = d.n_legs == 6 && d.ps_class == GRB_CLASS_C && d.proc == GRB_PROC_PP_CON && d.n_nodes == 6;
  if (d.n_legs != 2 && !me6 && !me8) {
    err = "the device path covers 2 first-level central legs, and 4 / 6 stable legs for PP[CON]<C> (the MRegge::ME6 / ME8 ladders)";
    return GRB_ERR_UNSUPPORTED;
  }
  if ((me6 || me8) && d.excitation != 0) {  // (the reference's own ladder integration with N* ends in "Too many NaN weights")
    err = "forward excitation with the ME6 / ME8 ladders is outside the device path";
    return GRB_ERR_UNSUPPORTED;
  }
  if (d.lipsdim > GRB_MAX_DIM) { err = "lipsdim exceeds GRB_MAX_DIM"; return GRB_ERR_ARG; }
  if (d.n_nodes < 2 || d.n_nodes > GRB_MAX_NODES) { err = "bad n_nodes"; return GRB_ERR_ARG; }
  if (d.lipsdim != (d.ps_class == GRB_CLASS_C ? 3 * (d.n_legs + 2) - 4 : 6) + d.excitation) { err = "lipsdim must be 6 + NSTARS for <F> and 3 Nf - 4 + NSTARS for <C>"; return GRB_ERR_ARG; }
  if (std::fabs(d.pbeam1_pz + d.pbeam2_pz) > 1e-6) { err = "asymmetric beams are outside the device path"; return GRB_ERR_UNSUPPORTED; }
  if (d.n_veto > GRB_MAX_VETO) { err = "too many veto domains"; return GRB_ERR_ARG; }
  if (d.fcuts.usercuts != 0 && std::abs(d.fcuts.usercuts) != 3 && std::abs(d.fcuts.usercuts) != 111) {
    err = "USERCUTS id not supported on the device path";
    return GRB_ERR_UNSUPPORTED;
  }
  P.proc = d.proc; P.screening = d.screening; P.n_nodes = d.n_nodes; P.n_legs = d.n_legs; P.lipsdim = d.lipsdim;
  P.fid_active = d.fcuts.active; P.usercuts = d.fcuts.usercuts; P.veto_active = d.veto_active; P.n_veto = d.n_veto;
  P.ps_active = d.ps_active;
  P.s = d.s; P.sqrt_s = d.sqrt_s; P.s_pow15 = std::pow(d.s, 3.0 / 2.0);
  P.pb1z = d.pbeam1_pz; P.pb1e = d.pbeam1_E; P.pb2z = d.pbeam2_pz; P.pb2e = d.pbeam2_E;
  P.bsum_z = d.pbeam1_pz + d.pbeam2_pz; P.bsum_e = d.pbeam1_E + d.pbeam2_E;
  P.mb1 = d.beam1_mass; P.mb2 = d.beam2_mass;
  P.M_MAX0 = d.sqrt_s - (m_of_beam(d.pbeam1_pz, d.pbeam1_E) + m_of_beam(d.pbeam2_pz, d.pbeam2_E));
  P.log_pt0 = std::log(d.gcuts.forward_pt_min + 1e-12);
  P.dlog_pt = std::log(d.gcuts.forward_pt_max) - std::log(d.gcuts.forward_pt_min + 1e-12);
  P.Y_min = d.gcuts.Y_min; P.dY = d.gcuts.Y_max - d.gcuts.Y_min;
  P.gM_min = d.gcuts.M_min; P.gM_max = d.gcuts.M_max;
  P.ps_class = d.ps_class;
  P.excitation = d.excitation;
  P.f2_model = d.soft.f2_model;
  P.log_M2f_min = std::log(1.09 * 1.09);  // neutron + pi+ threshold, MProcess.cc:1324-1328
  P.dlog_M2f = std::log(d.gcuts.XI_max * d.s) - P.log_M2f_min;
  if (d.excitation != 0 && !(d.gcuts.XI_max * d.s > 1.09 * 1.09)) { err = "MProcess::SampleForwardMasses: Forward leg Xi max is below the inelastic threshold"; return GRB_ERR_ARG; }
  P.exc_vertex = std::sqrt(std::max(0.0, d.soft.g3P * d.soft.gN_P)) / d.soft.gN_P;
  P.f_fM_min = d.fcuts.forward_M_min; P.f_fM_max = d.fcuts.forward_M_max;
  P.kt_min = d.gcuts.kt_min; P.dkt = d.gcuts.kt_max - d.gcuts.kt_min;
  P.rap_min = d.gcuts.rap_min; P.drap = d.gcuts.rap_max - d.gcuts.rap_min;
  P.offshell = d.offshell;
  const grb_fidcut &f = d.fcuts;
  P.f_eta_min = f.eta_min; P.f_eta_max = f.eta_max; P.f_rap_min = f.rap_min; P.f_rap_max = f.rap_max;
  P.f_pt_min = f.pt_min; P.f_pt_max = f.pt_max; P.f_Et_min = f.Et_min; P.f_Et_max = f.Et_max;
  P.f_M_min = f.M_min; P.f_M_max = f.M_max; P.f_Y_min = f.Y_min; P.f_Y_max = f.Y_max;
  P.f_Pt_min = f.Pt_min; P.f_Pt_max = f.Pt_max; P.f_t_min = f.forward_t_min; P.f_t_max = f.forward_t_max;
  for (int i = 0; i < d.n_veto; ++i) P.veto[i] = d.veto[i];
  // weight constant, reference MFactorized.cc:173-174,432-450, MProcess.cc:1390-1400, MPDG.h:64-78
  const double PI       = GRB_PI;
  const double GeV2barn = ((6.582119514E-25 * 2.99792458E8) * (6.582119514E-25 * 2.99792458E8)) / 1E-28;
  const double twopi    = 2.0 * PI;
  P.w_const = (1.0 / d.S_factor) * GeV2barn / (2.0 * d.s) * (twopi * twopi) * (P.dlog_pt * P.dlog_pt) * P.dY *
              0.5 / (twopi * twopi * twopi * twopi * twopi) * 0.25;
  // <C>, Kf = 2 (MContinuum.cc:181-183,619-658): 2^-4 (2pi)^-8 / 4 from the phase-space weight, dkt 2pi drap^2 and the
  // forward volume (2pi)^2 dlog^2 from the integral volume
  {
    const int Kf = d.n_legs, Nf = Kf + 2;
    P.w_const_C = (1.0 / d.S_factor) * GeV2barn / (2.0 * d.s) * (1.0 / std::pow(2.0, 2 * (Nf - 2))) / std::pow(twopi, 3 * Nf - 4) * 0.25 *
                  (std::pow(P.dkt, Kf - 1) * std::pow(twopi, Kf - 1) * std::pow(P.drap, Kf)) * ((twopi * twopi) * (P.dlog_pt * P.dlog_pt));
  }
  // Regge
  const grb_param_regge &R = d.regge;
  for (int k = 0; k < 3; ++k) {
    P.a0[k] = R.a0[k]; P.ap[k] = R.ap[k]; P.regge_n[k] = R.n[k];
    // ReggeEta(a0, sgn), reference src/MForm.cc:382-385
    const double denom = std::sin(PI * R.a0[k]);
    const double y     = (-1.0 * PI) * R.a0[k];
    const double vr = 1.0 + R.sgn[k] * std::cos(y), vi = R.sgn[k] * std::sin(y);
    P.eta0_re[k] = -vr / denom; P.eta0_im[k] = -vi / denom;
    P.phase_c[k] = ((-1.0 * PI) / 2.0 * R.ap[k]);
  }
  P.ln_inv_s0 = -std::log(R.s0);
  P.offshellFF = R.offshellFF;
  P.gN_P = d.soft.gN_P; P.fc1 = d.soft.fc1; P.fc2 = d.soft.fc2;
  P.gpp_P = R.c[0] / d.soft.gN_P;
  P.M2_leg0 = d.nodes[0].mass * d.nodes[0].mass;
  P.M2_leg2 = d.n_nodes > 2 ? d.nodes[2].mass * d.nodes[2].mass : P.M2_leg0;
  P.M2_leg4 = d.n_nodes > 4 ? d.nodes[4].mass * d.nodes[4].mass : P.M2_leg2;
  P.b_EXP = R.b_EXP; P.a_OREAR = R.a_OREAR; P.b_OREAR = R.b_OREAR; P.b_POW = R.b_POW;
  P.me4_sign = (d.proc == GRB_PROC_PP_CON_MINUS) ? -1.0 : 1.0;
  P.inv_fc1 = 1.0 / P.fc1; P.inv_fc2 = 1.0 / P.fc2; P.inv_b_POW = 1.0 / P.b_POW;
  P.g2_const = (P.gpp_P * P.gN_P) * (P.gpp_P * P.gN_P);
  P.eta0sq_re = P.eta0_re[0] * P.eta0_re[0] - P.eta0_im[0] * P.eta0_im[0];
  P.eta0sq_im = 2.0 * P.eta0_re[0] * P.eta0_im[0];
  // tensor pomeron 2xS (reference src/MTensorPomeron.cc:726-763,1388-1399,2203-2215,2288-2298)
  {
    const double gP = (d.nodes[0].mass < 0.2) ? d.tpom.gPpipi : d.tpom.gPKK;
    P.tp_coupling = 36.0 * d.tpom.gPNN * d.tpom.gPNN * gP * gP;
    P.tp_delta = d.tpom.delta_P; P.tp_ap = d.tpom.ap_P; P.tp_ln_ap = std::log(d.tpom.ap_P);
    P.tp_phase = -GRB_PI / 2.0;
    P.four_mp2 = (2 * 0.938272081) * (2 * 0.938272081);  // PDG::mp, reference MPDG.h:54
    P.em_kelly = d.em_ff;
    P.amp_kind = (d.proc == GRB_PROC_PP_CONTENSOR) ? 1 : (d.proc == GRB_PROC_PP_RES ? 2 : (d.proc == GRB_PROC_PP_RES_CON ? 3 : 0));
    if (d.proc == GRB_PROC_YY_CON) P.amp_kind = 4;
    if (d.proc == GRB_PROC_PP_CONTENSOR && d.nodes[0].spinX2 == 2) {  // 2xV (MTensorPomeron.cc:797-813)
      P.amp_kind = 5;
      const double *gv = d.nodes[0].pdg == 113 ? d.tpom.gPrhorho : d.tpom.gPphiphi;
      P.tp_g1 = gv[0]; P.tp_g2 = gv[1];
      P.tp_vv_coupling = 9.0 * d.tpom.gPNN * d.tpom.gPNN;
    }
    if (d.proc == GRB_PROC_PP_CONTENSOR && d.nodes[0].spinX2 == 1) P.amp_kind = 6;  // 2xF (MTensorPomeron.cc:765-804)
    if (d.proc == GRB_PROC_YY_QED) P.amp_kind = 7;                                    // QED branch (:664-713)
    if (d.proc == GRB_PROC_PP_RESTENSOR) P.amp_kind = 10;       // MTensorPomeron::ME3 (grb_tres.cuh)
    if (d.proc == GRB_PROC_PP_RES_CONTENSOR) P.amp_kind = 11;   // ME4 (2xS) + ME3, PROC_14
    if (d.proc == GRB_PROC_PP_RESTENSOR || d.proc == GRB_PROC_PP_RES_CONTENSOR) P.tp_vv_coupling = 9.0 * d.tpom.gPNN * d.tpom.gPNN;
    if (me6) P.amp_kind = 8;                                                          // MRegge::ME6 ladder
    if (me8) P.amp_kind = 9;                                                          // MRegge::ME8 ladder
    P.tp_ff_coupling = 81.0 * std::pow(d.tpom.gPNN, 4);
    P.mass_leg0 = d.nodes[0].mass;
    P.qed_e = std::sqrt(std::max(0.0, 0.0072973525693 * 4.0 * GRB_PI));  // qed::alpha_QED() = alpha_0, MQED.h:19-24,57
    P.inv_2mp = 1.0 / (2 * 0.938272081);
    for (int b = 0; b < 2; ++b) {  // MDirac::uHelDirac(pbeam, -1/+1), src/MDirac.cc:224-246 (px = py = 0: phi = 0)
      const double pz = b == 0 ? d.pbeam1_pz : d.pbeam2_pz, E = b == 0 ? d.pbeam1_E : d.pbeam2_E;
      const double m = m_of_beam(pz, E), p3 = std::sqrt(std::max(0.0, pz * pz)), N = std::sqrt(std::max(0.0, E + m));
      const double th2 = ((pz == 0.0) ? 0.0 : std::atan2(0.0, pz)) / 2.0, c = std::cos(th2), s = std::sin(th2);
      const double r = N * p3 / (E + m);
      const double um[4] = {-N * s, N * c, r * s, -r * c}, up[4] = {N * c, N * s, r * c, r * s};
      for (int k = 0; k < 4; ++k) { P.bsp[b][0][k] = um[k]; P.bsp[b][1][k] = up[k]; }
    }
    P.omega = d.regge.omega;
    // QED_alpha = ZERO: alpha(0); otherwise the MG5 model coupling, aEWM1 = 132.507 (Parameters_sm.cc:31,45-49,93-101)
    const double ee = d.qed_alpha == 0 ? 2. * std::sqrt(1.0 / 137.03599908) * std::sqrt(M_PI) : 2 * std::sqrt(1. / 1.325070e+02) * std::sqrt(M_PI);
    P.yy_ee4 = (ee * ee) * (ee * ee);
  }
  // nodes
  for (int i = 0; i < d.n_nodes; ++i) {
    const grb_node &n = d.nodes[i];
    DevNode &       o = P.nodes[i];
    if (n.n_children != 0 && n.n_children != 2) { err = "only two-body decays are on the device path"; return GRB_ERR_UNSUPPORTED; }
    if (n.n_children && (n.first_child <= i || n.first_child + 1 >= d.n_nodes)) { err = "bad decay tree indices"; return GRB_ERR_ARG; }
    o.mass = n.mass; o.width = n.width; o.first_child = n.first_child; o.n_children = n.n_children;
    o.bw_active = !(n.width < 1e-40);
    o.m2 = n.mass * n.mass; o.m0G = n.mass * n.width;
    const double hi = n.mass + d.offshell * n.width, lo = n.mass - d.offshell * n.width;
    o.m2max = hi * hi; o.m2lo = lo * lo;
    // RelativisticBWRandom constants for M_MIN = 0, reference src/MRandom.cc:82-107
    const double m2min = std::max(std::max(0.0, 0.0), o.m2lo);
    o.bwA0  = std::atan2(m2min - o.m2, o.m0G);
    o.bwBA0 = std::atan2(o.m2max - o.m2, o.m0G) - std::atan2(m2min - o.m2, o.m0G);
  }
  // depth-first decay order (reference MProcess::ConstructDecayKinematics recursion, MProcess.cc:1231-1284)
  P.n_decays = 0;
  std::vector<int> stack;
  for (int i = d.n_legs - 1; i >= 0; --i) stack.push_back(i);
  while (!stack.empty()) {
    const int n = stack.back();
    stack.pop_back();
    if (d.nodes[n].n_children) {
      P.decay_order[P.n_decays++] = n;
      for (int c = d.nodes[n].n_children - 1; c >= 0; --c) stack.push_back(d.nodes[n].first_child + c);
    }
  }
  return GRB_OK;
}

static cudaEvent_t next_event(grb_engine *e) {
  if (e->n_evk >= (int)e->evk.size()) {
    cudaEvent_t ev;
    cudaEventCreate(&ev);
    e->evk.push_back(ev);
  }
  return e->evk[e->n_evk++];
}

// ---- NCCL, resolved at run time so that the library neither links a second NCCL into a torch process nor
// needs one for single-GPU use -------------------------------------------------------------------------------
namespace {
struct NcclId {  // ncclUniqueId (nccl.h: 128 opaque bytes), passed by value to ncclCommInitRank
  char b[128];
};
struct NcclApi {
  void *h = nullptr;
  int (*GetUniqueId)(void *)                                                     = nullptr;
  int (*CommInitRank)(void **, int, NcclId, int)                                 = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
  int (*CommDestroy)(void *)                                                     = nullptr;
  const char *(*GetErrorString)(int)                                             = nullptr;
  bool ok = false;
};
NcclApi &nccl() {
  static NcclApi a;
  if (a.h) return a;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    a.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (a.h) break;
  }
  if (!a.h) return a;
  a.GetUniqueId    = (decltype(a.GetUniqueId))dlsym(a.h, "ncclGetUniqueId");
  a.CommInitRank   = (decltype(a.CommInitRank))dlsym(a.h, "ncclCommInitRank");
  a.AllReduce      = (decltype(a.AllReduce))dlsym(a.h, "ncclAllReduce");
  a.CommDestroy    = (decltype(a.CommDestroy))dlsym(a.h, "ncclCommDestroy");
  a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.h, "ncclGetErrorString");
  a.ok = a.GetUniqueId && a.CommInitRank && a.AllReduce && a.CommDestroy;
  return a;
}
constexpr int kNcclFloat64 = 8, kNcclSum = 0, kNcclMax = 2;  // ncclDataType_t / ncclRedOp_t values (nccl.h)
constexpr int kAccSum = 10;                                   // leading sum-reduced doubles of grb_accum
}  // namespace

static double hfact(int n) {
  double x = 1;
  for (int i = 1; i <= n; ++i) x *= i;
  return x;
}

// grb_resonance[] -> DevRes (host libm for the per-resonance constants)
static int digest_res(const grb_process_desc &d, DevRes &R, std::string &err) {
  memset(&R, 0, sizeof(R));
  if (d.n_res < 0 || d.n_res > GRB_MAX_RES) { err = "bad n_res"; return GRB_ERR_ARG; }
  const bool tensor_res = d.proc == GRB_PROC_PP_RESTENSOR || d.proc == GRB_PROC_PP_RES_CONTENSOR;
  if (tensor_res && !(d.nodes[0].spinX2 == 0 && d.nodes[1].spinX2 == 0 && d.n_nodes == 2 && d.n_legs == 2)) {
    err = "PP[RESTENSOR] / PP[RES+CONTENSOR]: two pseudoscalar daughters have a device amplitude (vector / photon pairs do not)";
    return GRB_ERR_UNSUPPORTED;
  }
  if ((d.proc == GRB_PROC_PP_RES || d.proc == GRB_PROC_PP_RES_CON) && d.n_res > 0 &&
      !(d.nodes[0].spinX2 == 0 && d.nodes[1].spinX2 == 0 && d.n_nodes == 2)) {
    err = "resonances: only two spinless final-state daughters are on the device path";
    return GRB_ERR_UNSUPPORTED;
  }
  R.n = d.n_res;
  for (int i = 0; i < d.n_res; ++i) {
    const grb_resonance &s = d.res[i];
    DevResonance &       o = R.r[i];
    if (s.spinX2 < 0 || s.spinX2 > 4 || (s.spinX2 % 2) != 0) { err = "resonance spin must be J = 0, 1 or 2 on the device path"; return GRB_ERR_UNSUPPORTED; }
    if (s.T_rows != 1 || s.T_cols != 1) { err = "resonance decay T matrix must be 1x1 (spinless daughters)"; return GRB_ERR_UNSUPPORTED; }
    if (tensor_res) {   // MTensorPomeron::ME3: J^P = 0+ and J = 2 (src/MTensorPomeron.cc:232-270,414-560)
      const int J = (int)(s.spinX2 / 2.0);
      if (!((J == 0 && s.P == 1) || J == 2)) {
        err = "MTensorPomeron::ME3: J^P = 0+ and J = 2 resonances have a device amplitude (1- photoproduction and 0- -> gamma gamma do not)";
        return GRB_ERR_UNSUPPORTED;
      }
      tres::ResPar &t = R.tp[i];
      t.J = J; t.on = 0; t.M0sq = s.mass * s.mass; t.M0G = s.mass * s.width; t.gd = s.g_decay_tensor[0];
      for (int k = 0; k < 7; ++k) {
        t.g[k] = s.g_Tensor[k];
        if (k >= 1 && std::abs(s.g_Tensor[k]) > 1e-6) t.on |= 1 << k;   // EPSILON of iG_PPT_total, :1660-1672
      }
    }
    o.spinX2 = s.spinX2; o.P = s.P; o.BW = s.BW; o.spingen = s.spingen; o.spindec = s.spindec; o.frame = s.frame;
    o.photo = ((int)(s.spinX2 / 2.0) == 1 && s.P == -1) ? 1 : 0;
    o.mass = s.mass; o.width = s.width; o.M2 = s.mass * s.mass; o.g_re = s.g_re; o.g_im = s.g_im;
    const double g2 = s.g_FF * s.g_FF;
    o.inv_gFF4 = 1.0 / (g2 * g2);
    o.g_decay = s.g_decay;
    o.B0 = std::pow(std::max(0.7 * 0.7, o.M2) / 2.0, -GRB_PI / 2.0) + 3 * GRB_PI / 2.0;
    for (int k = 0; k < 9; ++k) { o.T_re[k] = s.T_re[k]; o.T_im[k] = s.T_im[k]; }
    for (int k = 0; k < 25; ++k) { o.rho_re[k] = s.rho_re[k]; o.rho_im[k] = s.rho_im[k]; }
    if (o.photo) R.any_photo = 1;
    if (o.BW == 3) R.bw3 = 1;
  }
  for (int J = 1; J <= 2; ++J)
    for (int i = 0; i <= 2 * J; ++i)
      for (int j = 0; j <= 2 * J; ++j) {
        const int m = i - J, mp = j - J;
        const double factor = std::sqrt(hfact(J + m) * hfact(J - m) * hfact(J + mp) * hfact(J - mp));
        for (int k = 0; k <= 2 * J; ++k) {
          double c = 0.0;
          if (J - mp - k >= 0 && J + m - k >= 0 && k + mp - m >= 0)
            c = factor * std::pow(-1.0, mp - m + k) / (hfact(J - mp - k) * hfact(J + m - k) * hfact(k + mp - m) * hfact(k));
          R.wd[J - 1][(i * 5 + j) * 5 + k] = c;
        }
      }
  return GRB_OK;
}

// call with const_slot(e->device).mu held; release_constants() after the last kernel of the call was enqueued
// exchange helpers of grb_allreduce: tail slots [nd + 2 r, nd + 2 r + 1] of the accumulator buffer belong to rank r
__global__ void k_park_maxima(double *acc, int nd, int imax, int rank, int nranks) {
  const int t = threadIdx.x;
  for (int k = t; k < 2 * nranks; k += 32) acc[nd + k] = (k >> 1) == rank ? acc[imax + (k & 1)] : 0.0;
}
__global__ void k_fold_maxima(double *acc, int nd, int imax, int nranks) {
  if (threadIdx.x < 2) {
    double m = 0.0;
    for (int r = 0; r < nranks; ++r) m = fmax(m, acc[nd + 2 * r + threadIdx.x]);
    acc[imax + threadIdx.x] = m;
  }
}

static int ensure_constants(grb_engine *e) {
  ConstSlot &c = const_slot(e->device);
  if (c.owner != e || e->const_dirty) {
    if (c.owner != e && c.last_use) CK(cudaStreamWaitEvent(e->stream, c.last_use, 0));
    CK(upload_proc(e->dproc, e->stream));
    CK(upload_res(e->dres, e->stream));
    if (e->have_loop) CK(upload_loop(e->dloop, e->stream));
    c.owner        = e;
    e->const_dirty = false;
  }
  return GRB_OK;
}
static int release_constants(grb_engine *e) {
  ConstSlot &c = const_slot(e->device);
  if (!c.last_use) CK(cudaEventCreateWithFlags(&c.last_use, cudaEventDisableTiming));
  CK(cudaEventRecord(c.last_use, e->stream));
  return GRB_OK;
}

extern "C" {

int grb_abi_sizeof(int what) {
  switch (what) {
    case 0: return (int)sizeof(grb_process_desc);
    case 1: return (int)sizeof(grb_accum);
    case 2: return (int)sizeof(grb_resonance);
    case 3: return (int)sizeof(grb_node);
    default: return -1;
  }
}

int grb_create(const grb_process_desc *desc, int device, grb_engine **out) {
  if (!desc || !out) return GRB_ERR_ARG;
  *out          = nullptr;
  grb_engine *e = new grb_engine();
  e->desc       = *desc;
  e->device     = device;
  int rc        = digest(*desc, e->dproc, e->err);
  if (rc == GRB_OK) rc = digest_res(*desc, e->dres, e->err);
  if (rc != GRB_OK) {
    g_last_global_error = e->err;
    delete e;
    return rc;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= device) {
    g_last_global_error = "no CUDA device " + std::to_string(device) + " (the engine has no CPU fallback)";
    delete e;
    return GRB_ERR_CUDA;
  }
  auto fail = [&](const char *what, cudaError_t ce) {
    g_last_global_error = std::string(what) + ": " + cudaGetErrorString(ce);
    grb_destroy(e);
    return GRB_ERR_CUDA;
  };
  cudaError_t ce;
  if ((ce = cudaSetDevice(device)) != cudaSuccess) return fail("cudaSetDevice", ce);
  cudaDeviceProp prop;
  if ((ce = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail("cudaGetDeviceProperties", ce);
  e->sm_count = prop.multiProcessorCount;
  if ((ce = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", ce);
  cudaEventCreate(&e->ev0);
  cudaEventCreate(&e->ev1);
  if ((ce = cudaStreamCreateWithFlags(&e->stream_in, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", ce);
  if ((ce = cudaStreamCreateWithFlags(&e->stream_out, cudaStreamNonBlocking)) != cudaSuccess) return fail("cudaStreamCreate", ce);
  cudaEventCreateWithFlags(&e->ev_in, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&e->ev_done, cudaEventDisableTiming);
  const char *env = getenv("GRB_CHUNK");
  e->chunk        = env ? std::max<int64_t>(1024, atoll(env)) : (int64_t)1 << 23;
  // k_points / k_screen overlap: GRB_OVERLAP = CTAs per SM left to k_screen; default off -- measured on B200 (profiles/
  // r02_overlap_sweep.txt): k_screen needs all 16 warps per SM (29.3 instead of 26.7 ms per launch at 3 CTAs per SM), which
  // costs more than the 2 ms of k_points the overlap hides
  const char *ov  = getenv("GRB_OVERLAP");
  e->overlap      = ov ? std::max(0, atoi(ov)) : 0;
  const char *hs  = getenv("GRB_HOST_SLICE");  // points per pipeline slice of grb_eval with host buffers
  e->host_slice   = hs ? std::max<int64_t>(1024, atoll(hs)) : (int64_t)1 << 20;
  if (desc->screening) {
    if ((ce = cudaMalloc(&e->d_items, sizeof(double) * ITEM_NF * e->chunk)) != cudaSuccess) return fail("cudaMalloc items", ce);
    if ((ce = cudaMalloc(&e->d_item_idx, sizeof(int32_t) * e->chunk)) != cudaSuccess) return fail("cudaMalloc", ce);
    if ((ce = cudaMalloc(&e->d_item_bins, sizeof(uint64_t) * e->chunk)) != cudaSuccess) return fail("cudaMalloc", ce);
    if (desc->lipsdim > 8 && (ce = cudaMalloc(&e->d_item_bins_hi, sizeof(uint64_t) * e->chunk * (desc->lipsdim > 16 ? 2 : 1))) != cudaSuccess) return fail("cudaMalloc", ce);
  }
  if ((ce = cudaMalloc(&e->d_item_count, sizeof(unsigned int) * 8)) != cudaSuccess) return fail("cudaMalloc", ce);   // [0..2] work list, [4] k_points tiles
  // (+ 2 x 64 doubles: the per-rank maxima slots of grb_allreduce)
  if ((ce = cudaMalloc(&e->d_acc, sizeof(grb_accum) + 128 * sizeof(double))) != cudaSuccess) return fail("cudaMalloc", ce);
  if ((ce = cudaMalloc(&e->d_xmat, sizeof(double) * GRB_MAX_BINS * GRB_MAX_DIM)) != cudaSuccess) return fail("cudaMalloc", ce);
  cudaMemsetAsync(e->d_acc, 0, sizeof(grb_accum), e->stream);
  cudaMemsetAsync(e->d_item_count, 0, sizeof(unsigned int) * 8, e->stream);
  *out = e;
  return GRB_OK;
}

int grb_create_from_card(const char *card_path, const char *process_override, int pomloop_override,
                         const char *graniitti_base, int device, grb_engine **out) {
  if (!out) return GRB_ERR_ARG;
  grb_process_desc d;
  char             buf[1024];
  int rc = grb_parse_card(card_path, process_override, pomloop_override, graniitti_base, &d, buf, sizeof(buf));
  if (rc != GRB_OK) {
    g_last_global_error = buf;
    return rc;
  }
  return grb_create(&d, device, out);
}

int grb_get_desc(const grb_engine *e, grb_process_desc *out) {
  if (!e || !out) return GRB_ERR_ARG;
  *out = e->desc;
  return GRB_OK;
}

const char *grb_last_error(const grb_engine *e) { return e ? e->err.c_str() : g_last_global_error.c_str(); }

void grb_destroy(grb_engine *e) {
  if (!e) return;
  cudaSetDevice(e->device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  {
    ConstSlot &c = const_slot(e->device);
    std::lock_guard<std::mutex> lk(c.mu);
    if (c.owner == e) c.owner = nullptr;
  }
  cudaFree(e->d_items); cudaFree(e->d_item_idx); cudaFree(e->d_item_bins); cudaFree(e->d_item_bins_hi); cudaFree(e->d_item_count);
  cudaFree(e->d_items2); cudaFree(e->d_item_idx2); cudaFree(e->d_item_bins2); cudaFree(e->d_item_bins_hi2); cudaFree(e->d_item_count2);
  for (int k = 0; k < 2; ++k) { if (e->ev_pts[k]) cudaEventDestroy(e->ev_pts[k]); if (e->ev_free[k]) cudaEventDestroy(e->ev_free[k]); }
  if (e->ev_start) cudaEventDestroy(e->ev_start);
  if (e->stream_pts) cudaStreamDestroy(e->stream_pts);
  cudaFree(e->d_acc); cudaFree(e->d_xmat); cudaFree(e->d_scratch); cudaFree(e->d_max2);
  cudaFree(e->d_ev); cudaFree(e->d_pay); cudaFree(e->d_ev_count);
  if (e->nccl_comm && nccl().ok) nccl().CommDestroy(e->nccl_comm);
  for (auto ev : e->evk) cudaEventDestroy(ev);
  if (e->ev0) cudaEventDestroy(e->ev0);
  if (e->ev1) cudaEventDestroy(e->ev1);
  if (e->stream_in) cudaStreamDestroy(e->stream_in);
  if (e->stream_out) cudaStreamDestroy(e->stream_out);
  if (e->ev_in) cudaEventDestroy(e->ev_in);
  if (e->ev_done) cudaEventDestroy(e->ev_done);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

// Loop discretisation + A_eik(kt_j^2): everything S3ScreenedAmp2 needs that depends on (i,j) only
// (reference src/MProcess.cc:187-231, src/MEikonal.cc:489-512, include/Graniitti/MMath.h:502-516).
int grb_set_eikonal(grb_engine *e, const double *kt2, const double *re, const double *im, int n_points) {
  if (!e || !kt2 || !re || !im || n_points < 3) return GRB_ERR_ARG;
  const grb_process_desc &d = e->desc;
  const int NKT = d.loop_nkt, NPHI = d.loop_nphi;
  if (NKT < 3 || NPHI < 3 || NKT % 3 || NPHI % 3 || NKT + 1 > GRB_MAX_LOOP_KT || NPHI + 1 > GRB_MAX_LOOP_PHI) {
    e->err = "loop discretisation must be multiples of 3 within the compiled bounds";
    return GRB_ERR_ARG;
  }
  DevLoop &Q = e->dloop;
  memset(&Q, 0, sizeof(Q));
  Q.nkt = NKT; Q.nphi = NPHI;
  // IArray1D::Interpolate1D works on the array's own uniform grid (MIN, MAX, STEP = (MAX - MIN) / N; MEikonal.cc:489-512,
  // defaults MEikonalNumerics::MinKT2 = 1e-6, MaxKT2 = 25, MEikonal.h:27-28): taken from the table, which must be uniform
  const int    N    = n_points - 1;
  const double MIN = kt2[0], MAX = kt2[N];
  const double STEP = (MAX - MIN) / N;
  if (!(MAX > MIN) || !(MIN >= 0.0)) { e->err = "eikonal table: kt^2 grid must be increasing and non-negative"; return GRB_ERR_ARG; }
  for (int i = 0; i <= N; ++i)
    if (std::fabs(kt2[i] - (MIN + i * STEP)) > 1e-9 * MAX) {
      e->err = "eikonal table: kt^2 grid is not uniform (log-spaced tables are outside the device path)";
      return GRB_ERR_ARG;
    }
  const double StepKT  = (d.loop_max_kt - d.loop_min_kt) / NKT;
  const double StepPhi = (2.0 * GRB_PI - 0.0) / NPHI;
  auto w38 = [](int i, int n) { return (i == 0 || i == n) ? 1.0 : ((i % 3 == 0) ? 2.0 : 3.0); };
  for (int j = 0; j <= NKT; ++j) {
    const double kt = d.loop_min_kt + j * StepKT;
    double       a  = kt * kt;
    if (a < MIN) a = MIN;
    if (a > MAX * (1 + 1e-5)) { e->err = "loop kt^2 outside the eikonal table"; return GRB_ERR_ARG; }
    int i = (int)std::floor((a - MIN) / STEP);
    if (i < 0) i = 0;
    if (i >= N) i = N - 1;
    const double dx = a - kt2[i], h = kt2[i + 1] - kt2[i];
    Q.kt[j]      = kt;
    Q.aeik_re[j] = re[i] + dx * ((re[i + 1] - re[i]) / h);
    Q.aeik_im[j] = im[i] + dx * ((im[i + 1] - im[i]) / h);
    Q.wkt[j]     = w38(j, NKT);
    Q.b_re[j]    = Q.aeik_re[j] * kt * Q.wkt[j];
    Q.b_im[j]    = Q.aeik_im[j] * kt * Q.wkt[j];
  }
  for (int i = 0; i <= NPHI; ++i) {
    const double phi = 0.0 + i * StepPhi;
    Q.cphi[i] = std::cos(phi);
    Q.sphi[i] = std::sin(phi);
    Q.wphi[i] = w38(i, NPHI);
  }
  Q.norm = (9 * (StepPhi * StepKT) / 64) * (1.0 / (8.0 * (GRB_PI * GRB_PI) * d.s));
  // phi = 0 and phi = 2 pi are the same node up to sin(2 pi) = -2.4e-16: visit it once with both Simpson
  // end weights (1 + 1).  Saves 1/(NPHI+1) of the loop; GRB_NO_PHI_FOLD=1 keeps the reference's N+1 nodes.
  Q.nphi_nodes = NPHI + 1;
  if (!getenv("GRB_NO_PHI_FOLD")) {
    Q.wphi[0] += Q.wphi[NPHI];
    Q.nphi_nodes = NPHI;
  }
  e->have_loop   = true;
  e->const_dirty = true;  // force re-upload
  return GRB_OK;
}

static bool is_device_ptr(const void *p) {
  if (!p) return false;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// rng: internal draws from the Philox stream (seed, iteration, first_index + i) instead of a caller tape
static int eval_impl(grb_engine *e, int64_t n, const double *u, int dim, const double *tape, int dtape, bool rng, uint64_t seed,
                     uint64_t iteration, int64_t first_index, double *w, uint8_t *flags, double *dump) {
  if (!e || n < 0 || !u || !w || !flags || dim < e->desc.lipsdim || dtape < 0 || (dtape > 0 && !tape)) return GRB_ERR_ARG;
  if (e->desc.screening && !e->have_loop) { e->err = "screening is on but no eikonal table was set"; return GRB_ERR_STATE; }
  if (n == 0) return GRB_OK;
  CK(cudaSetDevice(e->device));
  const bool dev_in = is_device_ptr(u), dev_out = is_device_ptr(w);
  if ((tape && is_device_ptr(tape) != dev_in) || is_device_ptr(flags) != dev_out || (dump && is_device_ptr(dump) != dev_out)) {
    e->err = "inputs (u,tape) must live on the same side, and outputs (w,flags,dump) likewise";
    return GRB_ERR_ARG;
  }
  // stage host buffers
  const size_t b_u = sizeof(double) * dim * n, b_t = sizeof(double) * (size_t)dtape * n, b_w = sizeof(double) * n;
  const size_t b_f = (size_t)((n + 7) / 8 * 8), b_d = dump ? sizeof(double) * GRB_DUMP_LEN * n : 0;
  size_t       need = 0;
  if (!dev_in) need += b_u + b_t;
  if (!dev_out) need += b_w + b_f + b_d;
  if (need > e->scratch_bytes) {
    cudaFree(e->d_scratch);
    e->d_scratch = nullptr;
    e->scratch_bytes = 0;
    CK(cudaMalloc(&e->d_scratch, need));
    e->scratch_bytes = need;
  }
  char *        sp = (char *)e->d_scratch;
  const double *du = u, *dt = tape;
  double *      dw = w, *dd = dump;
  uint8_t *     df = flags;
  if (!dev_in) {
    du = (double *)sp; sp += b_u;
    dt = dtape ? (double *)sp : nullptr; sp += b_t;
  }
  if (!dev_out) {
    dw = (double *)sp; sp += b_w;
    df = (uint8_t *)sp; sp += b_f;
    dd = dump ? (double *)sp : nullptr;
  }
  std::unique_lock<std::mutex> const_lock(const_slot(e->device).mu);
  int rc = ensure_constants(e);
  if (rc != GRB_OK) return rc;
  e->last_launches = 0;
  e->n_evk = 0;
  // Host buffers are pipelined: slice k+1 is copied in (stream_in) and slice k-1 copied out (stream_out, the other DMA
  // engine) while slice k computes.  The SoA rows keep their leading dimension n on the device, so a slice is a 2-D copy
  // of `len` columns.  Device-resident callers run whole chunks on the compute stream as before.
  const bool    staged = !dev_in || !dev_out;
  const int64_t step   = staged ? std::min<int64_t>(e->chunk, e->host_slice) : e->chunk;
  const size_t  pitch  = sizeof(double) * (size_t)n;
  CK(cudaEventRecord(e->ev0, e->stream));
  // The first copy-in cannot overlap anything: the pipeline is started with a quarter and a half slice
  int64_t len = 0;
  int     k_slice = 0;
  for (int64_t off = 0; off < n; off += len, ++k_slice) {
    len = std::min<int64_t>((!dev_in && k_slice < 2 && step >= 4096) ? step >> (2 - k_slice) : step, n - off);
    if (!dev_in) {
      CK(cudaMemcpy2DAsync((void *)(du + off), pitch, u + off, pitch, sizeof(double) * len, dim, cudaMemcpyHostToDevice, e->stream_in));
      if (dtape) CK(cudaMemcpy2DAsync((void *)(dt + off), pitch, tape + off, pitch, sizeof(double) * len, dtape, cudaMemcpyHostToDevice, e->stream_in));
      CK(cudaEventRecord(e->ev_in, e->stream_in));
      CK(cudaStreamWaitEvent(e->stream, e->ev_in, 0));
    }
    PointArgs a;
    memset(&a, 0, sizeof(a));
    a.n = len;
    a.u = du + off; a.tape = dt ? dt + off : nullptr; a.ld = n; a.dim = dim; a.dtape = dtape;
    a.vegas = 0; a.bins = 0;
    a.tape_rng = rng ? 1 : 0; a.seed = seed; a.iter = (uint32_t)iteration; a.first_call = first_index + off;
    a.w = dw + off; a.flags = df + off; a.dump = dd ? dd + off : nullptr;
    a.items = e->d_items; a.item_idx = e->d_item_idx; a.item_bins = e->d_item_bins; a.item_bins_hi = e->d_item_bins_hi; a.item_count = e->d_item_count;
    a.item_cap = e->chunk; a.acc = e->d_acc;
    a.tile_count = e->d_item_count + 4;
    CK(cudaMemsetAsync(e->d_item_count, 0, 8 * sizeof(unsigned int), e->stream));
    e->last_launches += launch_points(a, e->desc.n_nodes, e->dproc.amp_kind, e->sm_count, e->stream);
    if (e->desc.screening) {
      CK(cudaEventRecord(next_event(e), e->stream));
      e->last_launches += launch_screen(a, e->dproc.amp_kind, regge_fast_ff(e->dproc), e->sm_count, e->stream);
      CK(cudaEventRecord(next_event(e), e->stream));
    }
    if (!dev_out) {
      CK(cudaEventRecord(e->ev_done, e->stream));
      CK(cudaStreamWaitEvent(e->stream_out, e->ev_done, 0));
      CK(cudaMemcpyAsync(w + off, dw + off, sizeof(double) * len, cudaMemcpyDeviceToHost, e->stream_out));
      CK(cudaMemcpyAsync(flags + off, df + off, (size_t)len, cudaMemcpyDeviceToHost, e->stream_out));
      if (dump) CK(cudaMemcpy2DAsync(dump + off, pitch, dd + off, pitch, sizeof(double) * len, GRB_DUMP_LEN, cudaMemcpyDeviceToHost, e->stream_out));
    }
  }
  CK(cudaEventRecord(e->ev1, e->stream));
  rc = release_constants(e);
  const_lock.unlock();
  if (rc != GRB_OK) return rc;
  e->timed = true;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(e->stream));
  if (!dev_out) CK(cudaStreamSynchronize(e->stream_out));
  return GRB_OK;
}

int grb_eval(grb_engine *e, int64_t n, const double *u, int dim, const double *tape, int dtape, double *w,
             uint8_t *flags, double *dump) {
  return eval_impl(e, n, u, dim, tape, dtape, false, 0, 0, 0, w, flags, dump);
}

int grb_eval_rng(grb_engine *e, int64_t n, const double *u, int dim, uint64_t seed, uint64_t iteration, int64_t first_index,
                 double *w, uint8_t *flags, double *dump) {
  return eval_impl(e, n, u, dim, nullptr, 0, true, seed, iteration, first_index, w, flags, dump);
}

int grb_vegas_iterate_async(grb_engine *e, const double *xmat, int bins, int dim, int64_t first_call,
                            int64_t ncalls, uint64_t seed, uint64_t iteration, int mode, double maxf,
                            void **accum_device_ptr) {
  if (!e || !xmat || bins < 2 || bins > GRB_MAX_BINS || dim != e->desc.lipsdim || ncalls < 0) return GRB_ERR_ARG;
  if (e->desc.screening && !e->have_loop) { e->err = "screening is on but no eikonal table was set"; return GRB_ERR_STATE; }
  CK(cudaSetDevice(e->device));
  const int npart = 3 + e->desc.n_nodes;
  if (mode == GRB_MODE_GENERATE) {
    if (!(maxf > 0.0)) { e->err = "generation mode needs maxf > 0 (from the integration phase)"; return GRB_ERR_ARG; }
    if (!e->d_ev) {
      const char *ec = getenv("GRB_EVENT_CAP");
      e->ev_cap = ec ? std::max<int64_t>(1024, atoll(ec)) : (int64_t)1 << 22;
      CK(cudaMalloc(&e->d_ev, sizeof(double) * (4 * npart + 2) * e->ev_cap));
      CK(cudaMalloc(&e->d_ev_count, 2 * sizeof(unsigned int)));
    }
    // (screening may have been switched on by grb_update_desc after the first generation call)
    if (e->desc.screening && !e->d_pay) CK(cudaMalloc(&e->d_pay, sizeof(double) * 4 * npart * e->chunk));
    CK(cudaMemsetAsync(e->d_ev_count, 0, 2 * sizeof(unsigned int), e->stream));
  }
  std::unique_lock<std::mutex> const_lock(const_slot(e->device).mu);
  int rc = ensure_constants(e);
  if (rc != GRB_OK) return rc;
  CK(cudaMemcpyAsync(e->d_xmat, xmat, sizeof(double) * bins * dim, cudaMemcpyHostToDevice, e->stream));
  CK(cudaMemsetAsync(e->d_acc, 0, sizeof(grb_accum), e->stream));
  e->last_launches = 0;
  e->n_evk = 0;
  CK(cudaEventRecord(e->ev0, e->stream));
  // Optional (GRB_OVERLAP=n): k_points of chunk c+1 on a side stream beside k_screen of chunk c (two work lists), k_screen
  // keeping n of its 4 resident CTAs per SM and k_points one 128-thread CTA per SM in the freed part of the register file.
  const bool overlap = e->overlap > 0 && e->desc.screening && mode != GRB_MODE_GENERATE && ncalls > e->chunk;
  if (overlap && !e->d_items2) {
    CK(cudaMalloc(&e->d_items2, sizeof(double) * ITEM_NF * e->chunk));
    CK(cudaMalloc(&e->d_item_idx2, sizeof(int32_t) * e->chunk));
    CK(cudaMalloc(&e->d_item_bins2, sizeof(uint64_t) * e->chunk));
    if (e->desc.lipsdim > 8) CK(cudaMalloc(&e->d_item_bins_hi2, sizeof(uint64_t) * e->chunk * (e->desc.lipsdim > 16 ? 2 : 1)));
    CK(cudaMalloc(&e->d_item_count2, sizeof(unsigned int) * 8));
    CK(cudaStreamCreateWithFlags(&e->stream_pts, cudaStreamNonBlocking));
    for (int k = 0; k < 2; ++k) {
      CK(cudaEventCreateWithFlags(&e->ev_pts[k], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&e->ev_free[k], cudaEventDisableTiming));
    }
    CK(cudaEventCreateWithFlags(&e->ev_start, cudaEventDisableTiming));
  }
  if (overlap) {  // the side stream starts after the constants, the grid and the cleared accumulators are in place
    CK(cudaEventRecord(e->ev_start, e->stream));
    CK(cudaStreamWaitEvent(e->stream_pts, e->ev_start, 0));
  }
  int64_t chunk_no = 0;
  for (int64_t off = 0; off < ncalls; off += e->chunk, ++chunk_no) {
    const int b = overlap ? (int)(chunk_no & 1) : 0;
    PointArgs a;
    memset(&a, 0, sizeof(a));
    a.n = std::min<int64_t>(e->chunk, ncalls - off);
    a.ld = a.n; a.dim = dim;
    a.vegas = 1; a.bins = bins; a.mode = mode; a.xmat = e->d_xmat; a.seed = seed; a.iter = (uint32_t)iteration;
    a.first_call = first_call + off; a.maxf = maxf;
    a.ev_out = e->d_ev; a.ev_count = e->d_ev_count; a.ev_cap = e->ev_cap; a.npart = npart; a.pay = e->d_pay;
    if (b == 0) { a.items = e->d_items; a.item_idx = e->d_item_idx; a.item_bins = e->d_item_bins; a.item_bins_hi = e->d_item_bins_hi; a.item_count = e->d_item_count; }
    else { a.items = e->d_items2; a.item_idx = e->d_item_idx2; a.item_bins = e->d_item_bins2; a.item_bins_hi = e->d_item_bins_hi2; a.item_count = e->d_item_count2; }
    a.item_cap = e->chunk; a.acc = e->d_acc;
    cudaStream_t sp = overlap ? e->stream_pts : e->stream;
    if (overlap) {
      a.side_launch = chunk_no > 0;          // the first chunk has the device to itself
      a.screen_ctas = e->overlap;
      if (chunk_no >= 2) CK(cudaStreamWaitEvent(sp, e->ev_free[b], 0));   // k_finish of chunk c-2 has released this list
    }
    a.tile_count = a.item_count + 4;
    CK(cudaMemsetAsync(a.item_count, 0, 8 * sizeof(unsigned int), sp));
    e->last_launches += launch_points(a, e->desc.n_nodes, e->dproc.amp_kind, e->sm_count, sp);
    if (overlap) {
      CK(cudaEventRecord(e->ev_pts[b], sp));
      CK(cudaStreamWaitEvent(e->stream, e->ev_pts[b], 0));
    }
    if (e->desc.screening) {
      CK(cudaEventRecord(next_event(e), e->stream));
      e->last_launches += launch_screen(a, e->dproc.amp_kind, regge_fast_ff(e->dproc), e->sm_count, e->stream);
      CK(cudaEventRecord(next_event(e), e->stream));
      if (overlap) CK(cudaEventRecord(e->ev_free[b], e->stream));
    }
  }
  CK(cudaEventRecord(e->ev1, e->stream));
  rc = release_constants(e);
  const_lock.unlock();
  if (rc != GRB_OK) return rc;
  e->timed = true;
  CK(cudaGetLastError());
  if (accum_device_ptr) *accum_device_ptr = e->d_acc;
  return GRB_OK;
}

int grb_vegas_fetch(grb_engine *e, grb_accum *out) {
  if (!e || !out) return GRB_ERR_ARG;
  CK(cudaSetDevice(e->device));
  CK(cudaMemcpyAsync(out, e->d_acc, sizeof(grb_accum), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  // points that are not Valid() keep amplitude_ok = true in the reference (AuxIntData default)
  out->n_amp_ok = out->evaluations - out->n_nan_inf;
  if (e->d_ev_count) {
    unsigned int c[2] = {0, 0};
    CK(cudaMemcpy(c, e->d_ev_count, sizeof(c), cudaMemcpyDeviceToHost));
    out->n_accepted = (double)c[0];
    out->n_overflow = (double)c[1];
    e->ev_dropped   = std::max<int64_t>(0, (int64_t)c[0] - e->ev_cap);
  }
  return GRB_OK;
}

int grb_vegas_iterate(grb_engine *e, const double *xmat, int bins, int dim, int64_t first_call, int64_t ncalls,
                      uint64_t seed, uint64_t iteration, int mode, double maxf, grb_accum *out) {
  int rc = grb_vegas_iterate_async(e, xmat, bins, dim, first_call, ncalls, seed, iteration, mode, maxf, nullptr);
  if (rc != GRB_OK) return rc;
  return grb_vegas_fetch(e, out);
}

int grb_sync(grb_engine *e) {
  if (!e) return GRB_ERR_ARG;
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize(e->stream));
  return GRB_OK;
}

// VEGAS grid refinement (Lepage 1978; the damped-redistribution scheme of reference
// include/Graniitti/MVEGAS.h:87-142): 3-point smoothing of the per-bin <f^2>, importance
// r_i = ((1 - d_i/D) / (ln D - ln d_i))^lambda, then bin edges are moved so every new bin holds sum(r)/bins.
int grb_vegas_optimize_grid(double *xmat, const double *f2mat, int bins, int dim, double lambda) {
  if (!xmat || !f2mat || bins < 3 || bins > GRB_MAX_BINS || dim < 1 || dim > GRB_MAX_DIM) return GRB_ERR_ARG;
  const double        EPS = 1.0e-30;
  std::vector<double> d(bins), r(bins), xnew(bins);
  for (int j = 0; j < dim; ++j) {
    auto F = [&](int i) { return f2mat[i * GRB_MAX_DIM + j]; };
    // smoothing
    double total = 0.0;
    double zo = F(0), zn = F(1);
    d[0]  = (zo + zn) / 2.0;
    total = d[0];
    for (int i = 2; i < bins; ++i) {
      const double ac = zo + zn;
      zo = zn;
      zn = F(i);
      d[i - 1] = (ac + zn) / 3.0;
      total += d[i - 1];
    }
    d[bins - 1] = (zo + zn) / 2.0;
    total += d[bins - 1];
    // importance
    double rsum = 0.0;
    for (int i = 0; i < bins; ++i) {
      if (d[i] < EPS) d[i] = EPS;
      r[i] = std::pow((1.0 - d[i] / total) / (std::log(total) - std::log(d[i]) + EPS), lambda);
      if (std::isnan(r[i]) || std::isinf(r[i])) r[i] = EPS;
      rsum += r[i];
    }
    // rebin
    const double per = rsum / bins;
    int          k  = 0;
    double       dr = 0.0, xo = 0.0, xn = 0.0;
    for (int i = 0; i < bins - 1; ++i) {
      while (per > dr) dr += r[(++k) - 1];
      if (k > 1) xo = xmat[(k - 2) * dim + j];
      xn = xmat[(k - 1) * dim + j];
      dr -= per;
      xnew[i] = xn - (xn - xo) * dr / r[k - 1];
    }
    for (int i = 0; i < bins - 1; ++i) xmat[i * dim + j] = xnew[i];
    xmat[(bins - 1) * dim + j] = 1.0;
  }
  return GRB_OK;
}

int64_t grb_get_events(grb_engine *e, int64_t max_events, double *momenta, double *weights, int *n_particles) {
  if (!e || max_events < 0) return GRB_ERR_ARG;
  const int npart = 3 + e->desc.n_nodes;
  if (n_particles) *n_particles = npart;
  if (!e->d_ev) return 0;
  if (cudaSetDevice(e->device) != cudaSuccess) return GRB_ERR_CUDA;
  if (cudaStreamSynchronize(e->stream) != cudaSuccess) return GRB_ERR_CUDA;  // the engine's stream is non-blocking
  unsigned int c[2] = {0, 0};
  if (cudaMemcpy(c, e->d_ev_count, sizeof(c), cudaMemcpyDeviceToHost) != cudaSuccess) return GRB_ERR_CUDA;
  e->ev_dropped   = std::max<int64_t>(0, (int64_t)c[0] - e->ev_cap);
  const int64_t n = std::min<int64_t>(std::min<int64_t>((int64_t)c[0], e->ev_cap), max_events);
  if (n == 0) return 0;
  // device rows have stride ev_cap; the caller's arrays are [npart][4][n] and [2][n] (W, f)
  if (momenta && cudaMemcpy2D(momenta, sizeof(double) * n, e->d_ev, sizeof(double) * e->ev_cap, sizeof(double) * n,
                              4 * npart, cudaMemcpyDeviceToHost) != cudaSuccess) return GRB_ERR_CUDA;
  if (weights && cudaMemcpy2D(weights, sizeof(double) * n, e->d_ev + (int64_t)4 * npart * e->ev_cap,
                              sizeof(double) * e->ev_cap, sizeof(double) * n, 2, cudaMemcpyDeviceToHost) != cudaSuccess)
    return GRB_ERR_CUDA;
  return n;
}

int64_t grb_write_engine_events(grb_engine *e, const char *path, const grb_writer_opts *opts) {
  if (!e || !path || !opts) return GRB_ERR_ARG;
  const int npart = 3 + e->desc.n_nodes;
  if (!e->d_ev) return 0;
  if (cudaSetDevice(e->device) != cudaSuccess || cudaStreamSynchronize(e->stream) != cudaSuccess) return GRB_ERR_CUDA;
  unsigned int c[2] = {0, 0};
  if (cudaMemcpy(c, e->d_ev_count, sizeof(c), cudaMemcpyDeviceToHost) != cudaSuccess) return GRB_ERR_CUDA;
  const int64_t n = std::min<int64_t>((int64_t)c[0], e->ev_cap);
  if (n == 0) return grb_write_events(&e->desc, path, opts, 0, nullptr, nullptr);
  std::vector<double> mom((size_t)4 * npart * n), wts((size_t)2 * n);
  int np = 0;
  const int64_t got = grb_get_events(e, n, mom.data(), wts.data(), &np);
  if (got < 0) return got;
  return grb_write_events(&e->desc, path, opts, got, mom.data(), wts.data());
}

int64_t grb_events_dropped(const grb_engine *e) { return e ? e->ev_dropped : GRB_ERR_ARG; }

int grb_loop_nodes(const grb_engine *e, int *nkt_nodes, int *nphi_nodes) {
  if (!e) return GRB_ERR_ARG;
  if (!e->have_loop) return GRB_ERR_STATE;
  if (nkt_nodes) *nkt_nodes = e->dloop.nkt + 1;
  if (nphi_nodes) *nphi_nodes = e->dloop.nphi_nodes;
  return GRB_OK;
}

int grb_build_eikonal(grb_engine *e, int number_bt, int number_kt2, double *bt_out, double *omega_re_out,
                      double *omega_im_out, double *kt2_out, double *re_out, double *im_out, double *xs_tot_el_inel) {
  if (!e) return GRB_ERR_ARG;
  if (number_bt <= 0) number_bt = 10000;    // modeldata/NUMERICS.json NUMERICS_EIKONAL defaults
  if (number_kt2 <= 0) number_kt2 = 10000;
  CK(cudaSetDevice(e->device));
  EikonalTables T;
  const int rc = build_eikonal_tables(e->desc.soft, e->desc.s, number_bt, number_kt2, e->stream, T, e->err);
  if (rc > 0) return GRB_ERR_CUDA;
  if (rc < 0) return rc == -3 ? GRB_ERR_UNSUPPORTED : GRB_ERR_ARG;
  auto put = [](double *dst, const std::vector<double> &v) { if (dst) memcpy(dst, v.data(), sizeof(double) * v.size()); };
  put(bt_out, T.bt); put(omega_re_out, T.om_re); put(omega_im_out, T.om_im);
  put(kt2_out, T.kt2); put(re_out, T.sa_re); put(im_out, T.sa_im);
  if (xs_tot_el_inel) for (int k = 0; k < 3; ++k) xs_tot_el_inel[k] = T.xs[k] * 1e3;  // barn -> mb
  return grb_set_eikonal(e, T.kt2.data(), T.sa_re.data(), T.sa_im.data(), (int)T.kt2.size());
}

int grb_update_desc(grb_engine *e, const grb_process_desc *desc) {
  if (!e || !desc) return GRB_ERR_ARG;
  if (desc->n_nodes != e->desc.n_nodes || desc->n_legs != e->desc.n_legs || desc->ps_class != e->desc.ps_class) {
    e->err = "grb_update_desc: the decay-tree topology and the phase-space class must stay the same";
    return GRB_ERR_ARG;
  }
  DevProc P;
  DevRes  R;
  int rc = digest(*desc, P, e->err);
  if (rc == GRB_OK) rc = digest_res(*desc, R, e->err);
  if (rc != GRB_OK) return rc;
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize(e->stream));
  if (desc->screening && !e->d_items) {  // screening switched on: the work list was not allocated at creation
    CK(cudaMalloc(&e->d_items, sizeof(double) * ITEM_NF * e->chunk));
    CK(cudaMalloc(&e->d_item_idx, sizeof(int32_t) * e->chunk));
    CK(cudaMalloc(&e->d_item_bins, sizeof(uint64_t) * e->chunk));
  }
  if (desc->screening && desc->lipsdim > 8 && !e->d_item_bins_hi) CK(cudaMalloc(&e->d_item_bins_hi, sizeof(uint64_t) * e->chunk * (desc->lipsdim > 16 ? 2 : 1)));
  const bool same_loop = desc->s == e->desc.s && desc->loop_nkt == e->desc.loop_nkt && desc->loop_nphi == e->desc.loop_nphi &&
                         desc->loop_min_kt == e->desc.loop_min_kt && desc->loop_max_kt == e->desc.loop_max_kt;
  e->desc  = *desc;
  e->dproc = P;
  e->dres  = R;
  if (!same_loop) e->have_loop = false;   // A_eik(kt^2) and the loop norm depend on s and the discretisation
  e->const_dirty = true;  // force the constant upload on the next call
  return GRB_OK;
}

int grb_set_beam_energies(grb_engine *e, double ebeam1, double ebeam2) {
  if (!e || !(ebeam1 > 0) || !(ebeam2 > 0)) return GRB_ERR_ARG;
  grb_process_desc d = e->desc;
  auto msq = [](double x) { return std::sqrt(std::max(0.0, x)); };
  const double E1 = std::max(ebeam1, 1.001 * d.beam1_mass), E2 = std::max(ebeam2, 1.001 * d.beam2_mass);
  d.pbeam1_pz = msq(E1 * E1 - d.beam1_mass * d.beam1_mass); d.pbeam1_E = E1;
  d.pbeam2_pz = -msq(E2 * E2 - d.beam2_mass * d.beam2_mass); d.pbeam2_E = E2;
  const double sz = d.pbeam1_pz + d.pbeam2_pz, se = d.pbeam1_E + d.pbeam2_E;
  d.s      = se * se - (0.0 * 0.0 + 0.0 * 0.0 + sz * sz);   // (pbeam1 + pbeam2).M2()
  d.sqrt_s = msq(d.s);
  if (d.sqrt_s < d.beam1_mass + d.beam2_mass) { e->err = "MProcess::SetBeamEnergies: Error with input CMS energy"; return GRB_ERR_ARG; }
  // <C> class QED / EW processes: the default kt range is the kinematic maximum (MContinuum.cc:80-83)
  if (d.ps_class == GRB_CLASS_C && d.gcuts.kt_max == e->desc.sqrt_s / 2.0) d.gcuts.kt_max = d.sqrt_s / 2.0;
  return grb_update_desc(e, &d);
}

int grb_xscan(grb_engine *e, const double *sqrt_s, int n, const grb_vegas_param *par, grb_xscan_row *rows) {
  if (!e || !sqrt_s || n < 0 || !par || !rows) return GRB_ERR_ARG;
  for (int i = 0; i < n; ++i) {
    int rc = grb_set_beam_energies(e, sqrt_s[i] / 2, sqrt_s[i] / 2);
    if (rc != GRB_OK) return rc;
    double xs[3] = {0, 0, 0};
    rc = grb_build_eikonal(e, 0, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, xs);
    if (rc != GRB_OK) return rc;
    grb_vegas_result r;
    rc = grb_integrate(e, par, 0, 1, nullptr, nullptr, &r);
    if (rc != GRB_OK) return rc;
    rows[i].sqrt_s = e->desc.sqrt_s;
    rows[i].xs_tot_mb = xs[0]; rows[i].xs_el_mb = xs[1]; rows[i].xs_inel_mb = xs[2];
    rows[i].sigma = r.sigma; rows[i].sigma_err = r.sigma_err; rows[i].chi2 = r.chi2;
    rows[i].evaluations = r.evaluations; rows[i].seconds = r.seconds;
  }
  return GRB_OK;
}

int grb_integrate_batch(grb_engine *e, const grb_process_desc *descs, int k_sets, const grb_vegas_param *par, int rank, int nranks,
                        grb_reduce_fn reduce, void *ctx, grb_vegas_result *results) {
  if (!e || !descs || k_sets < 0 || !par || !results) return GRB_ERR_ARG;
  for (int k = 0; k < k_sets; ++k) {
    int rc = grb_update_desc(e, &descs[k]);
    if (rc != GRB_OK) return rc;
    if (e->desc.screening && !e->have_loop) { e->err = "grb_integrate_batch: set " + std::to_string(k) + " changes s or the loop discretisation: no eikonal table"; return GRB_ERR_STATE; }
    rc = grb_integrate(e, par, rank, nranks, reduce, ctx, &results[k]);
    if (rc != GRB_OK) return rc;
  }
  return GRB_OK;
}

int grb_test_fastmath(int device, int kind, int64_t n, const double *x, double *y) {
  if (!x || !y || n <= 0 || kind < 0 || kind > 11) return GRB_ERR_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return GRB_ERR_CUDA;
  return run_fastmath(kind, n, x, y) == cudaSuccess ? GRB_OK : GRB_ERR_CUDA;
}

int grb_comm_unique_id(char id_out[128]) {
  if (!id_out) return GRB_ERR_ARG;
  NcclApi &n = nccl();
  if (!n.ok) { g_last_global_error = "libnccl.so.2 not found"; return GRB_ERR_NCCL; }
  return n.GetUniqueId(id_out) == 0 ? GRB_OK : GRB_ERR_NCCL;
}

int grb_comm_init(grb_engine *e, int rank, int nranks, const char id[128]) {
  if (!e || !id || rank < 0 || nranks < 1 || rank >= nranks || nranks > 64) return GRB_ERR_ARG;
  NcclApi &n = nccl();
  if (!n.ok) { e->err = "libnccl.so.2 not found"; return GRB_ERR_NCCL; }
  CK(cudaSetDevice(e->device));
  NcclId uid;
  memcpy(uid.b, id, 128);
  const int rc = n.CommInitRank(&e->nccl_comm, nranks, uid, rank);
  if (rc != 0) { e->err = std::string("ncclCommInitRank: ") + (n.GetErrorString ? n.GetErrorString(rc) : "error"); return GRB_ERR_NCCL; }
  e->comm_rank = rank;
  e->comm_size = nranks;
  if (!e->d_max2) CK(cudaMalloc(&e->d_max2, 2 * sizeof(double)));
  return GRB_OK;
}

// ONE collective per iteration: the accumulator block is followed by 2 x nranks slots; every rank parks its own two maxima
// in its slots (the others stay zero), a single sum all-reduce then carries the sums AND an all-gather of the maxima, and
// a one-warp kernel folds the gathered maxima back into (maxW, maxf).  Halves the latency-bound exchange of the iteration
// turn-around (two dependent NCCL launches before).
int grb_allreduce(grb_engine *e) {
  if (!e) return GRB_ERR_ARG;
  if (!e->nccl_comm || e->comm_size <= 1) return GRB_OK;
  NcclApi &n = nccl();
  CK(cudaSetDevice(e->device));
  double *acc = reinterpret_cast<double *>(e->d_acc);
  const size_t nd = sizeof(grb_accum) / sizeof(double);
  k_park_maxima<<<1, 32, 0, e->stream>>>(acc, (int)nd, kAccSum, e->comm_rank, e->comm_size);
  const int rc = n.AllReduce(acc, acc, nd + 2 * (size_t)e->comm_size, kNcclFloat64, kNcclSum, e->nccl_comm, e->stream);
  if (rc != 0) { e->err = std::string("ncclAllReduce: ") + (n.GetErrorString ? n.GetErrorString(rc) : "error"); return GRB_ERR_NCCL; }
  k_fold_maxima<<<1, 32, 0, e->stream>>>(acc, (int)nd, kAccSum, e->comm_size);
  CK(cudaGetLastError());
  return GRB_OK;
}

// reference VEGASData::GetLocalCalls (include/Graniitti/MVEGAS.h:59-69): floor(calls/R) each, remainder to rank 0
int grb_local_calls(int64_t calls, int rank, int nranks, int64_t *first, int64_t *n) {
  if (calls < 0 || nranks < 1 || rank < 0 || rank >= nranks || !first || !n) return GRB_ERR_ARG;
  const int64_t per = calls / nranks, extra = calls - per * nranks;
  *n     = per + (rank == 0 ? extra : 0);
  *first = (rank == 0) ? 0 : extra + per * rank;
  return GRB_OK;
}

// ---- the integration loop (reference MGraniitti::SampleVegas + VEGAS, src/MGraniitti.cc:1074-1339) -----------
namespace {
struct VegasState {
  double sumdata = 0, sumchi2 = 0, sweight = 0;
  double sigma = 0, sigma_err = 0, chi2 = 0;
  double evals = 0, kin = 0, fid = 0, veto = 0, amp = 0, maxW = 0, maxf = 0, device_ms = 0;
  int    iterations = 0;
};

// one VEGAS() call of the reference: init 0 = burn-in (fresh grid + fresh integral), 1 = keep grid, fresh integral.
// returns 1 (done) or 10 (chi2 > 50: caller multiplies the calls by 10), negative on error
int vegas_phase(grb_engine *e, const grb_vegas_param &par, int init, int64_t calls, int itermin, int rank, int nranks,
                grb_reduce_fn reduce, void *ctx, std::vector<double> &xmat, uint64_t &iter_counter, VegasState &st) {
  const int    dim = e->desc.lipsdim, bins = par.bins;
  const double EPS = 1.0e-30;
  if (init == 0)
    for (int i = 0; i < bins; ++i)
      for (int j = 0; j < dim; ++j) xmat[i * dim + j] = (i + 1.0) / bins;
  st = VegasState();
  int64_t my_first = 0, my_n = 0;
  grb_local_calls(calls, rank, nranks, &my_first, &my_n);
  const int max_iter = par.max_iter > 0 ? par.max_iter : 1000;
  for (int iter = 0; iter < itermin; ++iter) {
    void *dacc = nullptr;
    int rc = grb_vegas_iterate_async(e, xmat.data(), bins, dim, my_first, my_n, par.seed, iter_counter++,
                                     init == 0 ? GRB_MODE_BURNIN : GRB_MODE_INTEGRATE, 0.0, &dacc);
    if (rc != GRB_OK) return rc;
    if (nranks > 1) {
      rc = reduce ? reduce(ctx, dacc) : grb_allreduce(e);
      if (rc != GRB_OK) { if (reduce) e->err = "reduce callback failed"; return rc < 0 ? rc : GRB_ERR_NCCL; }
    }
    grb_accum acc;
    rc = grb_vegas_fetch(e, &acc);
    if (rc != GRB_OK) return rc;
    double ms = 0;
    grb_last_timing(e, &ms, nullptr);
    st.device_ms += ms;
    // estimates of this iteration (src/MGraniitti.cc:1222-1251)
    double fsum = acc.fsum * (1.0 / (double)calls);
    double f2sum = acc.f2sum * ((1.0 / (double)calls) * (1.0 / (double)calls));
    const double I_this = fsum;
    f2sum = std::sqrt(std::max(0.0, f2sum * calls));
    f2sum = (f2sum - fsum) * (f2sum + fsum);
    if (f2sum <= 0.0) f2sum = EPS;
    const double I2_this = f2sum * 1.0 / (double)calls;
    const double alpha = 1.0 / I2_this;
    st.sumdata += alpha * I_this;
    st.sumchi2 += alpha * I_this * I_this;
    st.sweight += alpha;
    st.sigma     = st.sumdata / st.sweight;
    st.sigma_err = std::sqrt(std::max(0.0, 1.0 / st.sweight));
    double chi2this = (st.sumchi2 - st.sumdata * st.sumdata / st.sweight) / (iter + 1E-5);
    st.chi2 = chi2this < 0 ? 0.0 : chi2this;
    st.iterations = iter + 1;
    st.evals += acc.evaluations; st.kin += acc.n_kin_ok; st.fid += acc.n_fid_ok; st.veto += acc.n_veto_ok; st.amp += acc.n_amp_ok;
    if (init != 0) { st.maxW = std::max(st.maxW, acc.maxW); st.maxf = std::max(st.maxf, acc.maxf); }
    if (par.debug >= 0 && rank == 0)
      printf("VEGAS:: local iter = %4d integral = %0.5E +- std = %0.5E \t [global integral = %0.5E +- std = %0.5E] \t chi2this = %0.2f \n",
             iter + 1, I_this, std::sqrt(std::max(0.0, I2_this)), st.sigma, st.sigma_err, chi2this);
    if (iter > 0) {
      if (std::isnan(st.sigma) || std::isinf(st.sigma)) { e->err = "VEGAS:: Integral inf/nan: FATAL ERROR, cuts or parameters need fixing"; return GRB_ERR_STATE; }
      if (st.sigma < 1e-60) { e->err = "VEGAS:: Integral < 1E-60: Check VEGAS parameters, decaymode sanity, generation and fiducial cuts!"; return GRB_ERR_STATE; }
      if (st.chi2 > 50) return 10;
    }
    if (init > 0 && iter == itermin - 1 && itermin < max_iter &&
        (st.chi2 > par.chi2max || st.sigma_err / st.sigma > par.precision))
      ++itermin;
    grb_vegas_optimize_grid(xmat.data(), acc.f2mat, bins, dim, par.lambda);
  }
  return 1;
}
}  // namespace

int grb_integrate(grb_engine *e, const grb_vegas_param *par, int rank, int nranks, grb_reduce_fn reduce, void *ctx,
                  grb_vegas_result *out) {
  if (!e || !par || !out || nranks < 1 || rank < 0 || rank >= nranks) return GRB_ERR_ARG;
  if (par->bins < 2 || par->bins > GRB_MAX_BINS || (par->bins % 2) != 0) { e->err = "VEGAS::BINS should be an even number <= 128"; return GRB_ERR_ARG; }
  if (par->ncall < 50 || par->iter < 1 || par->lambda < 1.0 || par->lambda > 10.0) { e->err = "VEGAS parameters out of range"; return GRB_ERR_ARG; }
  if (nranks > 1 && !reduce && !e->nccl_comm) { e->err = "multi-rank integration needs a reduce callback or grb_comm_init"; return GRB_ERR_STATE; }
  const auto t0 = std::chrono::steady_clock::now();
  memset(out, 0, sizeof(*out));
  std::vector<double> xmat((size_t)par->bins * e->desc.lipsdim);
  uint64_t   iter_counter = 0;
  int64_t    ncall = par->ncall;
  int        burn = par->burnin_iter > 0 ? par->burnin_iter : 3;
  VegasState st;
  double     device_ms = 0;
  for (int attempt = 0; attempt < 8; ++attempt) {
    int factor;
    do {  // burn-in until stable (src/MGraniitti.cc:1096-1103)
      factor = vegas_phase(e, *par, 0, ncall, burn, rank, nranks, reduce, ctx, xmat, iter_counter, st);
      device_ms += st.device_ms;
      if (factor < 0) return factor;
      if (factor != 1) { ncall *= factor; ++out->restarts; }
    } while (factor != 1);
    factor = vegas_phase(e, *par, 1, ncall, par->iter, rank, nranks, reduce, ctx, xmat, iter_counter, st);
    device_ms += st.device_ms;
    if (factor < 0) return factor;
    if (factor == 1) break;
    ncall *= factor;
    burn *= 2;
    ++out->restarts;
  }
  out->sigma = st.sigma; out->sigma_err = st.sigma_err; out->chi2 = st.chi2;
  out->maxW = st.maxW; out->maxf = st.maxf;
  out->evaluations = st.evals; out->n_kin_ok = st.kin; out->n_fid_ok = st.fid; out->n_veto_ok = st.veto; out->n_amp_ok = st.amp;
  out->iterations = st.iterations; out->ncall_final = ncall; out->device_ms = device_ms;
  for (int i = 0; i < par->bins; ++i)
    for (int j = 0; j < e->desc.lipsdim; ++j) out->xmat[i * e->desc.lipsdim + j] = xmat[i * e->desc.lipsdim + j];
  out->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return GRB_OK;
}

int grb_sample_flat(grb_engine *e, const grb_flat_param *par, int rank, int nranks, grb_reduce_fn reduce, void *ctx,
                    grb_vegas_result *out) {
  if (!e || !par || !out || nranks < 1 || rank < 0 || rank >= nranks) return GRB_ERR_ARG;
  if (nranks > 1 && !reduce && !e->nccl_comm) { e->err = "multi-rank sampling needs a reduce callback or grb_comm_init"; return GRB_ERR_STATE; }
  const auto t0 = std::chrono::steady_clock::now();
  memset(out, 0, sizeof(*out));
  const int     dim = e->desc.lipsdim, bins = 2;   // a two-bin uniform grid: x = u, vegasweight = 1 exactly
  const int64_t batch = par->batch > 0 ? par->batch : (int64_t)1 << 22;
  const double  max_events = par->max_events > 0 ? (double)par->max_events : 1099511627776.0;
  const double  precision = par->precision > 0 ? par->precision : 0.05, min_events = par->min_events > 0 ? par->min_events : 1e6;
  std::vector<double> xmat((size_t)bins * dim);
  for (int i = 0; i < bins; ++i)
    for (int j = 0; j < dim; ++j) xmat[i * dim + j] = (i + 1.0) / bins;
  double Wsum = 0, W2sum = 0, evals = 0;
  int64_t my_first = 0, my_n = 0;
  grb_local_calls(batch, rank, nranks, &my_first, &my_n);
  for (uint64_t it = 0;; ++it) {
    void *dacc = nullptr;
    int rc = grb_vegas_iterate_async(e, xmat.data(), bins, dim, my_first, my_n, par->seed, it, GRB_MODE_INTEGRATE, 0.0, &dacc);
    if (rc != GRB_OK) return rc;
    if (nranks > 1) {
      rc = reduce ? reduce(ctx, dacc) : grb_allreduce(e);
      if (rc != GRB_OK) return rc < 0 ? rc : GRB_ERR_NCCL;
    }
    grb_accum acc;
    rc = grb_vegas_fetch(e, &acc);
    if (rc != GRB_OK) return rc;
    double ms = 0;
    grb_last_timing(e, &ms, nullptr);
    out->device_ms += ms;
    Wsum += acc.fsum; W2sum += acc.f2sum; evals += acc.evaluations;
    out->n_kin_ok += acc.n_kin_ok; out->n_fid_ok += acc.n_fid_ok; out->n_veto_ok += acc.n_veto_ok; out->n_amp_ok += acc.n_amp_ok;
    out->maxW = std::max(out->maxW, acc.maxW); out->maxf = std::max(out->maxf, acc.maxf);
    out->iterations = (int32_t)(it + 1);
    // Stats::CalculateCrossSection (include/Graniitti/MGraniitti.h:59-66)
    out->sigma     = Wsum / evals;
    out->sigma_err = std::sqrt(std::max(0.0, (W2sum / evals - out->sigma * out->sigma) / evals));
    if (std::isnan(out->sigma) || std::isinf(out->sigma)) { e->err = "FLAT:: Integral inf/nan"; return GRB_ERR_STATE; }
    if ((out->sigma > 0 && out->sigma_err / out->sigma < precision && evals > min_events) || evals >= max_events) break;
  }
  out->evaluations = evals; out->ncall_final = batch;
  for (int i = 0; i < bins * dim; ++i) out->xmat[i] = xmat[i];
  out->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return GRB_OK;
}

int grb_measure_dfma_peak(int device, double *tflops_out) {
  if (!tflops_out) return GRB_ERR_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= device) return GRB_ERR_CUDA;
  if (cudaSetDevice(device) != cudaSuccess) return GRB_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return GRB_ERR_CUDA;
  return run_dfma_peak(prop.multiProcessorCount, tflops_out) == cudaSuccess ? GRB_OK : GRB_ERR_CUDA;
}

int grb_last_timing(const grb_engine *e, double *kernel_ms, int *n_launches) {
  if (!e) return GRB_ERR_ARG;
  if (kernel_ms) {
    *kernel_ms = 0.0;
    if (e->timed) {
      cudaSetDevice(e->device);
      if (cudaEventSynchronize(e->ev1) == cudaSuccess) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, e->ev0, e->ev1) == cudaSuccess) *kernel_ms = ms;
      }
    }
  }
  if (n_launches) *n_launches = e->last_launches;
  return GRB_OK;
}

void *grb_get_stream(grb_engine *e) { return e ? (void *)e->stream : nullptr; }

int grb_last_timing_kernels(const grb_engine *e, double *ms_total, double *ms_screen, int *n_screen_launches) {
  if (!e) return GRB_ERR_ARG;
  double tot = 0.0, scr = 0.0;
  if (e->timed) {
    cudaSetDevice(e->device);
    if (cudaEventSynchronize(e->ev1) == cudaSuccess) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, e->ev0, e->ev1) == cudaSuccess) tot = ms;
      for (int i = 0; i + 1 < e->n_evk; i += 2)
        if (cudaEventElapsedTime(&ms, e->evk[i], e->evk[i + 1]) == cudaSuccess) scr += ms;
    }
  }
  if (ms_total) *ms_total = tot;
  if (ms_screen) *ms_screen = scr;
  if (n_screen_launches) *n_screen_launches = e->n_evk / 2;
  return GRB_OK;
}

}  // extern "C"
