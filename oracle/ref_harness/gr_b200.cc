// oracle/ref_harness/gr_b200.cc -- INTEGRATION.md route A, compiled: the UNMODIFIED reference library does the steering
// (MGraniitti::ReadInput -- JSON5 card, process string, @-commands, MODELPARAM tune, PDG table, resonance files, branching
// tables) and this file is the binding a maintainer adds: it flattens the reference's PARSED OBJECTS (proc->lts, gcuts,
// fcuts, vetocuts, PARAM_SOFT / PARAM_REGGE / PARAM_STRUCTURE / PARAM_FLAT, MTensorPomeronParam, lts.RESONANCES) into the
// POD grb_process_desc of include/grb.h, hands the reference's own screening table (proc->Eikonal.MSA) to grb_set_eikonal
// and runs grb_integrate in place of MGraniitti::CallIntegrator (reference src/MGraniitti.cc:1016-1071, Program/gr.cc:
// 167-219).  Nothing of the steering is re-parsed here.
//
// Built by oracle/build_ref.sh against the reference objects and graniitti_b200/libgrb.so (test infrastructure: it lives
// next to ref_eval and proves that the C ABI binds to the reference as INTEGRATION.md says).
//
//   gr_b200 desc <card> <process|null> <pomloop -1|0|1> <out.bin>            host only: writes the grb_process_desc bytes
//   gr_b200 run  <card> <process|null> <pomloop> <ncall> <iter> <precision> <seed>     needs a GPU: prints sigma +- err
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

#include "Graniitti/MAux.h"
#include "Graniitti/MForm.h"
#include "Graniitti/MGlobals.h"
#include "Graniitti/MGraniitti.h"
#include "Graniitti/MProcess.h"
#include "Graniitti/MRegge.h"
#include "Graniitti/MTensorPomeron.h"
#include "grb.h"

namespace {

// protected members of MProcess read through pointers-to-member formed in a derived class (no reference source is touched)
struct Peek : public gra::MProcess {
  static double S(const gra::MProcess &p) { return p.*(&Peek::S_factor); }
  static int    Excitation(const gra::MProcess &p) { return p.*(&Peek::EXCITATION); }
  static int    UserCuts(const gra::MProcess &p) { return p.*(&Peek::USERCUTS); }
  static int    FlatAmp(const gra::MProcess &p) { return p.*(&Peek::FLATAMP); }
  static std::string Cid(const gra::MProcess &p) { return p.*(&Peek::CID); }
};

int frame_id(const std::string &f) {  // gra::spin::GetRotation frames on the device path
  if (f == "CM") return 0;
  if (f == "HX") return 1;
  if (f == "CS") return 2;
  if (f == "GJ") return 3;
  throw std::invalid_argument("FRAME " + f + " is outside the device path");
}

int proc_id(const std::string &istate, const std::string &channel) {  // MSubProc registry keys -> GRB_PROC_*
  const std::string k = istate + "[" + channel + "]";
  if (k == "yy[CON]") return GRB_PROC_YY_CON;
  if (k == "yy[QED]") return GRB_PROC_YY_QED;
  if (k == "PP[RES]") return GRB_PROC_PP_RES;
  if (k == "PP[CONTENSOR]") return GRB_PROC_PP_CONTENSOR;
  if (k == "PP[CON]") return GRB_PROC_PP_CON;
  if (k == "PP[CON-]") return GRB_PROC_PP_CON_MINUS;
  if (k == "PP[RES+CON]") return GRB_PROC_PP_RES_CON;
  if (k == "PP[RESTENSOR]") return GRB_PROC_PP_RESTENSOR;
  if (k == "PP[RES+CONTENSOR]") return GRB_PROC_PP_RES_CONTENSOR;
  throw std::invalid_argument("process " + k + " has no device amplitude");
}

void setup(gra::MGraniitti &g, const std::string &card, const std::string &proc, int pomloop) {
  g.HILJAA = true;
  g.ReadInput(card, proc);
  if (pomloop >= 0) g.proc->SetScreening(pomloop != 0);
  g.proc->SetHistograms(0);
  g.proc->post_Constructor();
}

// the reference's decay tree (lts.decaytree, MDecayBranch) -> grb_node[]: first-level legs first, then breadth-first
void flatten_tree(const std::vector<gra::MDecayBranch> &legs, grb_process_desc &d) {
  std::vector<std::pair<const gra::MDecayBranch *, int>> queue;
  int  n   = 0;
  auto put = [&](const gra::MDecayBranch &b, int parent) {
    if (n >= GRB_MAX_NODES) throw std::invalid_argument("decay tree exceeds GRB_MAX_NODES");
    grb_node &o = d.nodes[n];
    o.pdg = b.p.pdg; o.spinX2 = b.p.spinX2; o.parent = parent; o.first_child = -1;
    o.n_children = (int)b.legs.size(); o.mass = b.p.mass; o.width = b.p.width;
    queue.push_back({&b, n});
    return n++;
  };
  for (const auto &b : legs) put(b, -1);
  for (size_t q = 0; q < queue.size(); ++q)
    if (!queue[q].first->legs.empty()) {
      d.nodes[queue[q].second].first_child = n;
      for (const auto &c : queue[q].first->legs) put(c, queue[q].second);
    }
  d.n_nodes = n;
  d.n_legs  = (int)legs.size();
  d.n_final = 0;
  for (int i = 0; i < n; ++i) d.n_final += d.nodes[i].n_children == 0;
}

void fill_desc(gra::MGraniitti &g, grb_process_desc &d) {
  memset(&d, 0, sizeof(d));
  gra::MProcess &p          = *g.proc;
  const std::string modelfile = gra::aux::GetBasePath(2) + "/modeldata/" + gra::MODELPARAM + "/GENERAL.json";
  d.abi_version = GRB_ABI_VERSION;
  const std::string cid = Peek::Cid(p);
  if (cid == "F") d.ps_class = GRB_CLASS_F;
  else if (cid == "C") d.ps_class = GRB_CLASS_C;
  else throw std::invalid_argument("phase-space class <" + cid + "> is outside the hot-path scope");
  d.proc       = proc_id(p.ProcPtr.ISTATE, p.ProcPtr.CHANNEL);
  d.screening  = p.GetScreening();
  d.excitation = Peek::Excitation(p);
  d.flatamp    = Peek::FlatAmp(p);
  d.ps_active  = !p.GetISOLATE();
  d.flatmass2  = p.GetFLATMASS2();
  d.lipsdim    = (int)p.GetdLIPSDim();
  flatten_tree(p.lts.decaytree, d);
  d.S_factor = Peek::S(p);
  d.offshell = p.GetOFFSHELL();
  // internal uniforms per point when no rejection loop repeats (one Breit-Wigner draw per unstable node, two decay angles
  // per two-body decay; MProcess.cc:753-812, MKinematics.h:405-409)
  {
    auto bw = [&](int i) { return d.nodes[i].width < 1e-40 ? 0 : 1; };
    auto draws = [&](int i) {
      if (!bw(i)) return 0;
      int k = 1;
      for (int c = 0; c < d.nodes[i].n_children; ++c) k += bw(d.nodes[i].first_child + c);
      return k;
    };
    int nt = 2;
    for (int i = 0; i < d.n_legs; ++i) nt += draws(i);
    for (int i = 0; i < d.n_nodes; ++i)
      if (d.nodes[i].n_children) {
        for (int c = 0; c < d.nodes[i].n_children; ++c) nt += draws(d.nodes[i].first_child + c);
        nt += 2;
      }
    d.n_tape = nt;
  }
  // beams and s (MProcess::SetInitialState / SetBeamEnergies)
  d.s = p.lts.s; d.sqrt_s = p.lts.sqrt_s;
  d.pbeam1_pz = p.lts.pbeam1.Pz(); d.pbeam1_E = p.lts.pbeam1.E();
  d.pbeam2_pz = p.lts.pbeam2.Pz(); d.pbeam2_E = p.lts.pbeam2.E();
  d.beam1_mass = p.lts.beam1.mass; d.beam2_mass = p.lts.beam2.mass;
  // screening loop discretisation (MEikonalNumerics after SetLoopDiscretization(ND))
  {
    const auto &N = p.Eikonal.Numerics;
    d.loop_nkt = (int)N.NumberLoopKT; d.loop_nphi = (int)N.NumberLoopPHI;
    d.loop_min_kt = N.MinLoopKT; d.loop_max_kt = N.MaxLoopKT;
  }
  // cuts as the reference holds them after post_Constructor (SetTechnicalBoundaries, MContinuum defaults)
  {
    const gra::GENCUT &c = p.gcuts;
    grb_gencut &o = d.gcuts;
    o.forward_pt_min = c.forward_pt_min; o.forward_pt_max = c.forward_pt_max; o.XI_min = c.XI_min; o.XI_max = c.XI_max;
    if (d.ps_class == GRB_CLASS_C) { o.kt_min = c.kt_min; o.kt_max = c.kt_max; o.rap_min = c.rap_min; o.rap_max = c.rap_max; }
    else { o.Y_min = c.Y_min; o.Y_max = c.Y_max; o.M_min = c.M_min; o.M_max = c.M_max; o.kt_min = 0.0; o.kt_max = -1.0; }
  }
  {
    const gra::FIDCUT &c = p.fcuts;
    grb_fidcut &o = d.fcuts;
    o.active = c.active; o.usercuts = c.active ? Peek::UserCuts(p) : 0;
    if (c.active) {
      o.eta_min = c.eta_min; o.eta_max = c.eta_max; o.rap_min = c.rap_min; o.rap_max = c.rap_max;
      o.pt_min = c.pt_min; o.pt_max = c.pt_max; o.Et_min = c.Et_min; o.Et_max = c.Et_max;
      o.M_min = c.M_min; o.M_max = c.M_max; o.Y_min = c.Y_min; o.Y_max = c.Y_max; o.Pt_min = c.Pt_min; o.Pt_max = c.Pt_max;
      o.forward_t_min = c.forward_t_min; o.forward_t_max = c.forward_t_max;
      o.forward_M_min = c.forward_M_min; o.forward_M_max = c.forward_M_max;
    }
  }
  d.veto_active = p.vetocuts.active;
  for (const auto &v : p.vetocuts.cuts) {
    if (d.n_veto >= GRB_MAX_VETO) throw std::invalid_argument("more veto domains than GRB_MAX_VETO");
    d.veto[d.n_veto++] = {v.eta_min, v.eta_max, v.pt_min, v.pt_max};
  }
  // model parameters: the reference's own readers fill its PARAM_* namespaces (MGraniitti::ReadModelParam did the generic
  // blocks; the amplitude blocks are read lazily by the amplitude classes -- MRegge::MRegge, MTensorPomeron -- here up front)
  gra::PARAM_STRUCTURE::ReadParameters(modelfile);
  gra::PARAM_FLAT::ReadParameters(modelfile);
  {
    grb_param_soft &o = d.soft;
    o.DELTA_P = gra::PARAM_SOFT::DELTA_P; o.ALPHA_P = gra::PARAM_SOFT::ALPHA_P; o.gN_P = gra::PARAM_SOFT::gN_P;
    o.gN_O = gra::PARAM_SOFT::gN_O; o.g3P = gra::PARAM_SOFT::g3P; o.gamma = gra::PARAM_SOFT::gamma;
    o.fc1 = gra::PARAM_SOFT::fc1; o.fc2 = gra::PARAM_SOFT::fc2; o.fc3 = gra::PARAM_SOFT::fc3;
    o.odderon_on = gra::PARAM_SOFT::ODDERON_ON;
    o.f2_model   = gra::PARAM_STRUCTURE::F2 == "DL" ? 1 : 0;
    d.flat_b     = gra::PARAM_FLAT::b;
    d.em_ff      = gra::PARAM_STRUCTURE::EM == "KELLY" ? 1 : 0;
    d.qed_alpha  = gra::PARAM_STRUCTURE::QED_alpha == "ZERO" ? 0 : 1;
  }
  if (!p.lts.decaytree.empty()) {
    gra::PARAM_REGGE::ReadParameters(std::abs(p.lts.decaytree[0].p.pdg), modelfile);  // as MRegge::MRegge does (MRegge.cc:37-56)
    grb_param_regge &o = d.regge;
    for (int k = 0; k < 3; ++k) {
      o.a0[k] = gra::PARAM_REGGE::a0.at(k); o.ap[k] = gra::PARAM_REGGE::ap.at(k); o.sgn[k] = gra::PARAM_REGGE::sgn.at(k);
      o.c[k] = gra::PARAM_REGGE::c.at(k); o.n[k] = gra::PARAM_REGGE::n.at(k);
    }
    o.s0 = gra::PARAM_REGGE::s0; o.offshellFF = gra::PARAM_REGGE::offshellFF; o.reggeize = gra::PARAM_REGGE::reggeize;
    o.b_EXP = gra::PARAM_REGGE::b_EXP; o.a_OREAR = gra::PARAM_REGGE::a_OREAR; o.b_OREAR = gra::PARAM_REGGE::b_OREAR;
    o.b_POW = gra::PARAM_REGGE::b_POW; o.omega = gra::PARAM_REGGE::omega;
  }
  {
    gra::MTensorPomeronParam T;
    T.ReadParameters(modelfile);
    grb_param_tensorpom &o = d.tpom;
    o.delta_P = T.delta_P; o.delta_O = T.delta_O; o.delta_1R = T.delta_1R; o.delta_2R = T.delta_2R;
    o.ap_P = T.ap_P; o.ap_O = T.ap_O; o.ap_1R = T.ap_1R; o.ap_2R = T.ap_2R; o.eta_O = T.eta_O;
    o.gPNN = T.gPNN; o.gPpipi = T.gPpipi; o.gPKK = T.gPKK;
    o.gPrhorho[0] = T.gPrhorho.at(0); o.gPrhorho[1] = T.gPrhorho.at(1);
    o.gPphiphi[0] = T.gPphiphi.at(0); o.gPphiphi[1] = T.gPphiphi.at(1);
    o.M_O = T.M_O; o.M_1R = T.M_1R;
  }
  // resonances: lts.RESONANCES after SetResonances + SetupBranching (std::map: name order)
  for (const auto &kv : p.lts.RESONANCES) {
    if (d.n_res >= GRB_MAX_RES) throw std::invalid_argument("more resonances than GRB_MAX_RES");
    const gra::PARAM_RES &r = kv.second;
    grb_resonance &o = d.res[d.n_res++];
    o.pdg = r.p.pdg; o.spinX2 = r.p.spinX2; o.P = r.p.P; o.BW = r.BW;
    o.spingen = r.SPINGEN; o.spindec = r.SPINDEC; o.frame = frame_id(r.FRAME); o.jmax = r.JMAX;
    o.mass = r.p.mass; o.width = r.p.width; o.g_re = r.g.real(); o.g_im = r.g.imag(); o.g_FF = r.g_FF;
    o.g_decay = r.hel.g_decay; o.BR = r.hel.BR;
    for (size_t k = 0; k < r.g_Tensor.size() && k < 8; ++k) o.g_Tensor[k] = r.g_Tensor[k];
    for (size_t k = 0; k < r.hel.g_decay_tensor.size() && k < 2; ++k) o.g_decay_tensor[k] = r.hel.g_decay_tensor[k];
    if (r.p.spinX2 != 0) {
      if (r.rho.size_row() > GRB_RHO_DIM) throw std::invalid_argument("resonance spin exceeds GRB_RHO_DIM");
      for (size_t a = 0; a < r.rho.size_row(); ++a)
        for (size_t b = 0; b < r.rho.size_col(); ++b) {
          o.rho_re[a * GRB_RHO_DIM + b] = r.rho[a][b].real();
          o.rho_im[a * GRB_RHO_DIM + b] = r.rho[a][b].imag();
        }
    }
    o.T_rows = (int)r.hel.T.size_row(); o.T_cols = (int)r.hel.T.size_col();
    if (o.T_rows * o.T_cols > 9) throw std::invalid_argument("T matrix larger than 3x3");
    for (int a = 0; a < o.T_rows; ++a)
      for (int b = 0; b < o.T_cols; ++b) {
        o.T_re[a * o.T_cols + b] = r.hel.T[a][b].real();
        o.T_im[a * o.T_cols + b] = r.hel.T[a][b].imag();
      }
  }
}

}  // namespace

int main(int argc, char **argv) {
  if (argc < 6) {
    fprintf(stderr, "usage: gr_b200 desc|run <card> <process|null> <pomloop> ...\n");
    return 2;
  }
  const std::string mode = argv[1];
  try {
    gra::MGraniitti g;
    setup(g, argv[2], argv[3], atoi(argv[4]));
    grb_process_desc d;
    fill_desc(g, d);
    if (mode == "desc") {
      FILE *fo = fopen(argv[5], "wb");
      if (!fo) return 2;
      fwrite(&d, sizeof(d), 1, fo);
      fclose(fo);
      return 0;
    }
    if (mode != "run" || argc < 9) return 2;
    grb_engine *e = nullptr;
    int rc = grb_create(&d, 0, &e);
    if (rc != GRB_OK) {
      fprintf(stderr, "grb_create: %d %s\n", rc, grb_last_error(nullptr));
      return 1;
    }
    if (d.screening) {  // the reference's own MSA table (read from its eikonal/ cache or computed by S3Constructor)
      g.proc->Eikonal.S3Constructor(g.proc->GetMandelstam_s(), g.proc->GetInitialState(), false);
      const auto &F = g.proc->Eikonal.MSA.F;
      const unsigned n = g.proc->Eikonal.Numerics.NumberKT2;
      std::vector<double> x(n + 1), re(n + 1), im(n + 1);
      for (unsigned i = 0; i <= n; ++i) { x[i] = F[i][0].real(); re[i] = F[i][1].real(); im[i] = F[i][1].imag(); }
      rc = grb_set_eikonal(e, x.data(), re.data(), im.data(), (int)n + 1);
      if (rc != GRB_OK) { fprintf(stderr, "grb_set_eikonal: %s\n", grb_last_error(e)); return 1; }
    }
    grb_vegas_param par;
    memset(&par, 0, sizeof(par));
    par.bins = 128; par.iter = atoi(argv[6]); par.burnin_iter = 3; par.ncall = atoll(argv[5]); par.lambda = 1.5;
    par.chi2max = 10.0; par.precision = atof(argv[7]); par.seed = strtoull(argv[8], nullptr, 10); par.debug = -1;
    grb_vegas_result r;
    rc = grb_integrate(e, &par, 0, 1, nullptr, nullptr, &r);
    if (rc != GRB_OK) { fprintf(stderr, "grb_integrate: %s\n", grb_last_error(e)); return 1; }
    printf("RESULT sigma %.17g err %.17g chi2 %.17g iterations %d evaluations %.0f\n", r.sigma, r.sigma_err, r.chi2, r.iterations,
           r.evaluations);
    grb_destroy(e);
  } catch (const std::exception &ex) {
    fprintf(stderr, "gr_b200: exception: %s\n", ex.what());
    return 1;
  }
  return 0;
}
