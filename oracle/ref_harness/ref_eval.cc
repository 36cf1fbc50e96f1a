// oracle/ref_harness/ref_eval.cc -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Per-point driver around the UNMODIFIED GRANIITTI reference library (objects compiled from
// /root/reference by oracle/build_ref.sh). It calls the reference's seam A,
//     double MProcess::EventWeight(const std::vector<double>& randvec, AuxIntData& aux)
// (reference include/Graniitti/MProcess.h:84) on caller supplied points, with the *internal* uniforms
// (gra::MRandom::U, reference src/MRandom.cc:62-65 -- drawn inside EventWeight for the central decay
// angles, Breit-Wigner masses, ...) read from a caller supplied tape, so that a device implementation
// can be compared point by point on identical inputs.
//
// Call sequence follows reference src/MGraniitti.cc:1016-1038 (Initialize) and :1446-1482 (SampleFlat):
//   MGraniitti g; g.ReadInput(card, proc); proc->post_Constructor(); [Eikonal.S3Constructor]; EventWeight
//
// The strong definition of gra::MRandom::U below overrides the (weakened, objcopy --weaken-symbol)
// original in MRandom.o, including calls made from inside MRandom.o (RelativisticBWRandom).
//
// usage:
//   ref_eval eval  <card.json> <process|null> <pomloop 0|1|-1(card)> <in.bin> <out.bin> [kappa 0|1|2]
//   ref_eval info  <card.json> <process|null> <pomloop>              (prints flattened process facts)
//   ref_eval msa   <card.json> <process|null> <out.bin>              (dumps the MSA eikonal table)
//   ref_eval record <card.json> <process|null> <pomloop> <in.bin> <out.txt>   (the reference's event record of every valid point)
//
// in.bin : int64 n, int64 du, int64 dt, then n*(du+dt) doubles (point-major: u[du] then tape[dt])
// out.bin: int64 n, int64 rec, then n*rec doubles, record layout = see `struct Rec` below.

#include <chrono>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iostream>
#include <string>
#include <vector>

#include "Graniitti/MGraniitti.h"
#include "Graniitti/MProcess.h"
#include "Graniitti/MRandom.h"

namespace {
const double *g_tape     = nullptr;
int64_t       g_tape_len = 0;
int64_t       g_tape_pos = 0;
int64_t       g_overrun  = 0;
}  // namespace

// Tape driven replacement of reference src/MRandom.cc:62-65 (same a + (b-a)*value affine map).
double gra::MRandom::U(double a, double b) {
  double value = 0.5;
  if (g_tape != nullptr && g_tape_pos < g_tape_len) {
    value = g_tape[g_tape_pos];
  } else {
    ++g_overrun;
  }
  ++g_tape_pos;
  return a + (b - a) * value;
}

namespace {

constexpr int NHAMP = 16;  // helicity amplitudes stored per point (first NHAMP)
constexpr int NPART = 9;   // pfinal[0..8] stored per point

struct Rec {
  double W;
  double kin_ok, fid_ok, veto_ok, amp_ok;
  double tape_used;
  double t1, t2, s1, s2, m2, Y, Pt, t_hat, u_hat;
  double nhamp;
  double kappa;  // max_k |W(u +- ulp e_k) - W| / |W| (only if requested, else -1)
  double DW;     // central phase-space weight lts.DW.Integral()
  double hamp[2 * NHAMP];
  double p[4 * NPART];  // px,py,pz,E of pfinal[i]
};
constexpr int RECLEN = sizeof(Rec) / sizeof(double);

double eval_point(gra::MGraniitti &g, const double *row, int64_t du, int64_t dt, gra::AuxIntData &aux,
                  const std::vector<double> *uoverride = nullptr, const std::vector<double> *tapeoverride = nullptr) {
  std::vector<double> u(row, row + du);
  if (uoverride) u = *uoverride;
  g_tape     = tapeoverride ? tapeoverride->data() : row + du;
  g_tape_len = dt;
  g_tape_pos = 0;
  aux        = gra::AuxIntData();
  return g.proc->EventWeight(u, aux);
}

void setup(gra::MGraniitti &g, const std::string &card, const std::string &proc, int pomloop) {
  g.HILJAA = true;
  g.ReadInput(card, proc);
  if (pomloop >= 0) g.proc->SetScreening(pomloop != 0);
  g.proc->SetHistograms(0);
  g.proc->post_Constructor();
  if (g.proc->GetScreening()) {
    g.proc->Eikonal.S3Constructor(g.proc->GetMandelstam_s(), g.proc->GetInitialState(), false);
  }
}

}  // namespace

int main(int argc, char **argv) {
  if (argc < 5) {
    fprintf(stderr, "usage: ref_eval eval|info|msa <card> <process|null> ...\n");
    return 2;
  }
  const std::string mode = argv[1];
  const std::string card = argv[2];
  const std::string proc = argv[3];
  try {
    gra::MGraniitti g;
    if (mode == "msa") {
      setup(g, card, proc, 1);
      FILE *fo = fopen(argv[4], "wb");
      const auto &N = g.proc->Eikonal.Numerics;
      double hdr[8] = {(double)N.NumberKT2, N.MinKT2, N.MaxKT2, (double)N.NumberLoopKT,
                       (double)N.NumberLoopPHI, N.MinLoopKT, N.MaxLoopKT, 0};
      fwrite(hdr, sizeof(double), 8, fo);
      // the table rows as held in memory (= read back from the 15-digit ASCII cache)
      const auto &F = g.proc->Eikonal.MSA.F;
      for (unsigned int i = 0; i <= N.NumberKT2; ++i) {
        double row[3] = {F[i][0].real(), F[i][1].real(), F[i][1].imag()};
        fwrite(row, sizeof(double), 3, fo);
      }
      fclose(fo);
      return 0;
    }
    const int pomloop = atoi(argv[4]);
    setup(g, card, proc, pomloop);

    if (mode == "info") {
      auto &p = *g.proc;
      printf("INFO s %.17g\n", p.lts.s);
      printf("INFO sqrt_s %.17g\n", p.lts.sqrt_s);
      printf("INFO pbeam1 %.17g %.17g\n", p.lts.pbeam1.Pz(), p.lts.pbeam1.E());
      printf("INFO pbeam2 %.17g %.17g\n", p.lts.pbeam2.Pz(), p.lts.pbeam2.E());
      printf("INFO beam_mass %.17g %.17g\n", p.lts.beam1.mass, p.lts.beam2.mass);
      printf("INFO lipsdim %u\n", p.GetdLIPSDim());
      printf("INFO screening %d\n", (int)p.GetScreening());
      printf("INFO gcuts Y %.17g %.17g M %.17g %.17g fpt %.17g %.17g\n", p.gcuts.Y_min, p.gcuts.Y_max,
             p.gcuts.M_min, p.gcuts.M_max, p.gcuts.forward_pt_min, p.gcuts.forward_pt_max);
      printf("INFO fcuts active %d eta %.17g %.17g rap %.17g %.17g pt %.17g %.17g Et %.17g %.17g\n",
             (int)p.fcuts.active, p.fcuts.eta_min, p.fcuts.eta_max, p.fcuts.rap_min, p.fcuts.rap_max,
             p.fcuts.pt_min, p.fcuts.pt_max, p.fcuts.Et_min, p.fcuts.Et_max);
      printf("INFO fcuts M %.17g %.17g Y %.17g %.17g Pt %.17g %.17g t %.17g %.17g\n", p.fcuts.M_min,
             p.fcuts.M_max, p.fcuts.Y_min, p.fcuts.Y_max, p.fcuts.Pt_min, p.fcuts.Pt_max,
             p.fcuts.forward_t_min, p.fcuts.forward_t_max);
      for (size_t i = 0; i < p.lts.decaytree.size(); ++i) {
        const auto &b = p.lts.decaytree[i];
        printf("INFO leg %zu pdg %d mass %.17g width %.17g nlegs %zu\n", i, b.p.pdg, b.p.mass,
               b.p.width, b.legs.size());
      }
      for (const auto &kv : p.lts.RESONANCES) {
        const auto &r = kv.second;
        printf("INFO res %s pdg %d mass %.17g width %.17g spinX2 %d P %d BW %d g %.17g %.17g gFF %.17g BR %.17g gdecay %.17g "
               "spingen %d spindec %d frame %s jmax %d T %zu %zu",
               kv.first.c_str(), r.p.pdg, r.p.mass, r.p.width, r.p.spinX2, r.p.P, r.BW, r.g.real(), r.g.imag(), r.g_FF,
               r.hel.BR, r.hel.g_decay, (int)r.SPINGEN, (int)r.SPINDEC, r.FRAME.c_str(), r.JMAX, r.hel.T.size_row(),
               r.hel.T.size_col());
        for (size_t a = 0; a < r.hel.T.size_row(); ++a)
          for (size_t b = 0; b < r.hel.T.size_col(); ++b) printf(" %.17g %.17g", r.hel.T[a][b].real(), r.hel.T[a][b].imag());
        printf(" rho %zu", r.rho.size_row());
        for (size_t a = 0; a < r.rho.size_row(); ++a)
          for (size_t b = 0; b < r.rho.size_col(); ++b) printf(" %.17g %.17g", r.rho[a][b].real(), r.rho[a][b].imag());
        printf("\n");
      }
      if (p.GetScreening()) {
        const auto &N = p.Eikonal.Numerics;
        printf("INFO loop nkt %u nphi %u minkt %.17g maxkt %.17g\n", N.NumberLoopKT, N.NumberLoopPHI,
               N.MinLoopKT, N.MaxLoopKT);
      }
      return 0;
    }

    if (mode == "record" && argc >= 7) {
      // The reference's own event record of every valid point: EventWeight, then MProcess::EventRecord (= CommonRecord,
      // src/MProcess.cc:1588-1706) into the stub GenEvent, dumped as text: "EVENT k np nv", "P id pid status parent px py pz e"
      // (ids in the order HepMC3::GenEvent assigns them: incoming then outgoing particles of each vertex as it is attached),
      // "V id n_in in...".
      FILE *fi = fopen(argv[5], "rb");
      if (!fi) return 2;
      int64_t hdr[3];
      if (fread(hdr, sizeof(int64_t), 3, fi) != 3) return 2;
      const int64_t n = hdr[0], du = hdr[1], dt = hdr[2];
      std::vector<double> in(n * (du + dt));
      if ((int64_t)fread(in.data(), sizeof(double), in.size(), fi) != (int64_t)in.size()) return 2;
      fclose(fi);
      FILE *fo = fopen(argv[6], "w");
      for (int64_t k = 0; k < n; ++k) {
        gra::AuxIntData aux;
        const double    W = eval_point(g, &in[k * (du + dt)], du, dt, aux);
        if (!(aux.Valid() && W != 0.0)) continue;
        HepMC3::GenEvent evt(HepMC3::Units::GEV, HepMC3::Units::MM);
        if (!g.proc->EventRecord(evt)) continue;
        std::vector<const HepMC3::GenParticle *> order;
        auto id_of = [&](const HepMC3::GenParticlePtr &p) {
          for (size_t i = 0; i < order.size(); ++i)
            if (order[i] == p.get()) return (int)i + 1;
          order.push_back(p.get());
          return (int)order.size();
        };
        // the stub keeps vertices in add_vertex order; v4 (system -> daughters) is added before its particles are attached,
        // nested decay vertices are added (with their mother attached) before the next first-level daughter: a depth-first walk
        // from v4 reproduces the order in which the real GenEvent sees the particles
        std::vector<int> parent_vertex(64, 0);
        const auto &     vt = evt.vtx;
        std::function<void(size_t)> walk = [&](size_t vi) {
          for (const auto &p : vt[vi]->in) id_of(p);
          for (const auto &p : vt[vi]->out) {
            const int id = id_of(p);
            parent_vertex[id] = -(int)(vi + 1);
            for (size_t w = vi + 1; w < vt.size(); ++w)
              if (!vt[w]->in.empty() && vt[w]->in[0].get() == p.get() && vt[w]->in.size() == 1 && w >= 4) walk(w);
          }
        };
        for (size_t vi = 0; vi < vt.size() && vi < 4; ++vi) walk(vi);
        fprintf(fo, "EVENT %lld %zu %zu %.17g\n", (long long)k, order.size(), vt.size(), W);
        for (size_t i = 0; i < order.size(); ++i) {
          const auto &m = order[i]->momentum();
          fprintf(fo, "P %zu %d %d %d %.17g %.17g %.17g %.17g\n", i + 1, order[i]->pid(), order[i]->status(), parent_vertex[i + 1], m.px(),
                  m.py(), m.pz(), m.e());
        }
        for (size_t vi = 0; vi < vt.size(); ++vi) {
          fprintf(fo, "V %d %zu", -(int)(vi + 1), vt[vi]->in.size());
          for (const auto &p : vt[vi]->in) fprintf(fo, " %d", id_of(p));
          fprintf(fo, " %zu\n", vt[vi]->out.size());
        }
      }
      fclose(fo);
      return 0;
    }

    if (mode != "eval" || argc < 7) {
      fprintf(stderr, "bad arguments\n");
      return 2;
    }
    const int  kappa_mode = (argc >= 8) ? atoi(argv[7]) : 0;  // 1: perturb the coordinates, 2: and the consumed tape slots
    const bool want_kappa = kappa_mode != 0;
    FILE *     fi         = fopen(argv[5], "rb");
    if (!fi) {
      perror("in.bin");
      return 2;
    }
    int64_t hdr[3];
    if (fread(hdr, sizeof(int64_t), 3, fi) != 3) return 2;
    const int64_t n = hdr[0], du = hdr[1], dt = hdr[2];
    std::vector<double> in(n * (du + dt));
    if ((int64_t)fread(in.data(), sizeof(double), in.size(), fi) != (int64_t)in.size()) return 2;
    fclose(fi);

    std::vector<Rec> out(n);
    const auto       clk0 = std::chrono::steady_clock::now();
    for (int64_t k = 0; k < n; ++k) {
      const double *   row = &in[k * (du + dt)];
      gra::AuxIntData aux;
      Rec &            r = out[k];
      memset(&r, 0, sizeof(Rec));
      const double W = eval_point(g, row, du, dt, aux);
      r.W            = W;
      r.kin_ok       = aux.kinematics_ok;
      r.fid_ok       = aux.fidcuts_ok;
      r.veto_ok      = aux.vetocuts_ok;
      r.amp_ok       = aux.amplitude_ok;
      r.tape_used    = (double)g_tape_pos;
      const auto &l  = g.proc->lts;
      r.t1 = l.t1; r.t2 = l.t2; r.s1 = l.s1; r.s2 = l.s2; r.m2 = l.m2; r.Y = l.Y; r.Pt = l.Pt;
      r.t_hat = l.t_hat; r.u_hat = l.u_hat;
      r.nhamp = (double)l.hamp.size();
      r.DW    = l.DW.Integral();
      for (size_t h = 0; h < l.hamp.size() && h < (size_t)NHAMP; ++h) {
        r.hamp[2 * h]     = l.hamp[h].real();
        r.hamp[2 * h + 1] = l.hamp[h].imag();
      }
      for (size_t i = 0; i < l.pfinal.size() && i < (size_t)NPART; ++i) {
        r.p[4 * i + 0] = l.pfinal[i].Px();
        r.p[4 * i + 1] = l.pfinal[i].Py();
        r.p[4 * i + 2] = l.pfinal[i].Pz();
        r.p[4 * i + 3] = l.pfinal[i].E();
      }
      r.kappa = -1.0;
      if (want_kappa && W != 0.0) {
        // 1-ulp conditioning of the reference itself: move each u_k by +-1 ulp
        double kap = 0.0;
        for (int64_t d = 0; d < du; ++d) {
          for (int sgn = -1; sgn <= 1; sgn += 2) {
            std::vector<double> u(row, row + du);
            u[d] = std::nextafter(u[d], sgn > 0 ? 2.0 : -1.0);
            gra::AuxIntData a2;
            const double    W2 = eval_point(g, row, du, dt, a2, &u);
            if (W2 != 0.0) kap = std::max(kap, std::abs(W2 - W) / std::abs(W));
          }
        }
        // kappa mode 2: one step of the uniform grid (2^-53, what one ulp of 1.0 is -- a relative ulp of a small u is
        // swallowed by the first affine map, e.g. y = Y_min + dY u) in every coordinate and in every internal uniform the
        // point consumed (Breit-Wigner masses, decay angles)
        if (kappa_mode >= 2) {
          for (int64_t d = 0; d < du; ++d) {
            for (int sgn = -1; sgn <= 1; sgn += 2) {
              std::vector<double> u(row, row + du);
              u[d] += sgn * 0x1.0p-53;
              gra::AuxIntData a2;
              const double    W2 = eval_point(g, row, du, dt, a2, &u);
              if (W2 != 0.0) kap = std::max(kap, std::abs(W2 - W) / std::abs(W));
            }
          }
          const int64_t used = std::min<int64_t>((int64_t)r.tape_used, dt);
          for (int64_t d = 0; d < used; ++d) {
            for (int sgn = -1; sgn <= 1; sgn += 2) {
              std::vector<double> t(row + du, row + du + dt);
              t[d] += sgn * 0x1.0p-53;
              gra::AuxIntData a2;
              const double    W2 = eval_point(g, row, du, dt, a2, nullptr, &t);
              if (W2 != 0.0) kap = std::max(kap, std::abs(W2 - W) / std::abs(W));
            }
          }
        }
        r.kappa = kap;
      }
    }
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - clk0).count();
    printf("TIME %.6f %lld\n", secs, (long long)n);  // seconds spent in the EventWeight loop (incl. kappa calls)
    FILE *  fo      = fopen(argv[6], "wb");
    int64_t ohdr[2] = {n, RECLEN};
    fwrite(ohdr, sizeof(int64_t), 2, fo);
    fwrite(out.data(), sizeof(Rec), n, fo);
    fclose(fo);
    fprintf(stderr, "ref_eval: %lld points, tape overruns %lld\n", (long long)n, (long long)g_overrun);
  } catch (const std::exception &e) {
    fprintf(stderr, "ref_eval: exception: %s\n", e.what());
    return 1;
  } catch (...) {
    fprintf(stderr, "ref_eval: unknown exception\n");
    return 1;
  }
  return 0;
}
