// grb_events.cc -- event record + writer: at 1e8 - 1e9 weights/s the reference's serial
// record construction and ASCII writer under a mutex (MProcess::CommonRecord, src/MProcess.cc:1588-1706;
// MGraniitti::SaveEvent, src/MGraniitti.cc:1651-1726) is the generation bottleneck.  Here the accepted events arrive as
// compacted SoA momenta (grb_get_events); the reference's record -- beams (status 4), outgoing protons / N* systems,
// the two propagators (id 99, status 81), the central system (id 90, status 81), the decay tree depth first with one
// vertex per decay -- is built and formatted by several host threads into per-thread buffers, which one writer appends
// in event order.  Host-only code: no CUDA here.
//
// Formats: the ASCII layouts of HepMC3 3.2's WriterAscii ("hepmc3"), WriterAsciiHepMC2 ("hepmc2") and WriterHEPEVT
// ("hepevt"), which the reference instantiates (src/MGraniitti.cc:180-209).  HepMC3 itself is absent from this image
// (the reference's tarball is a missing blob), so the layouts follow the library's documented formats and were not
// byte-compared against it; the RECORD CONTENT (particles, ids, statuses, mothers, momenta, vertices) is pinned against
// the reference's own CommonRecord on the same events (tests/test_events.py).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/grb.h"

namespace {

// reference include/Graniitti/MPDG.h:41-50
constexpr int kNSTAR = 90210, kSystem = 90, kPropagator = 99;
constexpr int kBeam = 4, kStable = 1, kDecay = 2, kIntermediate = 81;
constexpr double kHbar = 6.582119514E-25, kC = 2.99792458E8;  // GeV s, m / s (MPDG.h, M4Vec.h:310)

struct V4 { double x, y, z, e; };
inline double m_of(const V4 &a) {  // HepMC3::FourVector::m(): signed square root of the invariant mass squared
  const double m2 = a.e * a.e - (a.x * a.x + a.y * a.y + a.z * a.z);
  return m2 < 0.0 ? -std::sqrt(-m2) : std::sqrt(m2);
}
inline uint64_t mix64(uint64_t z) {  // splitmix64 finaliser: counter-based draw for the decay lifetimes
  z += 0x9E3779B97F4A7C15ull; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

struct Builder {
  const grb_process_desc *d;
  grb_rec_particle *p;
  grb_rec_vertex *  v;
  int               np = 0, nv = 0;
  int add_particle(const V4 &q, int pid, int status, int parent) {
    grb_rec_particle &o = p[np];
    o.id = np + 1; o.pid = pid; o.status = status; o.parent = parent;
    o.px = q.x; o.py = q.y; o.pz = q.z; o.e = q.e; o.m = m_of(q);
    return ++np;
  }
  int add_vertex(int in0, int in1, const double pos[4]) {
    grb_rec_vertex &o = v[nv];
    memset(&o, 0, sizeof(o));
    o.id = -(nv + 1); o.n_in = in1 > 0 ? 2 : 1; o.in[0] = in0; o.in[1] = in1;
    if (pos) { o.x = pos[0]; o.y = pos[1]; o.z = pos[2]; o.t = pos[3]; }
    return -(++nv);
  }
};

// MProcess::WriteDecayKinematics (src/MProcess.cc:1287-1314): one vertex per decaying particle, daughters depth first
void write_decay(Builder &B, const V4 *node_p4, int node, int mother_id, const double *life_u) {
  const grb_node &n = B.d->nodes[node];
  if (n.n_children == 0) return;
  double pos[4] = {0, 0, 0, 0};
  if (life_u && n.width > 0.0) {
    // CommonRecord (:1676-1686): exponential proper lifetime of mean hbar / width, M4Vec::PropagatePosition(tau, 1e3 mm / m)
    const V4 &q = node_p4[node];
    const double m = m_of(q), pmod = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z);
    const double tau0 = -(kHbar / n.width) * std::log(1.0 - life_u[node]);
    const double gamma = q.e / m, tau = gamma * tau0, beta = pmod / q.e;
    if (pmod > 0.0 && m > 0.0) {
      pos[0] = tau * kC * beta * q.x / pmod * 1e3; pos[1] = tau * kC * beta * q.y / pmod * 1e3;
      pos[2] = tau * kC * beta * q.z / pmod * 1e3; pos[3] = tau * kC * 1e3;
    }
  }
  const int vid = B.add_vertex(mother_id, 0, pos);
  for (int c = 0; c < n.n_children; ++c) {
    const int k = n.first_child + c;
    const int id = B.add_particle(node_p4[k], B.d->nodes[k].pdg, B.d->nodes[k].n_children ? kDecay : kStable, vid);
    B.v[-vid - 1].n_out++;
    write_decay(B, node_p4, k, id, life_u);
  }
}

int build_record(const grb_process_desc *d, const double *mom, int beam1_pdg, int beam2_pdg, const double *life_u,
                 grb_rec_particle *p, int *np, grb_rec_vertex *v, int *nv) {
  const int npart = 3 + d->n_nodes;
  const V4 *q     = reinterpret_cast<const V4 *>(mom);  // [pX, p1, p2, nodes...] x (px, py, pz, E)
  Builder   B{d, p, v};
  const V4  pb1{0.0, 0.0, d->pbeam1_pz, d->pbeam1_E}, pb2{0.0, 0.0, d->pbeam2_pz, d->pbeam2_E};
  // excited forward systems: PDG_NSTAR with the beam's sign, status 81 (lts.excite1/2; their fragmentation is downstream)
  const bool ex1 = d->excitation > 0 && m_of(q[1]) > d->beam1_mass + 1e-3, ex2 = d->excitation > 0 && m_of(q[2]) > d->beam2_mass + 1e-3;
  const V4   q1{pb1.x - q[1].x, pb1.y - q[1].y, pb1.z - q[1].z, pb1.e - q[1].e};
  const V4   q2{pb2.x - q[2].x, pb2.y - q[2].y, pb2.z - q[2].z, pb2.e - q[2].e};
  // particle ids follow HepMC3::GenEvent::add_vertex (incoming, then outgoing particles of each vertex as it is attached)
  const int v1 = B.add_vertex(1, 0, nullptr);
  B.add_particle(pb1, beam1_pdg, kBeam, 0);
  B.add_particle(q[1], ex1 ? kNSTAR * (beam1_pdg < 0 ? -1 : 1) : beam1_pdg, ex1 ? kIntermediate : kStable, v1);
  const int iq1 = B.add_particle(q1, kPropagator, kIntermediate, v1);
  B.v[0].n_out = 2;
  const int v2 = B.add_vertex(4, 0, nullptr);
  B.add_particle(pb2, beam2_pdg, kBeam, 0);
  B.add_particle(q[2], ex2 ? kNSTAR * (beam2_pdg < 0 ? -1 : 1) : beam2_pdg, ex2 ? kIntermediate : kStable, v2);
  const int iq2 = B.add_particle(q2, kPropagator, kIntermediate, v2);
  B.v[1].n_out = 2;
  const int v3 = B.add_vertex(iq1, iq2, nullptr);
  const int isys = B.add_particle(q[0], kSystem, kIntermediate, v3);
  B.v[2].n_out = 1;
  const int v4 = B.add_vertex(isys, 0, nullptr);
  for (int i = 0; i < d->n_legs; ++i) {
    const int id = B.add_particle(q[3 + i], d->nodes[i].pdg, d->nodes[i].n_children ? kDecay : kStable, v4);
    B.v[-v4 - 1].n_out++;
    write_decay(B, q + 3, i, id, life_u);
  }
  (void)npart;
  *np = B.np;
  *nv = B.nv;
  return GRB_OK;
}

// ---- text layouts ------------------------------------------------------------------------------------------------------
struct Fmt {
  std::string out;
  char        buf[512];
  template <class... A> void add(const char *f, A... a) { out.append(buf, (size_t)snprintf(buf, sizeof(buf), f, a...)); }
};

// HepMC3 WriterAscii: the vertex line is written only for vertices with several incoming particles or a position; a
// particle produced by a single mother without a vertex line carries the mother's particle id in its second column
void emit_hepmc3(Fmt &F, const grb_rec_particle *p, int np, const grb_rec_vertex *v, int nv, int evno, double weight, double xs_pb, double xs_err_pb) {
  F.add("E %d %d %d\n", evno, nv, np);
  F.add("%s", "U GEV MM\n");
  F.add("W %.16e\n", weight);
  F.add("A 0 GenCrossSection %.8e %.8e -1 -1\n", xs_pb, xs_err_pb);
  std::vector<char> written((size_t)nv, 0);
  for (int i = 0; i < np; ++i) {
    int parent = 0;
    if (p[i].parent < 0) {
      const grb_rec_vertex &vx = v[-p[i].parent - 1];
      const bool has_pos = vx.x != 0.0 || vx.y != 0.0 || vx.z != 0.0 || vx.t != 0.0;
      if (vx.n_in > 1 || has_pos) {
        parent = vx.id;
        if (!written[-vx.id - 1]) {
          written[-vx.id - 1] = 1;
          if (vx.n_in > 1) F.add("V %d 0 [%d,%d]", vx.id, std::min(vx.in[0], vx.in[1]), std::max(vx.in[0], vx.in[1]));
          else F.add("V %d 0 [%d]", vx.id, vx.in[0]);
          if (has_pos) F.add(" @ %.16e %.16e %.16e %.16e\n", vx.x, vx.y, vx.z, vx.t);
          else F.add("%s", "\n");
        }
      } else
        parent = vx.in[0];
    }
    F.add("P %d %d %d %.16e %.16e %.16e %.16e %.16e %d\n", p[i].id, parent, p[i].pid, p[i].px, p[i].py, p[i].pz, p[i].e, p[i].m, p[i].status);
  }
}

// HepMC2 (IO_GenEvent) through HepMC3's WriterAsciiHepMC2: vertices with their outgoing particles, barcodes = ids
void emit_hepmc2(Fmt &F, const grb_rec_particle *p, int np, const grb_rec_vertex *v, int nv, int evno, double weight, double xs_pb, double xs_err_pb) {
  F.add("E %d -1 -1.0000000000000000e+00 -1.0000000000000000e+00 -1.0000000000000000e+00 0 0 %d 1 4 0 1 %.16e\n", evno, nv, weight);
  F.add("%s", "N 1 \"0\"\n");
  F.add("%s", "U GEV MM\n");
  F.add("C %.16e %.16e\n", xs_pb, xs_err_pb);
  std::vector<int> end_vertex((size_t)np + 1, 0);
  for (int k = 0; k < nv; ++k)
    for (int j = 0; j < v[k].n_in; ++j) end_vertex[(size_t)v[k].in[j]] = v[k].id;
  for (int k = 0; k < nv; ++k) {
    // orphan incoming particles (no production vertex) are listed with their end vertex: the two beams
    int orphans = 0;
    for (int j = 0; j < v[k].n_in; ++j) orphans += p[v[k].in[j] - 1].parent == 0;
    F.add("V %d 0 %.16e %.16e %.16e %.16e %d %d 0\n", v[k].id, v[k].x, v[k].y, v[k].z, v[k].t, orphans, v[k].n_out);
    auto line = [&](const grb_rec_particle &q) {
      F.add("P %d %d %.16e %.16e %.16e %.16e %.16e %d 0 0 %d 0\n", 10000 + q.id, q.pid, q.px, q.py, q.pz, q.e, q.m, q.status, end_vertex[(size_t)q.id]);
    };
    for (int j = 0; j < v[k].n_in; ++j)
      if (p[v[k].in[j] - 1].parent == 0) line(p[v[k].in[j] - 1]);
    for (int i = 0; i < np; ++i)
      if (p[i].parent == v[k].id) line(p[i]);
  }
}

// HEPEVT through HepMC3's WriterHEPEVT: one block per event, mothers / daughters as index ranges
void emit_hepevt(Fmt &F, const grb_rec_particle *p, int np, const grb_rec_vertex *v, int nv, int evno) {
  F.add("E % -8i %-8i\n", evno, np);
  for (int i = 0; i < np; ++i) {
    int mo1 = 0, mo2 = 0, da1 = 0, da2 = 0;
    double pos[4] = {0, 0, 0, 0};
    if (p[i].parent < 0) {
      const grb_rec_vertex &vx = v[-p[i].parent - 1];
      mo1 = vx.in[0]; mo2 = vx.n_in > 1 ? vx.in[1] : vx.in[0];
      if (mo1 > mo2) std::swap(mo1, mo2);
      pos[0] = vx.x; pos[1] = vx.y; pos[2] = vx.z; pos[3] = vx.t;
    }
    for (int k = 0; k < np; ++k)
      if (p[k].parent < 0) {
        const grb_rec_vertex &vx = v[-p[k].parent - 1];
        if (vx.in[0] == p[i].id || (vx.n_in > 1 && vx.in[1] == p[i].id)) {
          if (!da1) da1 = p[k].id;
          da2 = p[k].id;
        }
      }
    F.add("% 8i% 8i% 8i% 8i% 8i% 8i% 19.8E% 19.8E% 19.8E% 19.8E% 19.8E\n", p[i].status, p[i].pid, mo1, mo2, da1, da2, p[i].px, p[i].py, p[i].pz, p[i].e, p[i].m);
    F.add("%-48s% 19.8E% 19.8E% 19.8E% 19.8E\n", " ", pos[0], pos[1], pos[2], pos[3]);
  }
  (void)nv;
}

int format_id(const char *f) {
  if (!f) return -1;
  if (!strcmp(f, "hepmc3")) return 0;
  if (!strcmp(f, "hepmc2")) return 1;
  if (!strcmp(f, "hepevt")) return 2;
  return -1;
}

}  // namespace

extern "C" {

int grb_event_record(const grb_process_desc *d, const double *momenta, int beam1_pdg, int beam2_pdg, const double *lifetime_u,
                     grb_rec_particle *particles, int *n_particles, grb_rec_vertex *vertices, int *n_vertices) {
  if (!d || !momenta || !particles || !n_particles || !vertices || !n_vertices) return GRB_ERR_ARG;
  if (d->n_nodes < 1 || d->n_nodes > GRB_MAX_NODES || d->n_legs < 1 || d->n_legs > d->n_nodes) return GRB_ERR_ARG;
  return build_record(d, momenta, beam1_pdg ? beam1_pdg : 2212, beam2_pdg ? beam2_pdg : 2212, lifetime_u, particles, n_particles,
                      vertices, n_vertices);
}

int64_t grb_write_events(const grb_process_desc *d, const char *path, const grb_writer_opts *o, int64_t n, const double *momenta,
                         const double *weights) {
  if (!d || !path || !o || n < 0 || (n > 0 && !momenta)) return GRB_ERR_ARG;
  const int fmt = format_id(o->format);
  if (fmt < 0) return GRB_ERR_ARG;  // reference: "MGraniitti::SaveEvent: Unknown output FORMAT"
  const int npart = 3 + d->n_nodes;
  const int b1 = o->beam1_pdg ? o->beam1_pdg : 2212, b2 = o->beam2_pdg ? o->beam2_pdg : 2212;
  FILE *fo = fopen(path, o->append ? "ab" : "wb");
  if (!fo) return GRB_ERR_ARG;
  if (!o->append) {  // run header: the two tools the reference registers (generator, steering card), src/MGraniitti.cc:190-200
    if (fmt == 0) {
      fprintf(fo, "HepMC::Version 3.02.00\nHepMC::Asciiv3-START_EVENT_LISTING\n");
      fprintf(fo, "T %s\\|%s\\|Generator\n", o->generator_name ? o->generator_name : "GRANIITTI (TUNE0)", o->generator_version ? o->generator_version : "1.051");
      fprintf(fo, "T %s\\|1.0\\|Steering card\n", o->card_path ? o->card_path : "");
    } else if (fmt == 1)
      fprintf(fo, "HepMC::Version 3.02.00\nHepMC::IO_GenEvent-START_EVENT_LISTING\n");
  }
  int nthreads = o->nthreads > 0 ? o->nthreads : (int)std::max(1u, std::thread::hardware_concurrency());
  const int64_t block = 4096;   // events per formatted buffer: the writer appends whole buffers in event order
  const int64_t nblocks = (n + block - 1) / block;
  nthreads = (int)std::max<int64_t>(1, std::min<int64_t>(nthreads, nblocks));
  int64_t written = 0;
  for (int64_t b0 = 0; b0 < nblocks; b0 += nthreads) {
    const int nb = (int)std::min<int64_t>(nthreads, nblocks - b0);
    std::vector<std::string> bufs((size_t)nb);
    auto work = [&](int t) {
      Fmt F;
      F.out.reserve((size_t)block * (size_t)(npart + 6) * 160);
      grb_rec_particle P[GRB_MAX_REC_PARTICLES];
      grb_rec_vertex   V[GRB_MAX_REC_VERTICES];
      double           mom[(3 + GRB_MAX_NODES) * 4], life[GRB_MAX_NODES];
      const int64_t lo = (b0 + t) * block, hi = std::min<int64_t>(n, lo + block);
      for (int64_t i = lo; i < hi; ++i) {
        for (int r = 0; r < 4 * npart; ++r) mom[r] = momenta[(int64_t)r * n + i];   // SoA [npart][4][n] -> one event
        for (int k = 0; k < d->n_nodes; ++k)
          life[k] = (double)(mix64(o->lifetime_seed ^ mix64((uint64_t)(o->first_event + i) * GRB_MAX_NODES + (uint64_t)k)) >> 11) * 0x1.0p-53;
        int np = 0, nv = 0;
        build_record(d, mom, b1, b2, o->lifetimes ? life : nullptr, P, &np, V, &nv);
        const double w = (o->weighted && weights) ? weights[i] : 1.0;   // unweighted events carry weight 1 (SaveEvent)
        const int evno = o->number_events ? (int)(o->first_event + i) : 0;
        if (fmt == 0) emit_hepmc3(F, P, np, V, nv, evno, w, o->sigma_pb, o->sigma_err_pb);
        else if (fmt == 1) emit_hepmc2(F, P, np, V, nv, evno, w, o->sigma_pb, o->sigma_err_pb);
        else emit_hepevt(F, P, np, V, nv, evno);
      }
      bufs[(size_t)t].swap(F.out);
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nb; ++t) th.emplace_back(work, t);
    work(0);
    for (auto &x : th) x.join();
    for (int t = 0; t < nb; ++t) {
      if (fwrite(bufs[(size_t)t].data(), 1, bufs[(size_t)t].size(), fo) != bufs[(size_t)t].size()) { fclose(fo); return GRB_ERR_ARG; }
      written += std::min<int64_t>(n, (b0 + t + 1) * block) - (b0 + t) * block;
    }
  }
  if (o->close_listing) {
    if (fmt == 0) fprintf(fo, "HepMC::Asciiv3-END_EVENT_LISTING\n\n");
    else if (fmt == 1) fprintf(fo, "HepMC::IO_GenEvent-END_EVENT_LISTING\n\n");
  }
  fclose(fo);
  return written;
}

}  // extern "C"
