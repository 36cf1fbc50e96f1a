// Minimal stand-in for the HepMC3 3.2.0 API surface that the GRANIITTI reference library uses
// (13 symbols; `grep HepMC3:: src include` in the reference). The real tarball is absent from
// /root/reference (.MISSING_LARGE_BLOBS), so the oracle build links against this header-only stub.
// Events are discarded: the oracle harness only needs per-point weights, never event files.
// TEST INFRASTRUCTURE ONLY (part of oracle/): never linked into the product library.
#pragma once
#include <memory>
#include <string>
#include <vector>
namespace HepMC3 {
struct Units {
  enum MomentumUnit { MEV, GEV };
  enum LengthUnit { MM, CM };
};
class FourVector {
 public:
  FourVector() : x_(0), y_(0), z_(0), t_(0) {}
  FourVector(double x, double y, double z, double t) : x_(x), y_(y), z_(z), t_(t) {}
  double px() const { return x_; }
  double py() const { return y_; }
  double pz() const { return z_; }
  double e() const { return t_; }
  double x() const { return x_; }
  double y() const { return y_; }
  double z() const { return z_; }
  double t() const { return t_; }
  double x_, y_, z_, t_;
};
class Attribute {
 public:
  virtual ~Attribute() {}
};
class GenCrossSection : public Attribute {
 public:
  void set_cross_section(double a, double b) { xs = a; err = b; }
  double xs = 0, err = 0;
};
class GenVertex;
class GenParticle {
 public:
  GenParticle(const FourVector& p = FourVector(), int pid = 0, int status = 0)
      : p_(p), pid_(pid), status_(status) {}
  const FourVector& momentum() const { return p_; }
  int pid() const { return pid_; }
  int status() const { return status_; }
  FourVector p_;
  int pid_;
  int status_;
};
using GenParticlePtr = std::shared_ptr<GenParticle>;
class GenVertex {
 public:
  GenVertex(const FourVector& p = FourVector()) : pos(p) {}
  void add_particle_in(GenParticlePtr p) { in.push_back(p); }
  void add_particle_out(GenParticlePtr p) { out.push_back(p); }
  std::vector<GenParticlePtr> in, out;
  FourVector pos;
};
using GenVertexPtr = std::shared_ptr<GenVertex>;
class GenRunInfo {
 public:
  struct ToolInfo {
    std::string name, version, description;
  };
  std::vector<ToolInfo>& tools() { return t_; }
  std::vector<ToolInfo> t_;
};
class GenEvent {
 public:
  GenEvent(Units::MomentumUnit = Units::GEV, Units::LengthUnit = Units::MM) {}
  void add_vertex(GenVertexPtr v) { vtx.push_back(v); }
  void add_attribute(const std::string&, std::shared_ptr<Attribute> a) { att.push_back(a); }
  std::vector<double>& weights() { return w; }
  std::vector<GenVertexPtr> vtx;
  std::vector<std::shared_ptr<Attribute>> att;
  std::vector<double> w;
};
class WriterAscii {
 public:
  WriterAscii(const std::string&, std::shared_ptr<GenRunInfo> = nullptr) {}
  void write_event(const GenEvent&) {}
  void close() {}
};
class WriterAsciiHepMC2 : public WriterAscii {
 public:
  using WriterAscii::WriterAscii;
};
class WriterHEPEVT {
 public:
  WriterHEPEVT(const std::string&) {}
  void write_event(const GenEvent&) {}
  void close() {}
};
struct Print {
  static void listing(const GenEvent&) {}
  static void content(const GenEvent&) {}
};
}  // namespace HepMC3
