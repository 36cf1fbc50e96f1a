// grb_host.cc -- host-side (C++) mirror of the reference's steering interface for the hot path.
//
// The reference keeps its JSON steering card, the process mini-language
//     ISTATE[CHANNEL]<class> -> final > {daughters} @CMD[target]{k:v}
// and the model tune directory (reference src/MGraniitti.cc:317-963, src/MProcess.cc:380-468,1732-1853,
// src/MPDG.cc:20-560, src/MRegge.cc:984-1033, src/MForm.cc:71-209).  This file reads exactly those inputs
// and flattens them into the POD grb_process_desc that crosses the C ABI, so a steering card written for
// `gr` drives the device engine unchanged.  Names, argument meaning and error behaviour follow the
// reference (std::invalid_argument there -> GRB_ERR_* + message here).
#include "grb_host.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "grb_builtin_data.h"
#include "grb_json5.h"

namespace grbhost {

struct Unsupported : std::runtime_error {
  using std::runtime_error::runtime_error;
};

static std::string read_file(const std::string &path) {
  std::ifstream f(path);
  if (!f.is_open()) throw std::invalid_argument("cannot open " + path);
  std::stringstream ss;
  ss << f.rdbuf();
  return ss.str();
}

static std::string trim(const std::string &s) {
  size_t a = 0, b = s.size();
  while (a < b && isspace((unsigned char)s[a])) ++a;
  while (b > a && isspace((unsigned char)s[b - 1])) --b;
  return s.substr(a, b - a);
}

static std::vector<std::string> split(const std::string &s, char sep) {
  std::vector<std::string> out;
  std::string              cur;
  for (char c : s) {
    if (c == sep) {
      out.push_back(cur);
      cur.clear();
    } else {
      cur += c;
    }
  }
  out.push_back(cur);
  return out;
}

// ---- particle table (reference src/MPDG.cc:20-440 ReadParticleData) --------------------------------------
struct Particle {
  int         pdg = 0, spinX2 = 0, chargeX3 = 0;
  double      mass = 0.0, width = 0.0;
  std::string name;
};

class ParticleTable {
 public:
  void add_entry(const std::vector<int> &ids, double mass, double width, const std::string &name,
                 const std::vector<int> &charges) {
    for (size_t k = 0; k < ids.size() && k < charges.size(); ++k) {
      Particle p;
      p.pdg = ids[k]; p.mass = mass; p.width = width; p.chargeX3 = charges[k]; p.name = name;
      p.spinX2 = (p.pdg % 10) - 1;
      const bool quark_or_nu = (p.pdg >= 1 && p.pdg <= 6) || p.pdg == 12 || p.pdg == 14 || p.pdg == 16;
      if (p.chargeX3 == 0 && p.pdg > 100) {
        p.name = name + "0";
      } else if (p.chargeX3 != 0 && !quark_or_nu) {
        p.name = name + std::string(std::abs(p.chargeX3 / 3), p.chargeX3 > 0 ? '+' : '-');
      }
      if (p.pdg == 21 || p.pdg == 22 || std::abs(p.pdg) == 24 || p.pdg == 23) p.spinX2 = 2;
      if (p.pdg == 25) p.spinX2 = 0;
      if (p.pdg >= 1 && p.pdg <= 16 && p.pdg != 9 && p.pdg != 10) p.spinX2 = 1;  // SM fermions
      tab[p.pdg] = p;
      const bool fermion = (p.spinX2 % 2) != 0;
      if ((fermion && ids.size() < 3) || (p.chargeX3 != 0 && ids.size() < 3)) {
        Particle a = p;
        if (quark_or_nu) a.name = name + "~";
        else if (p.chargeX3 != 0) a.name = name + std::string(std::abs(p.chargeX3 / 3), p.chargeX3 > 0 ? '-' : '+');
        else a.name = name + "0~";
        a.pdg = -p.pdg;
        a.chargeX3 = -p.chargeX3;
        tab[a.pdg] = a;
      }
    }
  }
  static std::vector<int> parse_charges(const std::string &s) {
    std::vector<int> q;
    for (auto &t : split(s, ',')) {
      const std::string c = trim(t);
      if (c.find("++") != std::string::npos) q.push_back(6);
      else if (c.find("--") != std::string::npos) q.push_back(-6);
      else if (c.find("-1/3") != std::string::npos) q.push_back(-1);
      else if (c.find("+2/3") != std::string::npos) q.push_back(2);
      else if (c.find("+") != std::string::npos) q.push_back(3);
      else if (c.find("-") != std::string::npos) q.push_back(-3);
      else if (c.find("0") != std::string::npos) q.push_back(0);
    }
    return q;
  }
  void load_builtin() {
    for (auto &b : kBuiltinPDG) {
      std::vector<int> ids = {b.pdg};
      if (b.pdg2) ids.push_back(b.pdg2);
      add_entry(ids, b.mass, b.width, b.name, parse_charges(b.charges));
    }
  }
  // PDG "mass_width" fixed-column format: ids in columns 1-32, mass 34-51, width 71-88, name+charges 108-128
  void load_mcd(const std::string &path) {
    std::ifstream f(path);
    if (!f.is_open()) throw std::invalid_argument("cannot open PDG table " + path);
    std::string line;
    while (std::getline(f, line)) {
      if (line.empty() || line[0] == '*') continue;
      if (line.size() < 108) line.resize(128, ' ');
      std::vector<int>   ids;
      std::istringstream si(line.substr(0, 32));
      int                v;
      while (si >> v)
        if (v != 0) ids.push_back(v);
      if (ids.empty()) continue;
      double             mass = 0.0, width = 0.0;
      std::istringstream sm(line.substr(33, 18)), sw(line.substr(70, 18));
      sm >> mass;
      sw >> width;
      std::istringstream sn(line.substr(107));
      std::string        name, rest;
      sn >> name;
      std::getline(sn, rest);
      add_entry(ids, mass, width, name, parse_charges(rest));
    }
  }
  const Particle &by_name(const std::string &n) const {
    for (auto &kv : tab)
      if (kv.second.name == n) return kv.second;
    // PDG number input is allowed too (reference src/MPDG.cc:572-576, aux::IsIntegerDigits)
    if (!n.empty() && n.find_first_not_of("-0123456789") == std::string::npos) {
      auto it = tab.find(std::stoi(n));
      if (it != tab.end()) return it->second;
      throw std::invalid_argument("MPDG::FindByPDG: Unknown PDG ID: " + n);
    }
    throw std::invalid_argument("MPDG::FindByPDGName: Unknown particle name: " + n);
  }
  const Particle &by_pdg(int pdg) const {
    auto it = tab.find(pdg);
    if (it == tab.end()) throw std::invalid_argument("MPDG::FindByPDG: Unknown PDG ID: " + std::to_string(pdg));
    return it->second;
  }
  std::map<int, Particle> tab;
};

// ---- @ command blocks (reference src/MAux.cc:690-811 SplitCommands) -----------------------------------------
struct OneCmd {
  std::string                        id;
  std::vector<std::string>           target;
  std::map<std::string, std::string> arg;
};

static std::vector<OneCmd> split_commands(const std::string &full) {
  std::vector<OneCmd> out;
  size_t              pos = full.find('@');
  while (pos != std::string::npos) {
    const size_t next = full.find('@', pos + 1);
    std::string  sub  = trim(full.substr(pos + 1, next == std::string::npos ? std::string::npos : next - pos - 1));
    OneCmd       c;
    const size_t l = sub.find('['), r = sub.find(']');
    if ((l == std::string::npos) != (r == std::string::npos))
      throw std::invalid_argument("SplitCommands: Incomplete @R[]{} syntax with missing [ or ]");
    if (l != std::string::npos) {
      const std::string t = sub.substr(l + 1, r - l - 1);
      if (!t.empty())
        for (auto &x : split(t, ',')) c.target.push_back(trim(x));
      sub.erase(l, r - l + 1);
    }
    const size_t L = sub.find('{'), R = sub.find('}');
    if ((L == std::string::npos) != (R == std::string::npos))
      throw std::invalid_argument("SplitCommands: Incomplete @R[]{} syntax with missing { or }");
    if (L == std::string::npos) {
      auto strip = split(sub, ':');
      c.id       = trim(strip[0]);
      if (strip.size() == 1) c.arg["_SINGLET_"] = "true";
      else if (strip.size() == 2) c.arg["_SINGLET_"] = trim(strip[1]);
      else throw std::invalid_argument("gra::SplitCommands: @Syntax invalid with '" + sub + "'");
    } else {
      c.id = trim(sub.substr(0, L));
      for (auto &kv : split(sub.substr(L + 1, R - L - 1), ',')) {
        auto strip = split(kv, ':');
        if (strip.size() != 2) throw std::invalid_argument("gra::SplitCommands: @Syntax not good with brackets {" + kv + "}");
        c.arg[trim(strip[0])] = trim(strip[1]);
      }
    }
    out.push_back(c);
    pos = next;
  }
  return out;
}

// ---- decay string -> flattened tree (reference src/MPDG.cc:446-532 TokenizeProcess) -------------------------
struct Branch {
  Particle            p;
  std::vector<Branch> legs;
};

static std::vector<std::string> words(const std::string &s) {
  std::vector<std::string> w;
  std::istringstream       ss(s);
  std::string              t;
  while (ss >> t) w.push_back(t);
  return w;
}

static void tokenize(const std::string &str, const ParticleTable &pdg, std::vector<Branch> &out) {
  size_t i = 0, start = 0;
  while (i < str.size()) {
    if (str[i] == '>') {
      // particles before '>' : the last one is the mother
      for (auto &w : words(str.substr(start, i - start))) {
        Branch b;
        b.p = pdg.by_name(w);
        out.push_back(b);
      }
      if (out.empty()) throw std::invalid_argument("decay arrow '>' without a mother particle");
      // matching braces
      size_t a = str.find('{', i);
      if (a == std::string::npos) throw std::invalid_argument("decay arrow '>' without { }");
      int    depth = 0;
      size_t b     = a;
      for (; b < str.size(); ++b) {
        if (str[b] == '{') ++depth;
        if (str[b] == '}' && --depth == 0) break;
      }
      if (b >= str.size()) throw std::invalid_argument("unbalanced { } in the decay string");
      tokenize(str.substr(a + 1, b - a - 1), pdg, out.back().legs);
      i = start = b + 1;
    } else {
      ++i;
    }
  }
  for (auto &w : words(str.substr(start))) {
    Branch b;
    b.p = pdg.by_name(w);
    out.push_back(b);
  }
}

// breadth-first flattening: first-level legs at [0, n_legs), daughters of one mother contiguous
static void flatten(const std::vector<Branch> &legs, grb_process_desc &d) {
  struct Q {
    const Branch *b;
    int           idx;
  };
  std::vector<Q> queue;
  int            n = 0;
  auto put = [&](const Branch &b, int parent) {
    if (n >= GRB_MAX_NODES) throw Unsupported("decay tree exceeds GRB_MAX_NODES");
    grb_node &o = d.nodes[n];
    o.pdg = b.p.pdg; o.spinX2 = b.p.spinX2; o.parent = parent; o.first_child = -1;
    o.n_children = (int)b.legs.size(); o.mass = b.p.mass; o.width = b.p.width;
    queue.push_back({&b, n});
    return n++;
  };
  for (auto &b : legs) put(b, -1);
  for (size_t q = 0; q < queue.size(); ++q) {
    const Branch *b = queue[q].b;
    const int     i = queue[q].idx;
    if (!b->legs.empty()) {
      d.nodes[i].first_child = n;
      for (auto &c : b->legs) put(c, i);
    }
  }
  d.n_nodes = n;
  d.n_legs  = (int)legs.size();
  d.n_final = 0;
  for (int i = 0; i < n; ++i)
    if (d.nodes[i].n_children == 0) ++d.n_final;
}

static double factorial(int n) {
  double f = 1.0;
  for (int i = 2; i <= n; ++i) f *= i;
  return f;
}

static void range2(const Json &j, const char *what, double &lo, double &hi) {
  auto v = j.as_vec();
  if (v.size() != 2) throw std::invalid_argument(std::string(what) + ": expected [min, max]");
  if (v[0] > v[1]) throw std::invalid_argument(std::string(what) + ": min > max");
  lo = v[0];
  hi = v[1];
}

// ---- resonances (reference form::ReadResonance src/MForm.cc:214-375, MProcess::ProcessHelicityDecay
// src/MProcess.cc:526-748, spin::InitTMatrix src/MSpin.cc:54-245) -------------------------------------------
static double fact(int n) {  // gra::math::factorial, MMath.h:201-205
  double x = 1;
  for (int i = 1; i <= n; ++i) x *= i;
  return x;
}
static bool is_int(double x) { return std::abs(x - std::floor(x)) < 1E-9; }
static bool is_halfint(double x) { return std::abs(2 * x - std::floor(2 * x)) < 1E-9; }
static bool is_eq(double x, double y) { return std::abs(x - y) < 1E-9; }

// Wigner 3j by the Racah formula (reference src/MSpin.cc:800-838) with its selection rules (:841-892)
static double w3j(double j1, double j2, double j3, double m1, double m2, double m3) {
  if (!is_int(j1 + j2 + j3) || !is_eq(m1 + m2 + m3, 0) || !is_eq(j1 - m1, std::floor(j1 - m1)) ||
      !is_eq(j2 - m2, std::floor(j2 - m2)) || !is_eq(j3 - m3, std::floor(j3 - m3)) || j3 > j1 + j2 ||
      j3 < std::abs(j1 - j2) || std::abs(m1) > j1 || std::abs(m2) > j2 || std::abs(m3) > j3)
    return 0.0;
  const int c1 = -j2 + m1 + j3, c2 = -j1 - m2 + j3, c3 = j1 + j2 - j3, c4 = j1 - m1, c5 = j2 + m2;
  const int lower = std::max(std::max(0, -c1), -c2), upper = std::min(std::min(c3, c4), c5);
  double    fsum = 0;
  for (int k = lower; k <= upper; ++k)
    fsum += std::pow(-1.0, k) / (fact(k) * fact(c1 + k) * fact(c2 + k) * fact(c3 - k) * fact(c4 - k) * fact(c5 - k));
  const double tri = fact(j1 + j2 - j3) * fact(j1 - j2 + j3) * fact(-j1 + j2 + j3) / fact(j1 + j2 + j3 + 1);
  return std::pow(-1.0, j1 - j2 - m3) *
         std::sqrt(std::max(0.0, tri * fact(j1 + m1) * fact(j1 - m1) * fact(j2 + m2) * fact(j2 - m2) * fact(j3 + m3) * fact(j3 - m3))) * fsum;
}
// Clebsch-Gordan <j m | j1 j2 m1 m2> (reference src/MSpin.cc:706-713, rules :734-792)
static double clebsch(double j1, double j2, double m1, double m2, double j, double m) {
  if (!is_halfint(j1) || !is_halfint(j2) || !is_halfint(j) || !is_int(j1 + j2 + j) || !is_halfint(m1) ||
      !is_halfint(m2) || !is_halfint(m) || !is_eq(m1 + m2, m) || !is_eq(j1 - m1, std::floor(j1 - m1)) ||
      !is_eq(j2 - m2, std::floor(j2 - m2)) || !is_eq(j - m, std::floor(j - m)) || j > j1 + j2 || j < std::abs(j1 - j2) ||
      std::abs(m1) > j1 || std::abs(m2) > j2 || std::abs(m) > j)
    return 0.0;
  return std::pow(-1.0, m + j1 - j2) * std::sqrt(std::max(0.0, 2 * j + 1)) * w3j(j1, j2, j, m1, m2, -m);
}

static int frame_id(const std::string &f) {
  if (f == "CM") return 0;
  if (f == "HX") return 1;
  if (f == "CS") return 2;
  if (f == "GJ") return 3;
  throw std::invalid_argument("gra::spin::GetRotation: Unknown polarization Lorentz rest FRAME chosen: " + f);
}

// parity of a decay daughter; only (pseudo)scalar mesons are supported as resonance daughters for now
static int daughter_parity(const grb_node &n) {
  if (n.spinX2 != 0) throw Unsupported("resonance decays into daughters with spin are not on the device path yet");
  return -1;  // light pseudoscalar mesons (reference MPDG.cc setPCL table: n0qq1 -> 0-+)
}

// overrides: the {key: value} blocks of the @R[name]{...} commands that target this resonance, in command order.  The
// reference applies them to the freshly read resonance BEFORE SetResonances / SetupBranching (src/MGraniitti.cc:481-578),
// so the decay coupling below is computed from the overridden mass and width (src/MProcess.cc:676-712).
static void fill_resonance(const Json &R, const Json &tune, const Json &branching, const grb_process_desc &d,
                           const std::string &name, const std::vector<const std::map<std::string, std::string> *> &overrides,
                           grb_resonance &o) {
  memset(&o, 0, sizeof(o));
  const Json &sp = tune.at("PARAM_SPIN");
  o.spingen = sp.at("SPINGEN").as_bool(); o.spindec = sp.at("SPINDEC").as_bool();
  o.frame = frame_id(sp.at("FRAME").as_str()); o.jmax = sp.at("JMAX").as_int();
  const double g_A = R.at("g_A").as_num(), g_phi = R.at("g_phi").as_num();
  // g_A * exp(i g_phi): libstdc++ std::exp(complex(0, x)) = (cos x, sin x)
  o.g_re = g_A * std::cos(g_phi); o.g_im = g_A * std::sin(g_phi);
  o.g_FF = R.at("g_FF").as_num();
  o.pdg = R.at("PDG").as_int();
  o.mass = R.at("M").as_num(); o.width = R.at("W").as_num();
  if (o.mass < 0 || o.width < 0) throw std::invalid_argument("MAux:ReadResonance:: <" + name + "> Invalid M or W < 0 !");
  const double J = R.at("J").as_num();
  o.spinX2 = (int)(J * 2);
  o.P = R.at("P").as_int();
  if (!(o.P == -1 || o.P == 1)) throw std::invalid_argument("MAux:ReadResonance:: <" + name + "> Invalid P (not -1 or 1) !");
  auto gT = R.at("g_Tensor").as_vec();
  for (size_t i = 0; i < gT.size() && i < 8; ++i) o.g_Tensor[i] = gT[i];
  o.BW = R.at("BW").as_int();
  if (o.BW < 1 || o.BW > 4) throw std::invalid_argument("CBW: Unknown BW (Breit-Wigner) parameter: " + std::to_string(o.BW));
  const int n = o.spinX2 + 1;
  if (o.spinX2 != 0) {
    if (n > GRB_RHO_DIM) throw Unsupported("resonance spin J > 2 exceeds GRB_RHO_DIM");
    if (R.at("random_rho").as_num() > 0) throw Unsupported("random_rho > 0 (RNG-drawn density matrices) is not supported");
    double tr = 0;
    for (int a = 0; a < n; ++a)
      for (int b = 0; b < n; ++b) {
        o.rho_re[a * GRB_RHO_DIM + b] = R.at("rho_real").arr.at(a).arr.at(b).as_num();
        o.rho_im[a * GRB_RHO_DIM + b] = R.at("rho_imag").arr.at(a).arr.at(b).as_num();
        if (a == b) tr += o.rho_re[a * GRB_RHO_DIM + b];
      }
    if (std::abs(tr - 1.0) > 1e-2) throw std::invalid_argument("gra::spin::Positivity: Density matrix is not properly normalized");
  }
  // ---- @R[name]{M:, W:, g<n>:, JZ<n>:} on-the-flight overrides (src/MGraniitti.cc:482-566); other keys match nothing ----
  for (const auto *ov : overrides) {
    const int N = o.spinX2 != 0 ? n : 1, J = (N - 1) / 2;
    std::vector<double> newrho((size_t)N * N, 0.0);
    bool density_touched = false;
    for (const auto &x : *ov) {
      if (x.first == "M") o.mass = std::stod(x.second);
      if (x.first == "W") o.width = std::stod(x.second);
      for (size_t k = 0; k < gT.size() && k < 8; ++k)
        if (x.first == "g" + std::to_string(k)) o.g_Tensor[k] = std::stod(x.second);
      for (int k = 0; k <= (N - 1) / 2; ++k)
        if (x.first == "JZ" + std::to_string(k)) {
          const double value = std::stod(x.second);
          if (value < 0) throw std::invalid_argument("MGraniitti::ReadProcessParam: @R[] JZ value < 0");
          newrho[(size_t)(N - 1 - J - k) * N + (N - 1 - J - k)] = value;
          newrho[(size_t)(N - 1 - J + k) * N + (N - 1 - J + k)] = value;
          density_touched = true;
        }
    }
    if (density_touched) {
      double trace = 0.0;
      for (int a = 0; a < N; ++a) trace += newrho[(size_t)a * N + a];
      if (!(std::abs(trace) > 0)) throw std::invalid_argument("MGraniitti::ReadProcessParam: @R[] Spin density error with <= 0 trace");
      const double inv = 1.0 / trace;
      for (int a = 0; a < GRB_RHO_DIM * GRB_RHO_DIM; ++a) { o.rho_re[a] = 0.0; o.rho_im[a] = 0.0; }
      for (int a = 0; a < N; ++a)
        for (int b = 0; b < N; ++b) o.rho_re[a * GRB_RHO_DIM + b] = newrho[(size_t)a * N + b] * inv;
    }
  }
  // ---- decay side: BR, alpha_ls -> T matrix, equivalent coupling g_decay --------------------------------
  if (d.n_legs != 2) throw Unsupported("resonances need a 2-body central final state on the device path");
  const std::string key = std::to_string(o.pdg);
  if (!branching.has(key))
    throw std::invalid_argument("MProcess::ProcessHelicityDecay: Did not find any branching ratio data for resonance with PDG: " + key);
  const Json &tab = branching.at(key);
  std::vector<int> want = {d.nodes[0].pdg, d.nodes[1].pdg};
  std::sort(want.begin(), want.end());
  double BR = 1.0;
  bool   Pcons = true, found = false;
  std::vector<double> gdt;
  std::map<std::pair<int, int>, std::pair<double, double>> alpha;  // (l,s) -> (re,im), default 1
  std::map<std::pair<int, int>, bool> alpha_set;
  for (auto &kv : tab.obj) {
    if (kv.second.type != Json::OBJ || !kv.second.has("PDG")) continue;
    std::vector<int> have;
    for (auto &x : kv.second.at("PDG").arr) have.push_back(x.as_int());
    std::sort(have.begin(), have.end());
    if (have != want) continue;
    BR = kv.second.at("BR").as_num();
    Pcons = kv.second.at("P_conservation").as_bool();
    if (kv.second.has("g_decay_tensor")) {   // user-set tensor-pomeron decay couplings (src/MProcess.cc:594-600)
      gdt = kv.second.at("g_decay_tensor").as_vec();
    }
    if (kv.second.has("alpha_ls"))
      for (auto &a : kv.second.at("alpha_ls").arr) {
        const auto key2 = std::make_pair(a.arr.at(0).as_int(), a.arr.at(1).as_int());
        alpha[key2] = {a.arr.at(2).as_num(), a.arr.at(3).as_num()};
        alpha_set[key2] = true;
      }
    found = true;
    break;
  }
  const bool all_set = !found;  // unknown decay mode: BR = 1, alpha_ls == 1 everywhere (reference warning branch)
  o.BR = BR;
  // InitTMatrix for daughters with spins s1, s2 (here 0): T[(2s1+1)][(2s2+1)]
  const double s1 = d.nodes[0].spinX2 / 2.0, s2 = d.nodes[1].spinX2 / 2.0;
  const int    P1 = daughter_parity(d.nodes[0]), P2 = daughter_parity(d.nodes[1]);
  const bool   identical = d.nodes[0].pdg == d.nodes[1].pdg;
  const int    n1 = (int)(2 * s1 + 1), n2 = (int)(2 * s2 + 1);
  if (n1 * n2 > 9) throw Unsupported("T matrix larger than 3x3");
  o.T_rows = n1; o.T_cols = n2;
  std::map<std::pair<int, int>, bool> active;
  for (int MODE = 0; MODE < 2; ++MODE) {
    int nonzero = 0;
    for (int s = 0; s <= (int)(s1 + s2); ++s)
      for (int l = 0; l <= (int)(J + s); ++l)
        for (int i = 0; i < n1; ++i)
          for (int j = 0; j < n2; ++j) {
            const double lambda1 = -s1 + i, lambda2 = -s2 + j, lambda = lambda1 - lambda2;
            if (!(std::abs(lambda1 - lambda2) <= J)) continue;
            const double NORM = std::sqrt((2.0 * l + 1.0) / (2.0 * J + 1.0));
            const double cg1 = clebsch(l, s, 0.0, lambda, J, lambda);
            const double cg2 = clebsch(s1, s2, lambda1, -lambda2, s, lambda);
            if (Pcons && (int)(P1 * P2 * std::pow(-1, l)) != o.P) continue;
            if (identical && ((l - s) % 2 != 0)) continue;  // Bose symmetry for identical bosons
            if (std::abs(cg1 * cg2) < 1e-6) continue;
            active[{l, s}] = true;
            if (MODE == 0) {
              if (!all_set && !alpha_set[{l, s}])
                throw std::invalid_argument("gra::spin::InitTMatrix: helicity decay alpha_ls-couplings missing from BRANCHING for (l,s) = (" +
                                            std::to_string(l) + "," + std::to_string(s) + ")");
              ++nonzero;
            } else {
              auto it = alpha.find({l, s});
              const double ar = it == alpha.end() ? 1.0 : it->second.first, ai = it == alpha.end() ? 0.0 : it->second.second;
              o.T_re[i * n2 + j] += ar * NORM * cg1 * cg2;
              o.T_im[i * n2 + j] += ai * NORM * cg1 * cg2;
            }
          }
    if (MODE == 0) {
      if (nonzero == 0) throw std::invalid_argument("gra::spin::InitTMatrix: Decay is impossible (spin-parity-statistics) for resonance " + name);
      double sum = 0;
      for (auto &kv : active) {
        auto it = alpha.find(kv.first);
        const double ar = it == alpha.end() ? 1.0 : it->second.first, ai = it == alpha.end() ? 0.0 : it->second.second;
        sum += ar * ar + ai * ai;
      }
      if (std::abs(sum - 1.0) > 1e-6)
        for (auto &kv : active) {
          auto it = alpha.find(kv.first);
          std::pair<double, double> v = it == alpha.end() ? std::make_pair(1.0, 0.0) : it->second;
          alpha[kv.first] = {v.first / std::sqrt(sum), v.second / std::sqrt(sum)};
        }
    }
  }
  // equivalent 2-body decay coupling (reference src/MProcess.cc:680-727, PDW2body MKinematics.h:331-334)
  auto kallen = [](double x, double y, double z) { return std::sqrt(std::max(0.0, x * x + y * y + z * z - 2.0 * (x * y + x * z + y * z))); };
  auto pdw2 = [&](double M2) {
    const double m = std::sqrt(std::max(0.0, M2));
    return kallen(M2, d.nodes[0].mass * d.nodes[0].mass, d.nodes[1].mass * d.nodes[1].mass) * 1.0 / (16.0 * 3.141592653589793238462643383279502884 * 1.0 * (m * (m * m)));
  };
  double PS = pdw2(o.mass * o.mass);
  if (std::abs(PS) < 1e-12)
    for (int i = 1; i <= 3; ++i) {
      PS = pdw2((o.mass + i * o.width) * (o.mass + i * o.width));
      if (PS > 1e-12) break;
    }
  // tensor-pomeron decay coupling (src/MProcess.cc:683-700, MTensorPomeron::GDecay src/MTensorPomeron.cc:138-156): from
  // BRANCHING.json when given there, else from J, M, W, BR and the first daughter's mass -- with the same mother-mass
  // shifts M + i W (i = 1..3) the phase-space loop above takes for a purely off-shell decay
  if (gdt.empty()) {
    auto gdecay = [&](double M) {
      const double pw = o.width * o.BR, PI_ = 3.141592653589793238462643383279502884;
      const double r = d.nodes[0].mass / M, P = std::sqrt(1 - 4 * (r * r));
      const int Jr = o.spinX2 / 2;
      if (Jr == 0) return std::sqrt(pw / (1.0 / (16 * PI_ * M) * (1.0 * 1.0) * P));
      if (Jr == 1) return std::sqrt(pw / (M / (192 * PI_) * (P * (P * P))));
      if (Jr == 2) return std::sqrt(pw / (M / (480 * PI_) * ((M / 1.0) * (M / 1.0)) * (P * ((P * P) * (P * P)))));
      if (Jr == 4) return 1.0;
      throw std::invalid_argument("MTensorPomeron::GDecay: Unknown input spin J = " + std::to_string(Jr));
    };
    gdt = {gdecay(o.mass)};
    if (std::abs(pdw2(o.mass * o.mass)) < 1e-12)
      for (int i = 1; i <= 3; ++i) {
        const double PSi = pdw2((o.mass + i * o.width) * (o.mass + i * o.width));
        gdt = {gdecay(o.mass + i * o.width)};
        if (PSi > 1e-12) break;
      }
  }
  for (size_t i = 0; i < gdt.size() && i < 2; ++i) o.g_decay_tensor[i] = gdt[i];
  o.g_decay = std::sqrt(std::max(0.0, o.BR * o.width / PS));
  if (std::isnan(o.g_decay) || std::isinf(o.g_decay))
    throw std::invalid_argument("MProcess::ProcessHelicityDecay: Kinematic coupling problem with resonance " + name + " (daughters too heavy?)");
}

// ---- the flattening proper ---------------------------------------------------------------------------------
static void parse_card_impl(const std::string &card_path, const char *process_override, int pomloop_override,
                            const char *base, grb_process_desc &d) {
  memset(&d, 0, sizeof(d));
  d.abi_version = GRB_ABI_VERSION;
  const Json card = parse_json5(read_file(card_path));

  // GENERALPARAM (reference MGraniitti::ReadGeneralParam, src/MGraniitti.cc:332-358)
  const Json &      gen        = card.at("GENERALPARAM");
  const std::string modelparam = gen.at("MODELPARAM").as_str();

  // model tune + particle table
  ParticleTable pdg;
  Json          tune;
  if (base && base[0]) {
    pdg.load_mcd(std::string(base) + "/modeldata/mass_width_2018.mcd");
    tune = parse_json5(read_file(std::string(base) + "/modeldata/" + modelparam + "/GENERAL.json"));
  } else {
    if (modelparam != "TUNE0") throw Unsupported("built-in model data holds TUNE0 only; pass the GRANIITTI base directory for MODELPARAM=" + modelparam);
    pdg.load_builtin();
    tune = parse_json5(kBuiltinTune0);
  }

  // PROCESSPARAM (reference MGraniitti::ReadProcessParam, src/MGraniitti.cc:375-609)
  const Json &pp   = card.at("PROCESSPARAM");
  std::string full = (process_override && process_override[0] && std::string(process_override) != "null")
                         ? std::string(process_override)
                         : pp.at("PROCESS").as_str();
  std::vector<OneCmd> syntax;
  const size_t        at = full.find('@');
  if (at != std::string::npos) {
    syntax = split_commands(full.substr(at));
    full   = full.substr(0, at);
  }
  std::string  process_part, decay_part;
  const size_t p1 = full.find("->"), p2 = full.find("&>");
  const size_t pa = (p1 != std::string::npos) ? p1 : p2;
  if (pa != std::string::npos) {
    process_part = trim(full.substr(0, pa));
    decay_part   = trim(full.substr(pa + 2));
  } else {
    process_part = trim(full);
  }
  d.ps_active = (p1 == std::string::npos && p2 != std::string::npos) ? 0 : 1;  // SetISOLATE, MProcess.h:100

  // @PDG[id]{M:..,W:..} overrides (reference MProcess::SetProcess, src/MProcess.cc:1740-1790)
  for (auto &c : syntax)
    if (c.id == "PDG") {
      if (c.target.size() != 1) throw std::invalid_argument("@Syntax error: invalid PDG[] target");
      const int id = std::stoi(c.target[0]);
      pdg.by_pdg(id);
      for (auto &kv : c.arg) {
        if (kv.first == "M") { pdg.tab[id].mass = std::stod(kv.second); if (pdg.tab.count(-id)) pdg.tab[-id].mass = pdg.tab[id].mass; }
        if (kv.first == "W") { pdg.tab[id].width = std::stod(kv.second); if (pdg.tab.count(-id)) pdg.tab[-id].width = pdg.tab[id].width; }
      }
    }

  // "PP[CON]<F>" -> istate, channel, class (reference MProcess::ParseCMD, src/MProcess.cc:1817-1850)
  std::string ps = process_part;
  ps.erase(std::remove(ps.begin(), ps.end(), ' '), ps.end());
  const size_t lb = ps.find('['), rb = ps.find(']'), la = ps.find('<'), ra = ps.find('>');
  if (lb == std::string::npos || rb == std::string::npos || la == std::string::npos || ra == std::string::npos)
    throw std::invalid_argument("MProcess::SetProcess: Unknown PROCESS: " + ps);
  const std::string istate = ps.substr(0, lb), channel = ps.substr(lb + 1, rb - lb - 1), cls = ps.substr(la + 1, ra - la - 1);
  if (cls == "F") d.ps_class = GRB_CLASS_F;
  else if (cls == "C") d.ps_class = GRB_CLASS_C;
  else throw Unsupported("phase-space class <" + cls + "> is outside the hot-path scope");
  const std::string key = istate + "[" + channel + "]";
  static const std::map<std::string, int> registry = {
      {"yy[CON]", GRB_PROC_YY_CON}, {"yy[QED]", GRB_PROC_YY_QED}, {"PP[RES]", GRB_PROC_PP_RES},
      {"PP[CONTENSOR]", GRB_PROC_PP_CONTENSOR}, {"PP[CON]", GRB_PROC_PP_CON}, {"PP[CON-]", GRB_PROC_PP_CON_MINUS},
      {"PP[RES+CON]", GRB_PROC_PP_RES_CON}, {"PP[RESTENSOR]", GRB_PROC_PP_RESTENSOR}, {"PP[RES+CONTENSOR]", GRB_PROC_PP_RES_CONTENSOR}};
  auto it = registry.find(key);
  if (it == registry.end()) throw Unsupported("MSubProc: process " + key + "<" + cls + "> is outside the hot-path scope");
  d.proc = it->second;

  d.screening  = (pomloop_override >= 0) ? (pomloop_override != 0) : pp.at("POMLOOP").as_bool();
  d.excitation = pp.at("NSTARS").as_int();
  if (d.excitation < 0 || d.excitation > 2) throw std::invalid_argument("MProcess::SetExcitation: Not valid input (0,1,2)");

  // decay mode (reference MProcess::SetDecayMode, src/MProcess.cc:380-437)
  {
    int L = 0, R = 0, arrows = 0;
    for (char c : decay_part) { L += c == '{'; R += c == '}'; arrows += c == '>'; }
    if (L != R) throw std::invalid_argument("MProces::SetDecayMode: ERROR: Decay tree has unbalanced { } brackets");
    if (arrows != L) throw std::invalid_argument("MProces::SetDecayMode: ERROR: Decay tree syntax with arrows > not equal to { } brackets!");
  }
  std::vector<Branch> tree;
  tokenize(decay_part, pdg, tree);
  flatten(tree, d);
  // symmetry factor over identical first-level legs (reference CalculateSymmetryFactor, MProcess.cc:449-468)
  {
    d.S_factor = 1.0;
    std::vector<bool> marked(tree.size(), false);
    for (size_t i = 0; i < tree.size(); ++i) {
      int n = 1;
      marked[i] = true;
      for (size_t j = 0; j < tree.size(); ++j)
        if (i != j && tree[i].p.pdg == tree[j].p.pdg && !marked[j]) { marked[j] = true; ++n; }
      d.S_factor *= factorial(n);
    }
  }

  // command syntax (reference src/MGraniitti.cc:585-596)
  d.offshell = 5.0;  // MProcess.h OFFSHELL default
  for (auto &c : syntax) {
    if (c.id == "FLATAMP") d.flatamp = std::stoi(c.arg["_SINGLET_"]);
    else if (c.id == "FLATMASS2") d.flatmass2 = c.arg["_SINGLET_"] == "true";
    else if (c.id == "OFFSHELL") d.offshell = std::stod(c.arg["_SINGLET_"]);
    else if (c.id == "PDG" || c.id == "RES" || c.id == "R" || c.id == "SPINGEN" || c.id == "SPINDEC" || c.id == "FRAME" || c.id == "JMAX") {}
    else throw std::invalid_argument("unknown @ command: " + c.id);
  }

  // beams (reference MProcess::SetInitialState/SetBeamEnergies, src/MProcess.cc:291-334)
  {
    const Json &beam = pp.at("BEAM"), &energy = pp.at("ENERGY");
    if (beam.arr.size() != 2 || energy.arr.size() != 2) throw std::invalid_argument("MProcess::SetInitialState: Input BEAM/ENERGY vector not dim 2!");
    const Particle &b1 = pdg.by_name(beam.arr[0].as_str()), &b2 = pdg.by_name(beam.arr[1].as_str());
    const double E1 = std::max(energy.arr[0].as_num(), 1.001 * b1.mass), E2 = std::max(energy.arr[1].as_num(), 1.001 * b2.mass);
    auto msq = [](double x) { return std::sqrt(std::max(0.0, x)); };
    d.beam1_mass = b1.mass; d.beam2_mass = b2.mass;
    d.pbeam1_pz = msq(E1 * E1 - b1.mass * b1.mass); d.pbeam1_E = E1;
    d.pbeam2_pz = -msq(E2 * E2 - b2.mass * b2.mass); d.pbeam2_E = E2;
    // s = (pbeam1 + pbeam2).M2(), reference M4Vec.h:123,139
    const double sz = d.pbeam1_pz + d.pbeam2_pz, se = d.pbeam1_E + d.pbeam2_E;
    d.s      = se * se - (0.0 * 0.0 + 0.0 * 0.0 + sz * sz);
    d.sqrt_s = msq(d.s);
    if (d.sqrt_s < b1.mass + b2.mass) throw std::invalid_argument("MProcess::SetBeamEnergies: Error with input CMS energy");
  }

  // INTEGRALPARAM: loop discretisation (reference src/MGraniitti.cc:610-628, MEikonal.h:44-51)
  {
    const int ND = card.at("INTEGRALPARAM").at("POMLOOP").at("ND").as_int();
    if (ND < -10 || ND > 10) throw std::invalid_argument("POMLOOP::ND out of range [-10, 10]");
    d.loop_nkt    = std::max(3, 3 * ND + 15);
    d.loop_nphi   = std::max(3, 3 * ND + 12);
    d.loop_min_kt = 1E-4;
    d.loop_max_kt = 1.75;
  }

  // GENCUTS (reference MGraniitti::ReadGenCuts, src/MGraniitti.cc:673-823) + SetTechnicalBoundaries
  {
    grb_gencut &g = d.gcuts;
    g.forward_pt_min = -1.0; g.forward_pt_max = -1.0; g.XI_min = 0.0; g.XI_max = 1.0;  // gra::GENCUT defaults, MKinematics.h:1554-1576
    const Json &gc = card.at("GENCUTS");
    g.kt_min = 0.0; g.kt_max = -1.0; g.rap_min = 0.0; g.rap_max = 0.0;
    if (d.ps_class == GRB_CLASS_C) {  // reference src/MGraniitti.cc:688-745 + MContinuum::post_Constructor (src/MContinuum.cc:64-91)
      if (!gc.has("<C>") || !gc.at("<C>").has("Rap"))
        throw std::invalid_argument("MGraniitti::ReadGenCuts: <C> phase space class requires from user: \"GENCUTS\" : { \"<C>\" : { \"Rap\" : [min, max] }}");
      const Json &Cc = gc.at("<C>");
      range2(Cc.at("Rap"), "GENCUTS::<C>::Rap", g.rap_min, g.rap_max);
      if (Cc.has("Kt")) range2(Cc.at("Kt"), "GENCUTS::<C>::Kt", g.kt_min, g.kt_max);
      if (Cc.has("Pt")) range2(Cc.at("Pt"), "GENCUTS::<C>::Pt", g.forward_pt_min, g.forward_pt_max);
      if (Cc.has("Xi")) range2(Cc.at("Xi"), "GENCUTS::<C>::Xi", g.XI_min, g.XI_max);
      g.Y_min = g.Y_max = g.M_min = g.M_max = 0.0;
      if (g.forward_pt_min < 0.0) g.forward_pt_min = 0.0;
      if (g.forward_pt_max < 0.0) g.forward_pt_max = d.excitation == 0 ? 2.5 : (d.excitation == 1 ? 50.0 : 100.0);
      if (g.kt_max < 0.0) {
        g.kt_min = 0.0;
        g.kt_max = 10.0 / d.n_legs;
        if (istate != "PP" && istate != "yP" && istate != "OP") g.kt_max = d.sqrt_s / 2.0;
      }
    } else {
    if (!gc.has("<F>") || !gc.at("<F>").has("Rap") || !gc.at("<F>").has("M"))
      throw std::invalid_argument("MGraniitti::ReadGenCuts: <F> phase space class requires from user: \"GENCUTS\" : { \"<F>\" : { \"Rap\" : [min, max], \"M\" : [min, max] }}");
    const Json &F = gc.at("<F>");
    range2(F.at("Rap"), "GENCUTS::<F>::Rap", g.Y_min, g.Y_max);
    range2(F.at("M"), "GENCUTS::<F>::M", g.M_min, g.M_max);
    if (F.has("Pt")) range2(F.at("Pt"), "GENCUTS::<F>::Pt", g.forward_pt_min, g.forward_pt_max);
    if (F.has("Xi")) range2(F.at("Xi"), "GENCUTS::<F>::Xi", g.XI_min, g.XI_max);
    if (g.forward_pt_min < 0.0) g.forward_pt_min = 0.0;                                       // MProcess.cc:1369-1386
    if (g.forward_pt_max < 0.0) g.forward_pt_max = d.excitation == 0 ? 2.5 : (d.excitation == 1 ? 50.0 : 100.0);
    }
  }
  d.lipsdim = d.ps_class == GRB_CLASS_C ? 3 * (d.n_legs + 2) - 4 + d.excitation  // MContinuum.cc:86-90
                                        : 5 + 1 + d.excitation;                 // MFactorized.cc:69-78

  // FIDCUTS (reference MGraniitti::ReadFidCuts, src/MGraniitti.cc:826-907); defaults of gra::FIDCUT
  {
    grb_fidcut &f = d.fcuts;
    f.active = 0; f.usercuts = 0;
    if (card.has("FIDCUTS") && card.at("FIDCUTS").has("active")) {
      const Json &fc = card.at("FIDCUTS");
      f.active = fc.at("active").as_bool();
      const Json &P = fc.at("PARTICLE"), &S = fc.at("SYSTEM"), &W = fc.at("FORWARD");
      range2(P.at("Eta"), "FIDCUTS::PARTICLE::Eta", f.eta_min, f.eta_max);
      range2(P.at("Rap"), "FIDCUTS::PARTICLE::Rap", f.rap_min, f.rap_max);
      range2(P.at("Pt"), "FIDCUTS::PARTICLE::Pt", f.pt_min, f.pt_max);
      range2(P.at("Et"), "FIDCUTS::PARTICLE::Et", f.Et_min, f.Et_max);
      range2(S.at("M"), "FIDCUTS::SYSTEM::M", f.M_min, f.M_max);
      range2(S.at("Rap"), "FIDCUTS::SYSTEM::Rap", f.Y_min, f.Y_max);
      range2(S.at("Pt"), "FIDCUTS::SYSTEM::Pt", f.Pt_min, f.Pt_max);
      range2(W.at("t"), "FIDCUTS::FORWARD::t", f.forward_t_min, f.forward_t_max);
      range2(W.at("M"), "FIDCUTS::FORWARD::M", f.forward_M_min, f.forward_M_max);
      const Json &uc = fc.at("USERCUTS");
      f.usercuts = (uc.type == Json::BOOL) ? (uc.b ? 1 : 0) : uc.as_int();
    }
  }

  // VETOCUTS (reference MGraniitti::ReadVetoCuts, src/MGraniitti.cc:910-963)
  if (card.has("VETOCUTS") && card.at("VETOCUTS").has("active")) {
    const Json &vc = card.at("VETOCUTS");
    d.veto_active  = vc.at("active").as_bool();
    for (int i = 0; i < 100; ++i) {
      const std::string k = std::to_string(i);
      if (!vc.has(k) || !vc.at(k).has("Eta") || !vc.at(k).has("Pt")) break;
      if (d.n_veto >= GRB_MAX_VETO) throw Unsupported("more veto domains than GRB_MAX_VETO");
      grb_vetodomain &v = d.veto[d.n_veto++];
      range2(vc.at(k).at("Eta"), "VETOCUTS::Eta", v.eta_min, v.eta_max);
      range2(vc.at(k).at("Pt"), "VETOCUTS::Pt", v.pt_min, v.pt_max);
    }
  }

  // model parameters: PARAM_SOFT (src/MForm.cc:110-135), PARAM_FLAT, PARAM_STRUCTURE
  {
    const Json &S = tune.at("PARAM_SOFT");
    grb_param_soft &o = d.soft;
    o.DELTA_P = S.at("DELTA_P").as_num(); o.ALPHA_P = S.at("ALPHA_P").as_num();
    o.gN_P = S.at("gN_P").as_num(); o.gN_O = S.at("gN_O").as_num();
    o.g3P = S.at("g3P").as_num() * o.gN_P;  // convention: file holds g3P/gN_P (MForm.cc:122)
    o.gamma = S.at("gamma").as_num();
    o.fc1 = S.at("fc1").as_num(); o.fc2 = S.at("fc2").as_num(); o.fc3 = S.at("fc3").as_num();
    o.odderon_on = S.at("ODDERON_ON").as_bool();
    d.flat_b = tune.at("PARAM_FLAT").at("b").as_num();
    const Json &T = tune.at("PARAM_STRUCTURE");
    const std::string em = T.at("EM").as_str(), qa = T.at("QED_alpha").as_str();
    d.em_ff     = (em == "KELLY") ? 1 : 0;
    d.soft.f2_model = (T.has("F2") && T.at("F2").as_str() == "DL") ? 1 : 0;   // PARAM_STRUCTURE.F2: CKMT (default) or DL
    d.qed_alpha = (qa == "ZERO") ? 0 : 1;
  }
  // PARAM_REGGE + per-particle couplings chosen by the first leg (src/MRegge.cc:984-1033)
  {
    const Json &R = tune.at("PARAM_REGGE");
    grb_param_regge &o = d.regge;
    auto a0 = R.at("a0").as_vec(), ap = R.at("ap").as_vec(), sg = R.at("sgn").as_vec();
    if (a0.size() != 3 || ap.size() != 3 || sg.size() != 3) throw std::invalid_argument("REGGE_PARAM::ReadParameters: expected 3 trajectories");
    for (int k = 0; k < 3; ++k) { o.a0[k] = a0[k]; o.ap[k] = ap[k]; o.sgn[k] = sg[k]; }
    o.s0 = R.at("s0").as_num(); o.offshellFF = R.at("offshellFF").as_int(); o.reggeize = R.at("reggeize").as_bool();
    o.b_EXP = R.at("b_EXP").as_num(); o.a_OREAR = R.at("a_OREAR").as_num(); o.b_OREAR = R.at("b_OREAR").as_num();
    o.b_POW = R.at("b_POW").as_num(); o.omega = R.at("omega").as_num();
    const int apdg = d.n_legs > 0 ? std::abs(d.nodes[0].pdg) : 0;
    const char *blk = "PARAM_X";
    if (apdg == 211 || apdg == 111) blk = "PARAM_PI";
    else if (apdg == 321 || apdg == 311) blk = "PARAM_K";
    else if (apdg == 2212 || apdg == 2112) blk = "PARAM_P";
    else if (apdg == 113 || apdg == 213) blk = "PARAM_RHO";
    else if (apdg == 331) blk = "PARAM_PHI";  // (sic) the reference keys phi(1020) couplings on 331
    else if (apdg == 9000221) blk = "PARAM_F0500";
    auto c = tune.at(blk).at("c").as_vec();
    const Json &n = tune.at(blk).at("n");
    for (int k = 0; k < 3; ++k) { o.c[k] = c[k]; o.n[k] = n.arr[k].as_bool(); }
  }
  {
    const Json &T = tune.at("PARAM_TENSORPOM");
    grb_param_tensorpom &o = d.tpom;
    o.delta_P = T.at("delta_P").as_num(); o.delta_O = T.at("delta_O").as_num();
    o.delta_1R = T.at("delta_1R").as_num(); o.delta_2R = T.at("delta_2R").as_num();
    o.ap_P = T.at("ap_P").as_num(); o.ap_O = T.at("ap_O").as_num(); o.ap_1R = T.at("ap_1R").as_num(); o.ap_2R = T.at("ap_2R").as_num();
    o.eta_O = T.at("eta_O").as_num(); o.gPNN = T.at("gPNN").as_num(); o.gPpipi = T.at("gPpipi").as_num(); o.gPKK = T.at("gPKK").as_num();
    auto rr = T.at("gPrhorho").as_vec(), ff = T.at("gPphiphi").as_vec();
    o.gPrhorho[0] = rr[0]; o.gPrhorho[1] = rr[1]; o.gPphiphi[0] = ff[0]; o.gPphiphi[1] = ff[1];
    o.M_O = T.at("M_O").as_num(); o.M_1R = T.at("M_1R").as_num();
  }

  // resonances (reference src/MGraniitti.cc:452-582): card "RES" list or @RES{name:true,...}; std::map order
  if (process_part.find("RES") != std::string::npos) {
    std::vector<std::string> names;
    for (auto &x : pp.at("RES").arr) names.push_back(x.as_str());
    for (auto &c : syntax)
      if (c.id == "RES") {
        names.clear();
        for (auto &kv : c.arg)
          if (kv.second == "true" || kv.second == "1") names.push_back(kv.first);
      }
    std::sort(names.begin(), names.end());
    names.erase(std::unique(names.begin(), names.end()), names.end());
    if ((int)names.size() > GRB_MAX_RES) throw Unsupported("more resonances than GRB_MAX_RES");
    Json branching;
    if (base && base[0]) branching = parse_json5(read_file(std::string(base) + "/modeldata/" + modelparam + "/BRANCHING.json"));
    else branching = parse_json5(kBuiltinBranching);
    for (auto &c : syntax)   // @R target checks, src/MGraniitti.cc:484-497
      if (c.id == "R") {
        if (c.target.empty()) throw std::invalid_argument("@Syntax error: invalid R[] without any target []");
        if (c.target.size() > 1) throw std::invalid_argument("@Syntax error: invalid R[] with multiple targets inside []");
        if (std::find(names.begin(), names.end(), c.target[0]) == names.end())
          throw std::invalid_argument("@Syntax error: invalid R[" + c.target[0] + "] not found");
      }
    for (auto &nm : names) {
      Json R;
      if (base && base[0]) {
        R = parse_json5(read_file(std::string(base) + "/modeldata/" + modelparam + "/RES_" + nm + ".json"));
      } else {
        const char *txt = nullptr;
        for (auto &b : kBuiltinRes)
          if (nm == b.name) txt = b.json;
        if (!txt) throw Unsupported("built-in model data has no resonance '" + nm + "'; pass the GRANIITTI base directory");
        R = parse_json5(txt);
      }
      grb_resonance &o = d.res[d.n_res++];
      std::vector<const std::map<std::string, std::string> *> overrides;
      for (auto &c : syntax)
        if (c.id == "R" && c.target[0] == nm) overrides.push_back(&c.arg);
      fill_resonance(R.at("PARAM_RES"), tune, branching, d, nm, overrides, o);
      for (auto &c : syntax) {
        if (c.id == "SPINGEN") o.spingen = c.arg.at("_SINGLET_") == "true";
        if (c.id == "SPINDEC") o.spindec = c.arg.at("_SINGLET_") == "true";
        if (c.id == "FRAME") o.frame = frame_id(c.arg.at("_SINGLET_"));
        if (c.id == "JMAX") o.jmax = std::stoi(c.arg.at("_SINGLET_"));
      }
    }
  }

  // internal uniforms consumed per point when no rejection loop repeats (for sizing tapes)
  {
    int nt = 0;
    auto bw = [&](int i) { return d.nodes[i].width < 1e-40 ? 0 : 1; };
    auto offshell_draws = [&](int i) {
      if (!bw(i)) return 0;
      int k = 1;
      for (int c = 0; c < d.nodes[i].n_children; ++c) k += bw(d.nodes[i].first_child + c);
      return k;
    };
    for (int i = 0; i < d.n_legs; ++i) nt += offshell_draws(i);
    nt += 2;
    for (int i = 0; i < d.n_nodes; ++i)
      if (d.nodes[i].n_children) {
        for (int c = 0; c < d.nodes[i].n_children; ++c) nt += offshell_draws(d.nodes[i].first_child + c);
        nt += 2;
      }
    d.n_tape = nt;
  }
}

}  // namespace grbhost

extern "C" int grb_parse_card(const char *card_path, const char *process_override, int pomloop_override,
                              const char *graniitti_base, grb_process_desc *out, char *errbuf, int errbuf_len) {
  auto seterr = [&](const std::string &m) {
    if (errbuf && errbuf_len > 0) snprintf(errbuf, errbuf_len, "%s", m.c_str());
  };
  if (!card_path || !out) {
    seterr("null argument");
    return GRB_ERR_ARG;
  }
  try {
    grbhost::parse_card_impl(card_path, process_override, pomloop_override, graniitti_base, *out);
  } catch (const grbhost::Unsupported &e) {
    seterr(e.what());
    return GRB_ERR_UNSUPPORTED;
  } catch (const std::invalid_argument &e) {
    seterr(e.what());
    return GRB_ERR_PARSE;
  } catch (const std::exception &e) {
    seterr(e.what());
    return GRB_ERR_PARSE;
  }
  seterr("");
  return GRB_OK;
}

// ---- the reference's eikonal table cache: names and file format (src/MEikonal.cc:60-128, 404-487) ----------

extern "C" int grb_eikonal_cache_names(const grb_process_desc *d, int pdg1, int pdg2, int number_bt, int number_kt2,
                                       char *mbt_name, char *msa_name, int cap) {
  if (!d || cap <= 0) return GRB_ERR_ARG;
  if (number_bt <= 0) number_bt = 10000;
  if (number_kt2 <= 0) number_kt2 = 10000;
  using std::to_string;
  const grb_param_soft &S = d->soft;
  const double GeV2fm = (6.582119514E-25 * 2.99792458E8) * 1E15;
  // PARAM_SOFT::GetHashString (src/MForm.cc:94-100), MEikonalNumerics::GetHashString (MEikonal.h:60-67)
  const std::string soft = to_string(S.DELTA_P) + to_string(S.ALPHA_P) + to_string(S.gN_P) + to_string(S.gN_O) + to_string(S.fc1) +
                           to_string(S.fc2) + to_string(S.fc3) + to_string(S.odderon_on != 0);
  const std::string num = to_string(1E-6) + to_string(25.0) + to_string((unsigned)number_kt2) + to_string(false) + to_string(1E-6) +
                          to_string(10.0 / GeV2fm) + to_string((unsigned)number_bt) + to_string(false) + to_string(1E-9) +
                          to_string(30.0) + to_string(10000u);
  auto djb2 = [](const std::string &s) {  // gra::aux::djb2hash, src/MAux.cc:281-287
    unsigned long h = 5381;
    for (auto c : s) h = (h << 5) + h + c;
    return h;
  };
  char sq[64];
  snprintf(sq, sizeof sq, "%.0f", std::sqrt(d->s > 0 ? d->s : 0.0));
  const std::string p1 = to_string(pdg1), p2 = to_string(pdg2), tail = "_" + p1 + "_" + p2 + "_" + sq + "_";
  const std::string mbt = "MBT" + tail + to_string(djb2(to_string(d->s) + p1 + "_" + p2 + soft + num));
  const std::string msa = "MSA" + tail + to_string(djb2(to_string(d->s) + p1 + p2 + soft + num));
  if (mbt_name) snprintf(mbt_name, cap, "%s", mbt.c_str());
  if (msa_name) snprintf(msa_name, cap, "%s", msa.c_str());
  return ((int)mbt.size() < cap && (int)msa.size() < cap) ? GRB_OK : GRB_ERR_ARG;
}

extern "C" int grb_write_eikonal_file(const char *path, const double *x, const double *re, const double *im, int n) {
  if (!path || !x || !re || !im || n <= 0) return GRB_ERR_ARG;
  FILE *f = fopen(path, "w");
  if (!f) return GRB_ERR_PARSE;
  for (int i = 0; i < n; ++i) fprintf(f, "%.15g,%.15g,%.15g\n", x[i], re[i], im[i]);  // IArray1D::WriteArray
  return fclose(f) == 0 ? GRB_OK : GRB_ERR_PARSE;
}

extern "C" int grb_read_eikonal_file(const char *path, double *x, double *re, double *im, int n) {
  if (!path || !x || !re || !im || n <= 0) return GRB_ERR_ARG;
  FILE *f = fopen(path, "r");
  if (!f) return GRB_ERR_PARSE;
  int got = 0;
  char line[256];
  while (got < n && fgets(line, sizeof line, f)) {
    char *p = line, *q;
    x[got] = strtod(p, &q);  if (q == p || *q != ',') break;  p = q + 1;
    re[got] = strtod(p, &q); if (q == p || *q != ',') break;  p = q + 1;
    im[got] = strtod(p, &q); if (q == p) break;
    ++got;
  }
  fclose(f);
  return got == n ? GRB_OK : GRB_ERR_PARSE;  // IArray1D::ReadArray: "Corrupted file" unless 3 (N+1) fills
}
