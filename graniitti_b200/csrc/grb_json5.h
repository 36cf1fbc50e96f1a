// grb_json5.h -- minimal JSON reader for GRANIITTI steering cards and model files, which are JSON with
// // and /* */ comments and trailing commas (the reference strips those and feeds nlohmann::json,
// reference src/MAux.cc GetInputData).  Header-only, exceptions stay inside the host layer.
#pragma once
#include <cctype>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace grbhost {

struct Json {
  enum Type { NUL, BOOL, NUM, STR, ARR, OBJ } type = NUL;
  bool                                      b   = false;
  double                                    num = 0.0;
  std::string                               str;
  std::vector<Json>                         arr;
  std::vector<std::pair<std::string, Json>> obj;

  bool has(const std::string &k) const {
    for (auto &kv : obj)
      if (kv.first == k) return true;
    return false;
  }
  const Json &at(const std::string &k) const {
    for (auto &kv : obj)
      if (kv.first == k) return kv.second;
    throw std::invalid_argument("missing key \"" + k + "\"");
  }
  double as_num() const {
    if (type == NUM) return num;
    if (type == BOOL) return b ? 1.0 : 0.0;
    throw std::invalid_argument("expected a number");
  }
  int  as_int() const { return (int)as_num(); }
  bool as_bool() const {
    if (type == BOOL) return b;
    if (type == NUM) return num != 0.0;
    throw std::invalid_argument("expected a boolean");
  }
  const std::string &as_str() const {
    if (type != STR) throw std::invalid_argument("expected a string");
    return str;
  }
  std::vector<double> as_vec() const {
    if (type != ARR) throw std::invalid_argument("expected an array");
    std::vector<double> v;
    for (auto &x : arr) v.push_back(x.as_num());
    return v;
  }
};

class Json5Parser {
 public:
  explicit Json5Parser(const std::string &t) : s(t) {}
  Json parse() {
    Json v = value();
    ws();
    if (p != s.size()) fail("trailing characters");
    return v;
  }

 private:
  const std::string &s;
  size_t             p = 0;

  [[noreturn]] void fail(const std::string &m) const {
    size_t line = 1;
    for (size_t i = 0; i < p && i < s.size(); ++i)
      if (s[i] == '\n') ++line;
    throw std::invalid_argument("JSON parse error at line " + std::to_string(line) + ": " + m);
  }
  void ws() {
    for (;;) {
      while (p < s.size() && isspace((unsigned char)s[p])) ++p;
      if (p + 1 < s.size() && s[p] == '/' && s[p + 1] == '/') {
        while (p < s.size() && s[p] != '\n') ++p;
      } else if (p + 1 < s.size() && s[p] == '/' && s[p + 1] == '*') {
        p += 2;
        while (p + 1 < s.size() && !(s[p] == '*' && s[p + 1] == '/')) ++p;
        p = (p + 2 <= s.size()) ? p + 2 : s.size();
      } else {
        return;
      }
    }
  }
  Json value() {
    ws();
    if (p >= s.size()) fail("unexpected end");
    const char c = s[p];
    if (c == '{') return object();
    if (c == '[') return array();
    if (c == '"') {
      Json v;
      v.type = Json::STR;
      v.str  = string();
      return v;
    }
    if (s.compare(p, 4, "true") == 0) { p += 4; Json v; v.type = Json::BOOL; v.b = true; return v; }
    if (s.compare(p, 5, "false") == 0) { p += 5; Json v; v.type = Json::BOOL; v.b = false; return v; }
    if (s.compare(p, 4, "null") == 0) { p += 4; return Json(); }
    // number
    char *      end = nullptr;
    const double x  = strtod(s.c_str() + p, &end);
    if (end == s.c_str() + p) fail("unexpected character");
    p = end - s.c_str();
    Json v;
    v.type = Json::NUM;
    v.num  = x;
    return v;
  }
  std::string string() {
    ++p;  // opening quote
    std::string o;
    while (p < s.size() && s[p] != '"') {
      if (s[p] == '\\' && p + 1 < s.size()) {
        ++p;
        switch (s[p]) {
          case 'n': o += '\n'; break;
          case 't': o += '\t'; break;
          default: o += s[p];
        }
      } else {
        o += s[p];
      }
      ++p;
    }
    if (p >= s.size()) fail("unterminated string");
    ++p;
    return o;
  }
  Json object() {
    Json v;
    v.type = Json::OBJ;
    ++p;
    for (;;) {
      ws();
      if (p < s.size() && s[p] == '}') { ++p; return v; }
      if (p >= s.size() || s[p] != '"') fail("expected a key");
      std::string k = string();
      ws();
      if (p >= s.size() || s[p] != ':') fail("expected ':'");
      ++p;
      v.obj.emplace_back(k, value());
      ws();
      if (p < s.size() && s[p] == ',') { ++p; continue; }
      if (p < s.size() && s[p] == '}') { ++p; return v; }
      fail("expected ',' or '}'");
    }
  }
  Json array() {
    Json v;
    v.type = Json::ARR;
    ++p;
    for (;;) {
      ws();
      if (p < s.size() && s[p] == ']') { ++p; return v; }
      v.arr.push_back(value());
      ws();
      if (p < s.size() && s[p] == ',') { ++p; continue; }
      if (p < s.size() && s[p] == ']') { ++p; return v; }
      fail("expected ',' or ']'");
    }
  }
};

inline Json parse_json5(const std::string &text) { return Json5Parser(text).parse(); }

}  // namespace grbhost
