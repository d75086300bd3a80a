// value.hpp -- the document model of a checkpoint and its three codecs.
//
// The reference serialises `EnergyMC<Any>` with serde into yaml / json / cbor chosen by the file
// extension (src/mc/mod.rs:110-120) and reads it back on --save-as / --resume-from (70-106).  This is
// the C++ host's counterpart: an ordered tree of null / bool / i64 / u64 / f64 / string / array / map,
// written and read in the three formats (externally tagged enums and Option::None = null are the
// caller's business, checkpoint.hpp).  Integers keep 64 unsigned bits (generator states), doubles
// round-trip bit for bit (17 significant digits; .nan / .inf in yaml, NaN / Infinity tokens in json as
// Python's json module writes them).  The yaml reader covers what serde_yaml and PyYAML emit for such
// documents: block maps and sequences, flow collections (possibly wrapped over lines), plain and quoted
// scalars.  Host side only; nothing here is on the hot path.
#pragma once
#include <cerrno>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace sadmc_host {

struct Value {
  enum Kind { Null, Bool, Int, UInt, Float, String, Array, Map } kind = Null;
  bool b = false;
  int64_t i = 0;
  uint64_t u = 0;
  double f = 0.0;
  std::string s;
  std::vector<Value> a;
  std::vector<std::pair<std::string, Value>> m;

  Value() {}
  static Value null() { return Value(); }
  static Value boolean(bool x) {
    Value v;
    v.kind = Bool;
    v.b = x;
    return v;
  }
  static Value integer(int64_t x) {
    Value v;
    if (x >= 0) {
      v.kind = UInt;
      v.u = (uint64_t)x;
    } else {
      v.kind = Int;
      v.i = x;
    }
    return v;
  }
  static Value uinteger(uint64_t x) {
    Value v;
    v.kind = UInt;
    v.u = x;
    return v;
  }
  static Value number(double x) {
    Value v;
    v.kind = Float;
    v.f = x;
    return v;
  }
  static Value optional(double x) { return std::isnan(x) ? null() : number(x); } // Option<f64>, NaN = None in the C ABI
  static Value string(const std::string& x) {
    Value v;
    v.kind = String;
    v.s = x;
    return v;
  }
  static Value array() {
    Value v;
    v.kind = Array;
    return v;
  }
  static Value map() {
    Value v;
    v.kind = Map;
    return v;
  }
  Value& set(const std::string& k, Value v) {
    for (auto& kv : m)
      if (kv.first == k) {
        kv.second = std::move(v);
        return *this;
      }
    m.emplace_back(k, std::move(v));
    return *this;
  }
  Value& push(Value v) {
    a.push_back(std::move(v));
    return *this;
  }
  const Value* find(const std::string& k) const {
    for (auto& kv : m)
      if (kv.first == k) return &kv.second;
    return nullptr;
  }
  const Value& at(const std::string& k) const {
    const Value* p = find(k);
    if (!p) throw std::runtime_error("checkpoint: missing field `" + k + "`");
    return *p;
  }
  bool is_null() const { return kind == Null; }
  double as_f64() const {
    if (kind == Float) return f;
    if (kind == UInt) return (double)u;
    if (kind == Int) return (double)i;
    throw std::runtime_error("checkpoint: expected a number");
  }
  uint64_t as_u64() const {
    if (kind == UInt) return u;
    if (kind == Int && i >= 0) return (uint64_t)i;
    if (kind == Float && f >= 0 && f == std::floor(f)) return (uint64_t)f;
    throw std::runtime_error("checkpoint: expected an unsigned integer");
  }
  int64_t as_i64() const {
    if (kind == Int) return i;
    if (kind == UInt) return (int64_t)u;
    if (kind == Float && f == std::floor(f)) return (int64_t)f;
    throw std::runtime_error("checkpoint: expected an integer");
  }
  bool as_bool() const {
    if (kind == Bool) return b;
    if (kind == UInt) return u != 0;
    throw std::runtime_error("checkpoint: expected a bool");
  }
  const std::string& as_string() const {
    if (kind != String) throw std::runtime_error("checkpoint: expected a string");
    return s;
  }
  // externally tagged enum: {"Tag": body} or the bare string "Tag" (unit variant)
  std::string tag() const {
    if (kind == String) return s;
    if (kind == Map && m.size() == 1) return m[0].first;
    throw std::runtime_error("checkpoint: expected an externally tagged enum");
  }
  const Value& body() const {
    if (kind == Map && m.size() == 1) return m[0].second;
    throw std::runtime_error("checkpoint: enum variant has no body");
  }
};

// ---- number formatting shared by json and yaml ------------------------------------------------
inline std::string fmt_f64(double x, bool yaml) {
  if (std::isnan(x)) return yaml ? ".nan" : "NaN";
  if (std::isinf(x)) return yaml ? (x > 0 ? ".inf" : "-.inf") : (x > 0 ? "Infinity" : "-Infinity");
  char buf[40];
  snprintf(buf, sizeof buf, "%.17g", x);
  std::string s = buf;
  // shortest representation that round-trips
  for (int p = 1; p < 17; p++) {
    char t[40];
    snprintf(t, sizeof t, "%.*g", p, x);
    if (strtod(t, nullptr) == x) {
      s = t;
      break;
    }
  }
  const size_t e = s.find('e');
  if (s.find('.') == std::string::npos) {
    if (e == std::string::npos)
      s += ".0";
    else
      s.insert(e, ".0"); // PyYAML only resolves exponents with a dot as floats
  }
  return s;
}

// ---- json ---------------------------------------------------------------------------------------
inline void json_escape(const std::string& s, std::string& out) {
  out += '"';
  for (unsigned char c : s) {
    if (c == '"' || c == '\\') {
      out += '\\';
      out += (char)c;
    } else if (c == '\n') {
      out += "\\n";
    } else if (c == '\t') {
      out += "\\t";
    } else if (c < 0x20) {
      char b[8];
      snprintf(b, sizeof b, "\\u%04x", c);
      out += b;
    } else {
      out += (char)c;
    }
  }
  out += '"';
}
inline void to_json(const Value& v, std::string& out) {
  switch (v.kind) {
    case Value::Null: out += "null"; break;
    case Value::Bool: out += v.b ? "true" : "false"; break;
    case Value::Int: out += std::to_string(v.i); break;
    case Value::UInt: out += std::to_string(v.u); break;
    case Value::Float: out += fmt_f64(v.f, false); break;
    case Value::String: json_escape(v.s, out); break;
    case Value::Array:
      out += '[';
      for (size_t k = 0; k < v.a.size(); k++) {
        if (k) out += ", ";
        to_json(v.a[k], out);
      }
      out += ']';
      break;
    case Value::Map:
      out += '{';
      for (size_t k = 0; k < v.m.size(); k++) {
        if (k) out += ", ";
        json_escape(v.m[k].first, out);
        out += ": ";
        to_json(v.m[k].second, out);
      }
      out += '}';
      break;
  }
}

inline Value scalar_from_text(const std::string& t, bool yaml); // below

struct JsonReader {
  const std::string& s;
  size_t p = 0;
  explicit JsonReader(const std::string& text) : s(text) {}
  void ws() {
    while (p < s.size() && (s[p] == ' ' || s[p] == '\n' || s[p] == '\t' || s[p] == '\r')) p++;
  }
  [[noreturn]] void fail(const char* what) const { throw std::runtime_error(std::string("json: ") + what + " at byte " + std::to_string(p)); }
  std::string str() {
    std::string out;
    p++; // opening quote
    while (p < s.size() && s[p] != '"') {
      if (s[p] == '\\') {
        p++;
        if (p >= s.size()) fail("bad escape");
        const char c = s[p++];
        if (c == 'n')
          out += '\n';
        else if (c == 't')
          out += '\t';
        else if (c == 'r')
          out += '\r';
        else if (c == 'b')
          out += '\b';
        else if (c == 'f')
          out += '\f';
        else if (c == 'u' || c == 'x' || c == 'U') { // \x and \U: yaml double-quoted scalars
          const size_t nd = c == 'u' ? 4 : (c == 'x' ? 2 : 8);
          const unsigned cp = (unsigned)strtoul(s.substr(p, nd).c_str(), nullptr, 16);
          p += nd;
          if (cp < 0x80)
            out += (char)cp;
          else if (cp < 0x800) {
            out += (char)(0xc0 | (cp >> 6));
            out += (char)(0x80 | (cp & 0x3f));
          } else {
            out += (char)(0xe0 | (cp >> 12));
            out += (char)(0x80 | ((cp >> 6) & 0x3f));
            out += (char)(0x80 | (cp & 0x3f));
          }
        } else
          out += c;
      } else {
        out += s[p++];
      }
    }
    if (p >= s.size()) fail("unterminated string");
    p++;
    return out;
  }
  Value value() {
    ws();
    if (p >= s.size()) fail("unexpected end");
    const char c = s[p];
    if (c == '{') {
      Value v = Value::map();
      p++;
      ws();
      if (s[p] == '}') {
        p++;
        return v;
      }
      for (;;) {
        ws();
        if (s[p] != '"') fail("expected a key");
        std::string k = str();
        ws();
        if (s[p] != ':') fail("expected ':'");
        p++;
        v.m.emplace_back(std::move(k), value());
        ws();
        if (s[p] == ',') {
          p++;
          continue;
        }
        if (s[p] == '}') {
          p++;
          return v;
        }
        fail("expected ',' or '}'");
      }
    }
    if (c == '[') {
      Value v = Value::array();
      p++;
      ws();
      if (s[p] == ']') {
        p++;
        return v;
      }
      for (;;) {
        v.a.push_back(value());
        ws();
        if (s[p] == ',') {
          p++;
          continue;
        }
        if (s[p] == ']') {
          p++;
          return v;
        }
        fail("expected ',' or ']'");
      }
    }
    if (c == '"') return Value::string(str());
    size_t q = p;
    while (q < s.size() && s[q] != ',' && s[q] != '}' && s[q] != ']' && s[q] != ' ' && s[q] != '\n' && s[q] != '\r' && s[q] != '\t') q++;
    const std::string tok = s.substr(p, q - p);
    p = q;
    Value v = scalar_from_text(tok, false);
    if (v.kind == Value::String) fail("bad token");
    return v;
  }
};
inline Value from_json(const std::string& text) {
  JsonReader r(text);
  Value v = r.value();
  r.ws();
  if (r.p != text.size()) r.fail("trailing characters");
  return v;
}

// A bare token -> null / bool / integer / float, else a string.
inline Value scalar_from_text(const std::string& t, bool yaml) {
  if (t == "null" || (yaml && (t == "~" || t.empty() || t == "Null" || t == "NULL"))) return Value::null();
  if (t == "true" || (yaml && (t == "True" || t == "TRUE"))) return Value::boolean(true);
  if (t == "false" || (yaml && (t == "False" || t == "FALSE"))) return Value::boolean(false);
  if (t == "NaN" || t == ".nan" || t == ".NaN" || t == ".NAN") return Value::number(std::nan(""));
  if (t == "Infinity" || t == ".inf" || t == "+.inf" || t == ".Inf") return Value::number(INFINITY);
  if (t == "-Infinity" || t == "-.inf" || t == "-.Inf") return Value::number(-INFINITY);
  if (t.empty()) return Value::string(t);
  // integer?
  size_t k = (t[0] == '-' || t[0] == '+') ? 1 : 0;
  bool digits = k < t.size();
  for (size_t q = k; q < t.size(); q++)
    if (t[q] < '0' || t[q] > '9') digits = false;
  if (digits) {
    errno = 0;
    if (t[0] == '-') {
      const long long x = strtoll(t.c_str(), nullptr, 10);
      if (errno == 0) return Value::integer(x);
    } else {
      const unsigned long long x = strtoull(t.c_str() + (t[0] == '+' ? 1 : 0), nullptr, 10);
      if (errno == 0) return Value::uinteger(x);
    }
    return Value::number(strtod(t.c_str(), nullptr)); // beyond 64 bits
  }
  // float? (json: any strtod-complete token; yaml 1.1 core: needs a digit and, for exponents, PyYAML wants a dot -- be liberal)
  char* end = nullptr;
  const double x = strtod(t.c_str(), &end);
  bool has_digit = false;
  for (char c : t)
    if (c >= '0' && c <= '9') has_digit = true;
  if (has_digit && end && *end == 0 && (t[0] == '-' || t[0] == '+' || t[0] == '.' || (t[0] >= '0' && t[0] <= '9'))) return Value::number(x);
  return Value::string(t);
}

// ---- cbor (RFC 8949 subset: the major types serde_cbor emits for such documents) ------------------
inline void cbor_head(int major, uint64_t n, std::string& out) {
  const unsigned char mt = (unsigned char)(major << 5);
  if (n < 24) {
    out += (char)(mt | n);
  } else if (n < 0x100) {
    out += (char)(mt | 24);
    out += (char)n;
  } else if (n < 0x10000) {
    out += (char)(mt | 25);
    out += (char)(n >> 8);
    out += (char)n;
  } else if (n < 0x100000000ull) {
    out += (char)(mt | 26);
    for (int k = 3; k >= 0; k--) out += (char)(n >> (8 * k));
  } else {
    out += (char)(mt | 27);
    for (int k = 7; k >= 0; k--) out += (char)(n >> (8 * k));
  }
}
inline void to_cbor(const Value& v, std::string& out) {
  switch (v.kind) {
    case Value::Null: out += (char)0xf6; break;
    case Value::Bool: out += (char)(v.b ? 0xf5 : 0xf4); break;
    case Value::UInt: cbor_head(0, v.u, out); break;
    case Value::Int:
      if (v.i >= 0)
        cbor_head(0, (uint64_t)v.i, out);
      else
        cbor_head(1, (uint64_t)(-(v.i + 1)), out);
      break;
    case Value::Float: {
      out += (char)0xfb;
      uint64_t bits;
      memcpy(&bits, &v.f, 8);
      for (int k = 7; k >= 0; k--) out += (char)(bits >> (8 * k));
      break;
    }
    case Value::String:
      cbor_head(3, v.s.size(), out);
      out += v.s;
      break;
    case Value::Array:
      cbor_head(4, v.a.size(), out);
      for (auto& x : v.a) to_cbor(x, out);
      break;
    case Value::Map:
      cbor_head(5, v.m.size(), out);
      for (auto& kv : v.m) {
        cbor_head(3, kv.first.size(), out);
        out += kv.first;
        to_cbor(kv.second, out);
      }
      break;
  }
}
struct CborReader {
  const std::string& s;
  size_t p = 0;
  explicit CborReader(const std::string& b) : s(b) {}
  unsigned char byte() {
    if (p >= s.size()) throw std::runtime_error("cbor: truncated");
    return (unsigned char)s[p++];
  }
  uint64_t be(int n) {
    uint64_t x = 0;
    for (int k = 0; k < n; k++) x = (x << 8) | byte();
    return x;
  }
  static double half(uint16_t h) {
    const int e = (h >> 10) & 0x1f, f = h & 0x3ff;
    double x = e == 0 ? std::ldexp((double)f, -24) : (e == 31 ? (f ? std::nan("") : INFINITY) : std::ldexp((double)(f + 1024), e - 25));
    return (h & 0x8000) ? -x : x;
  }
  Value value() {
    const unsigned char ib = byte();
    const int major = ib >> 5, info = ib & 31;
    if (major == 7) {
      if (info == 20) return Value::boolean(false);
      if (info == 21) return Value::boolean(true);
      if (info == 22 || info == 23) return Value::null();
      if (info == 25) return Value::number(half((uint16_t)be(2)));
      if (info == 26) {
        const uint32_t b = (uint32_t)be(4);
        float x;
        memcpy(&x, &b, 4);
        return Value::number((double)x);
      }
      if (info == 27) {
        const uint64_t b = be(8);
        double x;
        memcpy(&x, &b, 8);
        return Value::number(x);
      }
      throw std::runtime_error("cbor: unsupported simple value");
    }
    uint64_t n;
    if (info < 24)
      n = (uint64_t)info;
    else if (info == 24)
      n = be(1);
    else if (info == 25)
      n = be(2);
    else if (info == 26)
      n = be(4);
    else if (info == 27)
      n = be(8);
    else
      throw std::runtime_error("cbor: indefinite lengths are not supported");
    if (major == 0) return Value::uinteger(n);
    if (major == 1) return Value::integer(-1 - (int64_t)n);
    if (major == 2 || major == 3) {
      if (p + n > s.size()) throw std::runtime_error("cbor: truncated string");
      Value v = Value::string(s.substr(p, n));
      p += n;
      return v;
    }
    if (major == 4) {
      Value v = Value::array();
      for (uint64_t k = 0; k < n; k++) v.a.push_back(value());
      return v;
    }
    if (major == 5) {
      Value v = Value::map();
      for (uint64_t k = 0; k < n; k++) {
        Value key = value();
        v.m.emplace_back(key.kind == Value::String ? key.s : std::to_string(key.as_u64()), value());
      }
      return v;
    }
    throw std::runtime_error("cbor: unsupported major type");
  }
};
inline Value from_cbor(const std::string& bytes) {
  CborReader r(bytes);
  return r.value();
}

// ---- yaml ---------------------------------------------------------------------------------------
// A string may be written plain only if no yaml resolver (1.1 as PyYAML, 1.2 core as serde_yaml) could read it as
// anything but the same string: a word or path of [A-Za-z0-9_./-] starting with a letter or '/', not one of the
// boolean / null words of yaml 1.1.  Everything else is double-quoted.
inline bool yaml_plain_ok(const std::string& s) {
  if (s.empty()) return false;
  const unsigned char c0 = (unsigned char)s[0];
  if (!(isalpha(c0) || c0 == '/' || c0 == '_')) return false;
  for (unsigned char c : s)
    if (!(isalnum(c) || c == '_' || c == '.' || c == '/' || c == '-')) return false;
  std::string low;
  for (unsigned char c : s) low += (char)tolower(c);
  static const char* words[] = {"y", "n", "yes", "no", "on", "off", "true", "false", "null", "nan", "inf", "infinity"};
  for (const char* w : words)
    if (low == w) return false;
  return scalar_from_text(s, true).kind == Value::String;
}
inline void yaml_scalar(const Value& v, std::string& out) {
  switch (v.kind) {
    case Value::Null: out += "null"; break;
    case Value::Bool: out += v.b ? "true" : "false"; break;
    case Value::Int: out += std::to_string(v.i); break;
    case Value::UInt: out += std::to_string(v.u); break;
    case Value::Float: out += fmt_f64(v.f, true); break;
    case Value::String:
      if (yaml_plain_ok(v.s))
        out += v.s;
      else
        json_escape(v.s, out); // a double-quoted yaml scalar
      break;
    default: break;
  }
}
inline bool yaml_is_leaf(const Value& v) { // a collection of scalars only: written in flow style on one line
  if (v.kind == Value::Array) {
    for (auto& x : v.a)
      if (x.kind == Value::Array || x.kind == Value::Map) return false;
    return true;
  }
  if (v.kind == Value::Map) {
    for (auto& kv : v.m)
      if (kv.second.kind == Value::Array || kv.second.kind == Value::Map) return false;
    return true;
  }
  return false;
}
inline void yaml_flow(const Value& v, std::string& out) {
  if (v.kind == Value::Array) {
    out += '[';
    for (size_t k = 0; k < v.a.size(); k++) {
      if (k) out += ", ";
      yaml_flow(v.a[k], out);
    }
    out += ']';
  } else if (v.kind == Value::Map) {
    out += '{';
    for (size_t k = 0; k < v.m.size(); k++) {
      if (k) out += ", ";
      yaml_scalar(Value::string(v.m[k].first), out);
      out += ": ";
      yaml_flow(v.m[k].second, out);
    }
    out += '}';
  } else {
    yaml_scalar(v, out);
  }
}
inline void yaml_block(const Value& v, int indent, std::string& out) {
  const std::string pad((size_t)indent, ' ');
  if (v.kind == Value::Map) {
    for (auto& kv : v.m) {
      out += pad;
      yaml_scalar(Value::string(kv.first), out);
      out += ':';
      const Value& x = kv.second;
      if ((x.kind == Value::Map && !x.m.empty() && !yaml_is_leaf(x)) || (x.kind == Value::Array && !x.a.empty() && !yaml_is_leaf(x))) {
        out += '\n';
        yaml_block(x, x.kind == Value::Array ? indent : indent + 2, out);
      } else {
        out += ' ';
        yaml_flow(x, out);
        out += '\n';
      }
    }
  } else if (v.kind == Value::Array) {
    for (auto& x : v.a) {
      out += pad;
      out += "- ";
      if ((x.kind == Value::Map || x.kind == Value::Array) && !yaml_is_leaf(x)) {
        // nested block under a dash: first line shares the dash
        std::string inner;
        yaml_block(x, indent + 2, inner);
        out += inner.substr((size_t)indent + 2);
      } else {
        yaml_flow(x, out);
        out += '\n';
      }
    }
  } else {
    out += pad;
    yaml_scalar(v, out);
    out += '\n';
  }
}
inline std::string to_yaml(const Value& v) {
  std::string out;
  if ((v.kind == Value::Map && v.m.empty()) || (v.kind == Value::Array && v.a.empty()) || (v.kind != Value::Map && v.kind != Value::Array)) {
    yaml_flow(v, out);
    out += '\n';
  } else {
    yaml_block(v, 0, out);
  }
  return out;
}

// index just past the quoted scalar that starts at s[p] (' with '' escapes, " with backslash escapes); npos if unterminated
inline size_t skip_quoted(const std::string& s, size_t p) {
  const char q = s[p++];
  while (p < s.size()) {
    if (q == '"' && s[p] == '\\') {
      p += 2;
      continue;
    }
    if (s[p] == q) {
      if (q == '\'' && p + 1 < s.size() && s[p + 1] == '\'') {
        p += 2;
        continue;
      }
      return p + 1;
    }
    p++;
  }
  return std::string::npos;
}

// Reader: lines -> (indent, text); block structure by indentation, flow collections by a json-like scanner that
// may continue over the following lines.
struct YamlReader {
  struct Line {
    int indent;
    std::string text;
  };
  std::vector<Line> lines;
  size_t cur = 0;

  explicit YamlReader(const std::string& src) {
    size_t p = 0;
    Scan st; // carried from line to line: an open quoted scalar, the depth of an open flow collection
    while (p <= src.size()) {
      size_t e = src.find('\n', p);
      if (e == std::string::npos) e = src.size();
      std::string l = src.substr(p, e - p);
      p = e + 1;
      if (!l.empty() && l.back() == '\r') l.pop_back();
      int ind = 0;
      while ((size_t)ind < l.size() && l[(size_t)ind] == ' ') ind++;
      const bool continuation = st.open_quote != 0 || st.depth > 0;
      std::string t = strip_comment(l.substr((size_t)ind), st);
      if (!continuation && (t.empty() || t == "---" || t == "...")) continue;
      if (continuation) { // the rest of a flow collection or quoted scalar: folded into the line it began on (a line break is a space)
        if (!t.empty() && !lines.empty()) {
          lines.back().text += ' ';
          lines.back().text += t;
        }
        continue;
      }
      lines.push_back({ind, t});
    }
  }
  struct Scan {
    char open_quote = 0;
    int depth = 0;
  };
  // Cut a trailing comment.  A '#' is a comment only outside quoted scalars; a quote opens a scalar only where a
  // scalar can begin (start of a block value or key, or a token inside a flow collection); '[' / '{' open a flow
  // collection only where a block value begins.  Both can stay open over line ends.
  static std::string strip_comment(const std::string& t, Scan& st) {
    const size_t n = t.size();
    size_t end = n, k = 0;
    if (st.open_quote) { // find the end of the scalar that began on an earlier line
      const char q = st.open_quote;
      bool closed = false;
      while (k < n) {
        if (q == '"' && t[k] == '\\') {
          k += 2;
          continue;
        }
        if (t[k] == q) {
          if (q == '\'' && k + 1 < n && t[k + 1] == '\'') {
            k += 2;
            continue;
          }
          k++;
          closed = true;
          break;
        }
        k++;
      }
      if (!closed) return t;
      st.open_quote = 0;
    }
    bool value_start = st.depth == 0 && k == 0; // a key, a "- " entry or a scalar may begin here
    bool seen_key_colon = false;
    while (k < n) {
      const char c = t[k];
      if (st.depth == 0) {
        if (value_start) {
          if (c == ' ') {
            k++;
            continue;
          }
          if (c == '-' && (k + 1 == n || t[k + 1] == ' ')) { // block sequence entry
            k += 2;
            continue;
          }
          if (c == '"' || c == '\'') {
            const size_t e = skip_quoted(t, k);
            if (e == std::string::npos) {
              st.open_quote = c;
              return t;
            }
            k = e;
            value_start = false;
            continue;
          }
          if (c == '[' || c == '{') {
            st.depth = 1;
            k++;
            value_start = false;
            continue;
          }
          value_start = false; // a plain scalar or key begins: quotes and brackets inside it are text
        }
        if (c == ':' && (k + 1 == n || t[k + 1] == ' ') && !seen_key_colon) {
          seen_key_colon = true;
          value_start = true;
          k++;
          continue;
        }
        if (c == '#' && (k == 0 || t[k - 1] == ' ')) {
          end = k;
          break;
        }
        k++;
      } else { // inside a flow collection
        size_t b = k;
        while (b > 0 && t[b - 1] == ' ') b--;
        const bool token_start = b == 0 || t[b - 1] == '[' || t[b - 1] == '{' || t[b - 1] == ',' || (b < k && t[b - 1] == ':');
        if ((c == '"' || c == '\'') && token_start) {
          const size_t e = skip_quoted(t, k);
          if (e == std::string::npos) {
            st.open_quote = c;
            return t;
          }
          k = e;
          continue;
        }
        if (c == '[' || c == '{') st.depth++;
        if (c == ']' || c == '}') st.depth--;
        if (c == '#' && (k == 0 || t[k - 1] == ' ')) {
          end = k;
          break;
        }
        k++;
      }
    }
    while (end > 0 && (t[end - 1] == ' ' || t[end - 1] == '\t')) end--;
    return t.substr(0, end);
  }
  [[noreturn]] void fail(const std::string& what) const { throw std::runtime_error("yaml: " + what + " near line " + std::to_string(cur + 1)); }

  // position of the ": " / trailing ':' that ends a block-map key, or npos
  static size_t key_colon(const std::string& t) {
    if (t.empty() || t[0] == '[' || t[0] == '{') return std::string::npos;
    size_t k = 0;
    if (t[0] == '"' || t[0] == '\'') { // a quoted key
      k = skip_quoted(t, 0);
      if (k == std::string::npos) return std::string::npos;
      return (k < t.size() && t[k] == ':' && (k + 1 == t.size() || t[k + 1] == ' ')) ? k : std::string::npos;
    }
    for (; k < t.size(); k++)
      if (t[k] == ':' && (k + 1 == t.size() || t[k + 1] == ' ')) return k;
    return std::string::npos;
  }
  static std::string unquote(const std::string& t) {
    if (t.size() >= 2 && t[0] == '"' && t.back() == '"') {
      JsonReader r(t);
      return r.str();
    }
    if (t.size() >= 2 && t[0] == '\'' && t.back() == '\'') {
      std::string out;
      for (size_t k = 1; k + 1 < t.size(); k++) {
        out += t[k];
        if (t[k] == '\'' && t[k + 1] == '\'') k++;
      }
      return out;
    }
    return t;
  }
  static Value scalar(const std::string& t) {
    if (t.size() >= 2 && ((t[0] == '"' && t.back() == '"') || (t[0] == '\'' && t.back() == '\''))) return Value::string(unquote(t));
    return scalar_from_text(t, true);
  }

  // flow collection starting in `text` (first char '[' or '{'), possibly continuing on the following lines
  Value flow(std::string text) {
    auto balanced = [](const std::string& s) {
      int depth = 0;
      for (size_t k = 0; k < s.size();) {
        const char c = s[k];
        size_t b = k; // a quote opens a quoted scalar only at the start of a token
        while (b > 0 && s[b - 1] == ' ') b--;
        const bool token_start = b == 0 || s[b - 1] == '[' || s[b - 1] == '{' || s[b - 1] == ',' || (s[b - 1] == ':' && b < k);
        if ((c == '"' || c == '\'') && token_start) {
          const size_t e = skip_quoted(s, k);
          if (e == std::string::npos) return false; // the scalar continues on the next line
          k = e;
          continue;
        }
        if (c == '[' || c == '{') depth++;
        if (c == ']' || c == '}') depth--;
        k++;
      }
      return depth == 0;
    };
    while (!balanced(text)) {
      if (cur >= lines.size()) fail("unterminated flow collection");
      text += ' ';
      text += lines[cur++].text;
    }
    size_t p = 0;
    Value v = flow_value(text, p);
    return v;
  }
  static void fws(const std::string& s, size_t& p) {
    while (p < s.size() && (s[p] == ' ' || s[p] == '\t')) p++;
  }
  Value flow_value(const std::string& s, size_t& p) {
    fws(s, p);
    if (p >= s.size()) fail("empty flow value");
    if (s[p] == '[') {
      Value v = Value::array();
      p++;
      for (;;) {
        fws(s, p);
        if (p < s.size() && s[p] == ']') {
          p++;
          return v;
        }
        v.a.push_back(flow_value(s, p));
        fws(s, p);
        if (p < s.size() && s[p] == ',') p++;
      }
    }
    if (s[p] == '{') {
      Value v = Value::map();
      p++;
      for (;;) {
        fws(s, p);
        if (p < s.size() && s[p] == '}') {
          p++;
          return v;
        }
        std::string k = flow_token(s, p, true);
        fws(s, p);
        if (p >= s.size() || s[p] != ':') fail("expected ':' in a flow map");
        p++;
        Value x = flow_value(s, p);
        v.m.emplace_back(unquote(k), std::move(x));
        fws(s, p);
        if (p < s.size() && s[p] == ',') p++;
      }
    }
    return scalar(flow_token(s, p, false));
  }
  std::string flow_token(const std::string& s, size_t& p, bool key) {
    fws(s, p);
    const size_t b = p;
    if (p < s.size() && (s[p] == '"' || s[p] == '\'')) {
      const size_t e = skip_quoted(s, p);
      if (e == std::string::npos) fail("unterminated quoted scalar");
      p = e;
      return s.substr(b, p - b);
    }
    while (p < s.size() && s[p] != ',' && s[p] != ']' && s[p] != '}' && !(s[p] == ':' && (key || p + 1 == s.size() || s[p + 1] == ' '))) p++;
    size_t e = p;
    while (e > b && s[e - 1] == ' ') e--;
    return s.substr(b, e - b);
  }

  // value that follows "key:" or "- ": inline text if any, else the nested block below
  Value after(const std::string& rest, int parent_indent, bool in_seq_item) {
    if (!rest.empty()) {
      if (rest[0] == '[' || rest[0] == '{') return flow(rest);
      if (rest[0] == '"' || rest[0] == '\'') { // a quoted scalar folded over several lines
        std::string t = rest;
        while (skip_quoted(t, 0) == std::string::npos && cur < lines.size()) {
          t += ' ';
          t += lines[cur++].text;
        }
        return scalar(t);
      }
      // a plain scalar folded over several lines: after `key: text` deeper lines can only continue the text
      std::string t = rest;
      while (cur < lines.size() && lines[cur].indent > parent_indent) {
        t += ' ';
        t += lines[cur++].text;
      }
      return scalar(t);
    }
    if (cur >= lines.size()) return Value::null();
    const Line& nx = lines[cur];
    const bool seq_here = nx.text.rfind("- ", 0) == 0 || nx.text == "-";
    if (nx.indent > parent_indent || (seq_here && nx.indent == parent_indent && !in_seq_item)) return block(nx.indent);
    return Value::null();
  }
  Value block(int indent) {
    if (cur >= lines.size()) return Value::null();
    const std::string& first = lines[cur].text;
    if (first.rfind("- ", 0) == 0 || first == "-") {
      Value v = Value::array();
      while (cur < lines.size() && lines[cur].indent == indent && (lines[cur].text.rfind("- ", 0) == 0 || lines[cur].text == "-")) {
        std::string rest = lines[cur].text.size() > 2 ? lines[cur].text.substr(2) : "";
        while (!rest.empty() && rest[0] == ' ') rest.erase(0, 1);
        const size_t kc = key_colon(rest);
        const bool nested_seq = rest.rfind("- ", 0) == 0 || rest == "-";
        if (!rest.empty() && (kc != std::string::npos || nested_seq)) {
          // "- key: value" / "- - item" open a map / sequence whose entries sit at indent + 2: rewrite the line in place
          lines[cur].indent = indent + 2;
          lines[cur].text = rest;
          v.a.push_back(block(indent + 2));
        } else {
          cur++;
          v.a.push_back(after(rest, indent, true));
        }
      }
      return v;
    }
    const size_t kc0 = key_colon(first);
    if (kc0 == std::string::npos) {
      // a bare scalar or flow collection as the whole document / block
      std::string t = lines[cur++].text;
      if (t[0] == '[' || t[0] == '{') return flow(t);
      return scalar(t);
    }
    Value v = Value::map();
    while (cur < lines.size() && lines[cur].indent == indent) {
      const std::string t = lines[cur].text;
      const size_t kc = key_colon(t);
      if (kc == std::string::npos) fail("expected `key: value`");
      std::string key = unquote(t.substr(0, kc));
      std::string rest = kc + 1 < t.size() ? t.substr(kc + 1) : "";
      while (!rest.empty() && rest[0] == ' ') rest.erase(0, 1);
      cur++;
      v.m.emplace_back(std::move(key), after(rest, indent, false));
    }
    if (cur < lines.size() && lines[cur].indent > indent) fail("unexpected indentation");
    return v;
  }
};
inline Value from_yaml(const std::string& text) {
  YamlReader r(text);
  if (r.lines.empty()) return Value::null();
  Value v = r.block(r.lines[0].indent);
  if (r.cur != r.lines.size()) r.fail("trailing content");
  return v;
}

// ---- by extension (mc/mod.rs:72-79, 111-119) ---------------------------------------------------------
inline std::string extension_of(const std::string& path) {
  const size_t slash = path.find_last_of('/');
  const size_t dot = path.find_last_of('.');
  if (dot == std::string::npos || (slash != std::string::npos && dot < slash)) return "";
  return path.substr(dot + 1);
}
inline std::string dumps(const Value& v, const std::string& ext) {
  std::string out;
  if (ext == "yaml") return to_yaml(v);
  if (ext == "json") {
    to_json(v, out);
    return out;
  }
  if (ext == "cbor") {
    to_cbor(v, out);
    return out;
  }
  throw std::runtime_error("I don't know how to create file with extension \"" + ext + "\""); // mc/mod.rs:118
}
inline Value loads(const std::string& data, const std::string& ext) {
  if (ext == "yaml") return from_yaml(data);
  if (ext == "json") return from_json(data);
  if (ext == "cbor") return from_cbor(data);
  throw std::runtime_error("I don't know how to read file with extension \"" + ext + "\""); // mc/mod.rs:79,104
}

} // namespace sadmc_host
