// json.h — the small JSON value the host layer uses for keys, plans, statistics and the plan cache.
// The reference uses nlohmann::json (json_encoding.h:3); only the subset it needs is provided here:
// objects with sorted keys (so dumps are ordered like nlohmann's), arrays, strings, numbers,
// operator[], get<T>(), push_back, dump(indent), parse.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace rtat {

class Json {
 public:
  enum Kind { Null, Bool, Int, Float, String, Array, Object };

 private:
  Kind kind = Null;
  bool b = false;
  long long i = 0;
  double d = 0;
  std::string s;
  std::vector<Json> a;
  std::map<std::string, Json> o;

 public:
  Json() = default;
  Json(bool v) : kind(Bool), b(v) {}
  Json(int v) : kind(Int), i(v) {}
  Json(long v) : kind(Int), i(v) {}
  Json(long long v) : kind(Int), i(v) {}
  Json(unsigned v) : kind(Int), i(v) {}
  Json(unsigned long v) : kind(Int), i((long long)v) {}
  Json(double v) : kind(Float), d(v) {}
  Json(float v) : kind(Float), d(v) {}
  Json(const char *v) : kind(String), s(v) {}
  Json(const std::string &v) : kind(String), s(v) {}
  template <typename T>
  Json(const std::vector<T> &v) : kind(Array) {
    for (auto &x : v) a.emplace_back(x);
  }
  static Json array() { Json j; j.kind = Array; return j; }
  static Json object() { Json j; j.kind = Object; return j; }

  Kind type() const { return kind; }
  bool is_null() const { return kind == Null; }
  bool is_array() const { return kind == Array; }
  bool is_object() const { return kind == Object; }
  size_t size() const { return kind == Array ? a.size() : kind == Object ? o.size() : 0; }
  bool contains(const std::string &key) const { return kind == Object && o.count(key); }

  Json &operator[](const std::string &key) {
    if (kind == Null) kind = Object;
    if (kind != Object) throw std::runtime_error("json: not an object");
    return o[key];
  }
  const Json &operator[](const std::string &key) const {
    auto it = o.find(key);
    if (kind != Object || it == o.end()) throw std::runtime_error("json: missing key " + key);
    return it->second;
  }
  Json &operator[](const char *key) { return (*this)[std::string(key)]; }
  const Json &operator[](const char *key) const { return (*this)[std::string(key)]; }
  Json &operator[](size_t idx) { return a.at(idx); }
  const Json &operator[](size_t idx) const { return a.at(idx); }
  void push_back(const Json &v) {
    if (kind == Null) kind = Array;
    if (kind != Array) throw std::runtime_error("json: not an array");
    a.push_back(v);
  }
  const std::vector<Json> &items() const { return a; }
  const std::map<std::string, Json> &members() const { return o; }
  std::vector<Json>::const_iterator begin() const { return a.begin(); }
  std::vector<Json>::const_iterator end() const { return a.end(); }

  template <typename T>
  T get() const;

  bool operator==(const Json &r) const {
    if (kind != r.kind) {
      if ((kind == Int || kind == Float) && (r.kind == Int || r.kind == Float)) return number() == r.number();
      return false;
    }
    switch (kind) {
      case Null: return true;
      case Bool: return b == r.b;
      case Int: return i == r.i;
      case Float: return d == r.d;
      case String: return s == r.s;
      case Array: return a == r.a;
      case Object: return o == r.o;
    }
    return false;
  }
  bool operator!=(const Json &r) const { return !(*this == r); }

  std::string dump(int indent = -1) const {
    std::string out;
    write(out, indent, 0);
    return out;
  }

  static Json parse(const std::string &text) {
    size_t pos = 0;
    Json v = parse_value(text, pos);
    skip_ws(text, pos);
    if (pos != text.size()) throw std::runtime_error("json: trailing characters");
    return v;
  }

 private:
  double number() const { return kind == Int ? (double)i : d; }

  static void write_string(std::string &out, const std::string &v) {
    out += '"';
    for (char c : v) {
      switch (c) {
        case '"': out += "\\\""; break;
        case '\\': out += "\\\\"; break;
        case '\n': out += "\\n"; break;
        case '\t': out += "\\t"; break;
        case '\r': out += "\\r"; break;
        default:
          if ((unsigned char)c < 0x20) {
            char buf[8];
            std::snprintf(buf, sizeof buf, "\\u%04x", c);
            out += buf;
          } else {
            out += c;
          }
      }
    }
    out += '"';
  }
  static void write_double(std::string &out, double v) {
    if (!std::isfinite(v)) { out += "null"; return; }
    char buf[40];
    for (int prec = 15; prec <= 17; prec++) {  // shortest text that round-trips
      std::snprintf(buf, sizeof buf, "%.*g", prec, v);
      if (std::strtod(buf, nullptr) == v) break;
    }
    out += buf;
    if (!std::strpbrk(buf, ".eEn")) out += ".0";
  }
  void write(std::string &out, int indent, int depth) const {
    auto newline = [&](int dpt) {
      if (indent >= 0) { out += '\n'; out.append((size_t)indent * dpt, ' '); }
    };
    switch (kind) {
      case Null: out += "null"; break;
      case Bool: out += b ? "true" : "false"; break;
      case Int: out += std::to_string(i); break;
      case Float: write_double(out, d); break;
      case String: write_string(out, s); break;
      case Array: {
        if (a.empty()) { out += "[]"; break; }
        out += '[';
        for (size_t k = 0; k < a.size(); k++) {
          if (k) out += ',';
          newline(depth + 1);
          a[k].write(out, indent, depth + 1);
        }
        newline(depth);
        out += ']';
        break;
      }
      case Object: {
        if (o.empty()) { out += "{}"; break; }
        out += '{';
        bool first = true;
        for (auto &[key, val] : o) {
          if (!first) out += ',';
          first = false;
          newline(depth + 1);
          write_string(out, key);
          out += indent >= 0 ? ": " : ":";
          val.write(out, indent, depth + 1);
        }
        newline(depth);
        out += '}';
        break;
      }
    }
  }

  static void skip_ws(const std::string &t, size_t &p) {
    while (p < t.size() && (t[p] == ' ' || t[p] == '\n' || t[p] == '\t' || t[p] == '\r')) p++;
  }
  static std::string parse_string(const std::string &t, size_t &p) {
    std::string out;
    p++;  // opening quote
    while (p < t.size() && t[p] != '"') {
      if (t[p] == '\\') {
        if (++p >= t.size()) break;
        switch (t[p]) {
          case 'n': out += '\n'; break;
          case 't': out += '\t'; break;
          case 'r': out += '\r'; break;
          case 'b': out += '\b'; break;
          case 'f': out += '\f'; break;
          case 'u': {
            unsigned cp = (unsigned)std::strtoul(t.substr(p + 1, 4).c_str(), nullptr, 16);
            p += 4;
            if (cp < 0x80) out += (char)cp;
            else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
            else { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
            break;
          }
          default: out += t[p];
        }
        p++;
      } else {
        out += t[p++];
      }
    }
    if (p >= t.size()) throw std::runtime_error("json: unterminated string");
    p++;
    return out;
  }
  static Json parse_value(const std::string &t, size_t &p) {
    skip_ws(t, p);
    if (p >= t.size()) throw std::runtime_error("json: unexpected end");
    char c = t[p];
    if (c == '{') {
      Json obj = object();
      p++;
      skip_ws(t, p);
      if (p < t.size() && t[p] == '}') { p++; return obj; }
      while (true) {
        skip_ws(t, p);
        if (p >= t.size() || t[p] != '"') throw std::runtime_error("json: expected key");
        std::string key = parse_string(t, p);
        skip_ws(t, p);
        if (p >= t.size() || t[p] != ':') throw std::runtime_error("json: expected ':'");
        p++;
        obj.o[key] = parse_value(t, p);
        skip_ws(t, p);
        if (p < t.size() && t[p] == ',') { p++; continue; }
        if (p < t.size() && t[p] == '}') { p++; return obj; }
        throw std::runtime_error("json: expected ',' or '}'");
      }
    }
    if (c == '[') {
      Json arr = array();
      p++;
      skip_ws(t, p);
      if (p < t.size() && t[p] == ']') { p++; return arr; }
      while (true) {
        arr.a.push_back(parse_value(t, p));
        skip_ws(t, p);
        if (p < t.size() && t[p] == ',') { p++; continue; }
        if (p < t.size() && t[p] == ']') { p++; return arr; }
        throw std::runtime_error("json: expected ',' or ']'");
      }
    }
    if (c == '"') return Json(parse_string(t, p));
    if (t.compare(p, 4, "true") == 0) { p += 4; return Json(true); }
    if (t.compare(p, 5, "false") == 0) { p += 5; return Json(false); }
    if (t.compare(p, 4, "null") == 0) { p += 4; return Json(); }
    size_t start = p;
    bool is_float = false;
    while (p < t.size() && (std::isdigit((unsigned char)t[p]) || t[p] == '-' || t[p] == '+' || t[p] == '.' || t[p] == 'e' || t[p] == 'E')) {
      if (t[p] == '.' || t[p] == 'e' || t[p] == 'E') is_float = true;
      p++;
    }
    if (start == p) throw std::runtime_error(std::string("json: unexpected character '") + c + "'");
    std::string num = t.substr(start, p - start);
    return is_float ? Json(std::strtod(num.c_str(), nullptr)) : Json(std::strtoll(num.c_str(), nullptr, 10));
  }
};

template <>
inline std::string Json::get<std::string>() const {
  if (kind != String) throw std::runtime_error("json: not a string");
  return s;
}
template <>
inline int Json::get<int>() const {
  if (kind == Int) return (int)i;
  if (kind == Float) return (int)d;
  throw std::runtime_error("json: not a number");
}
template <>
inline long long Json::get<long long>() const {
  if (kind == Int) return i;
  if (kind == Float) return (long long)d;
  throw std::runtime_error("json: not a number");
}
template <>
inline double Json::get<double>() const {
  if (kind == Int) return (double)i;
  if (kind == Float) return d;
  throw std::runtime_error("json: not a number");
}
template <>
inline float Json::get<float>() const { return (float)get<double>(); }
template <>
inline bool Json::get<bool>() const {
  if (kind != Bool) throw std::runtime_error("json: not a bool");
  return b;
}

}  // namespace rtat
