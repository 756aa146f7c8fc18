// ko_json.h -- minimal JSON value/reader/writer for the ORACLE (test infrastructure only).
// Kubernetes objects reach the oracle as JSON manifests, the same shape kubectl/informers use.
// Not a general-purpose library: UTF-8 passthrough, \uXXXX only for the BMP, numbers kept as text.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace kojson {

struct Value;
using Object = std::vector<std::pair<std::string, Value>>;  // insertion order kept (spec order matters for overrides)
using Array = std::vector<Value>;

struct Value {
  enum Kind { Null, Bool, Number, String, Arr, Obj } kind = Null;
  bool b = false;
  std::string s;  // String payload, or Number literal text
  std::shared_ptr<Array> a;
  std::shared_ptr<Object> o;

  static Value null() { return Value(); }
  static Value boolean(bool v) { Value x; x.kind = Bool; x.b = v; return x; }
  static Value number(long long v) { Value x; x.kind = Number; x.s = std::to_string(v); return x; }
  static Value str(const std::string& v) { Value x; x.kind = String; x.s = v; return x; }
  static Value array() { Value x; x.kind = Arr; x.a = std::make_shared<Array>(); return x; }
  static Value object() { Value x; x.kind = Obj; x.o = std::make_shared<Object>(); return x; }

  bool is_null() const { return kind == Null; }
  bool is_obj() const { return kind == Obj; }
  bool is_arr() const { return kind == Arr; }
  bool is_str() const { return kind == String; }

  const Value* find(const std::string& k) const {
    if (kind != Obj) return nullptr;
    for (auto& kv : *o) if (kv.first == k) return &kv.second;
    return nullptr;
  }
  // get(): missing key / wrong kind => shared null
  const Value& get(const std::string& k) const {
    static const Value kNull;
    const Value* v = find(k);
    return v ? *v : kNull;
  }
  std::string str_or(const std::string& dflt) const { return kind == String ? s : dflt; }
  long long int_or(long long dflt) const {
    if (kind == Number) return std::strtoll(s.c_str(), nullptr, 10);
    return dflt;
  }
  bool bool_or(bool dflt) const { return kind == Bool ? b : dflt; }
  const Array& items() const { static const Array kEmpty; return kind == Arr ? *a : kEmpty; }
  const Object& members() const { static const Object kEmpty; return kind == Obj ? *o : kEmpty; }
  // A Quantity / scalar in a manifest may be a JSON string or a bare number: return its text.
  std::string scalar_text() const { return (kind == String || kind == Number) ? s : std::string(); }

  Value& set(const std::string& k, Value v) {
    for (auto& kv : *o) if (kv.first == k) { kv.second = std::move(v); return kv.second; }
    o->emplace_back(k, std::move(v));
    return o->back().second;
  }
  void push(Value v) { a->push_back(std::move(v)); }
};

class Parser {
 public:
  explicit Parser(const std::string& t) : t_(t) {}
  Value parse() {
    Value v = value();
    ws();
    if (i_ != t_.size()) fail("trailing characters");
    return v;
  }

 private:
  const std::string& t_;
  size_t i_ = 0;
  [[noreturn]] void fail(const char* m) { throw std::runtime_error(std::string("json: ") + m + " at " + std::to_string(i_)); }
  void ws() { while (i_ < t_.size() && (t_[i_] == ' ' || t_[i_] == '\n' || t_[i_] == '\t' || t_[i_] == '\r')) ++i_; }
  bool lit(const char* w) {
    size_t n = std::char_traits<char>::length(w);
    if (t_.compare(i_, n, w) == 0) { i_ += n; return true; }
    return false;
  }
  Value value() {
    ws();
    if (i_ >= t_.size()) fail("unexpected end");
    char c = t_[i_];
    if (c == '{') return object();
    if (c == '[') return array();
    if (c == '"') return Value::str(string());
    if (lit("null")) return Value::null();
    if (lit("true")) return Value::boolean(true);
    if (lit("false")) return Value::boolean(false);
    return number();
  }
  Value number() {
    size_t s = i_;
    while (i_ < t_.size() && (std::isdigit((unsigned char)t_[i_]) || t_[i_] == '-' || t_[i_] == '+' || t_[i_] == '.' || t_[i_] == 'e' || t_[i_] == 'E')) ++i_;
    if (s == i_) fail("bad value");
    Value x; x.kind = Value::Number; x.s = t_.substr(s, i_ - s);
    return x;
  }
  std::string string() {
    ++i_;
    std::string out;
    while (true) {
      if (i_ >= t_.size()) fail("unterminated string");
      char c = t_[i_++];
      if (c == '"') break;
      if (c != '\\') { out.push_back(c); continue; }
      if (i_ >= t_.size()) fail("bad escape");
      char e = t_[i_++];
      switch (e) {
        case 'n': out.push_back('\n'); break;
        case 't': out.push_back('\t'); break;
        case 'r': out.push_back('\r'); break;
        case 'b': out.push_back('\b'); break;
        case 'f': out.push_back('\f'); break;
        case 'u': {
          if (i_ + 4 > t_.size()) fail("bad \\u");
          unsigned cp = (unsigned)std::strtoul(t_.substr(i_, 4).c_str(), nullptr, 16);
          i_ += 4;
          if (cp < 0x80) out.push_back((char)cp);
          else if (cp < 0x800) { out.push_back((char)(0xC0 | (cp >> 6))); out.push_back((char)(0x80 | (cp & 0x3F))); }
          else { out.push_back((char)(0xE0 | (cp >> 12))); out.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out.push_back((char)(0x80 | (cp & 0x3F))); }
          break;
        }
        default: out.push_back(e);
      }
    }
    return out;
  }
  Value array() {
    ++i_;
    Value v = Value::array();
    ws();
    if (i_ < t_.size() && t_[i_] == ']') { ++i_; return v; }
    while (true) {
      v.push(value());
      ws();
      if (i_ >= t_.size()) fail("unterminated array");
      if (t_[i_] == ',') { ++i_; continue; }
      if (t_[i_] == ']') { ++i_; break; }
      fail("expected , or ]");
    }
    return v;
  }
  Value object() {
    ++i_;
    Value v = Value::object();
    ws();
    if (i_ < t_.size() && t_[i_] == '}') { ++i_; return v; }
    while (true) {
      ws();
      if (i_ >= t_.size() || t_[i_] != '"') fail("expected key");
      std::string k = string();
      ws();
      if (i_ >= t_.size() || t_[i_] != ':') fail("expected :");
      ++i_;
      v.set(k, value());
      ws();
      if (i_ >= t_.size()) fail("unterminated object");
      if (t_[i_] == ',') { ++i_; continue; }
      if (t_[i_] == '}') { ++i_; break; }
      fail("expected , or }");
    }
    return v;
  }
};

inline Value parse(const std::string& text) { return Parser(text).parse(); }

inline void dump(const Value& v, std::string& out) {
  switch (v.kind) {
    case Value::Null: out += "null"; break;
    case Value::Bool: out += v.b ? "true" : "false"; break;
    case Value::Number: out += v.s; break;
    case Value::String: {
      out.push_back('"');
      for (unsigned char c : v.s) {
        if (c == '"' || c == '\\') { out.push_back('\\'); out.push_back((char)c); }
        else if (c == '\n') out += "\\n";
        else if (c == '\t') out += "\\t";
        else if (c == '\r') out += "\\r";
        else if (c < 0x20) { char buf[8]; std::snprintf(buf, sizeof buf, "\\u%04x", c); out += buf; }
        else out.push_back((char)c);
      }
      out.push_back('"');
      break;
    }
    case Value::Arr: {
      out.push_back('[');
      bool first = true;
      for (auto& x : *v.a) { if (!first) out.push_back(','); first = false; dump(x, out); }
      out.push_back(']');
      break;
    }
    case Value::Obj: {
      out.push_back('{');
      bool first = true;
      for (auto& kv : *v.o) {
        if (!first) out.push_back(',');
        first = false;
        dump(Value::str(kv.first), out);
        out.push_back(':');
        dump(kv.second, out);
      }
      out.push_back('}');
      break;
    }
  }
}
inline std::string dump(const Value& v) { std::string s; dump(v, s); return s; }

}  // namespace kojson
