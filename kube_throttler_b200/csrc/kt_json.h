// kt_json.h -- the small JSON reader/writer of the host layer (Kubernetes manifests in, results out).
// Numbers keep their source text so that resource quantities written as bare numbers (cpu: 1, cpu: 0.5)
// reach the Quantity parser exactly as written.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace ktjson {

struct Node;
using NodePtr = std::shared_ptr<Node>;

struct Node {
  enum Type { Null, Bool, Num, Str, Arr, Obj } type = Null;
  bool b = false;
  std::string text;  // Str: the decoded string; Num: the literal as written
  std::vector<NodePtr> arr;
  std::vector<std::pair<std::string, NodePtr>> obj;  // insertion order preserved

  static const Node& nil() {
    static const Node n;
    return n;
  }
  bool is(Type t) const { return type == t; }
  const Node& operator[](const char* key) const {
    if (type == Obj)
      for (auto& kv : obj)
        if (kv.first == key) return *kv.second;
    return nil();
  }
  std::string str(const std::string& dflt = "") const { return type == Str ? text : dflt; }
  // strings and numbers both serve as scalar text (resource quantities)
  std::string scalar() const {
    if (type == Str || type == Num) return text;
    throw std::runtime_error("expected a string or a number");
  }
  long long integer(long long dflt = 0) const { return type == Num ? std::strtoll(text.c_str(), nullptr, 10) : dflt; }
  bool boolean(bool dflt = false) const { return type == Bool ? b : dflt; }
};

class Parser {
 public:
  explicit Parser(const char* s) : p_(s) {}
  NodePtr parse() {
    NodePtr v = value();
    ws();
    if (*p_) fail("trailing characters");
    return v;
  }

 private:
  const char* p_;
  int depth_ = 0;  // nesting of the value being parsed: bounded, the parser recurses
  struct Nest {
    Parser& p;
    explicit Nest(Parser& q) : p(q) { if (++p.depth_ > 256) p.fail("nested too deeply"); }
    ~Nest() { --p.depth_; }
  };
  [[noreturn]] void fail(const char* what) { throw std::runtime_error(std::string("json: ") + what); }
  void ws() {
    while (*p_ == ' ' || *p_ == '\t' || *p_ == '\n' || *p_ == '\r') ++p_;
  }
  bool eat(const char* lit) {
    const char* q = p_;
    while (*lit && *q == *lit) { ++q; ++lit; }
    if (*lit) return false;
    p_ = q;
    return true;
  }
  static void utf8(std::string& out, unsigned cp) {
    if (cp < 0x80) out += (char)cp;
    else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
    else if (cp < 0x10000) { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
    else { out += (char)(0xF0 | (cp >> 18)); out += (char)(0x80 | ((cp >> 12) & 0x3F)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
  }
  unsigned hex4() {
    unsigned v = 0;
    for (int i = 0; i < 4; ++i) {
      char c = *p_++;
      v <<= 4;
      if (c >= '0' && c <= '9') v |= c - '0';
      else if (c >= 'a' && c <= 'f') v |= c - 'a' + 10;
      else if (c >= 'A' && c <= 'F') v |= c - 'A' + 10;
      else fail("bad \\u escape");
    }
    return v;
  }
  std::string string() {
    if (*p_ != '"') fail("expected string");
    ++p_;
    std::string out;
    while (*p_ != '"') {
      if (!*p_) fail("unterminated string");
      if (*p_ == '\\') {
        ++p_;
        switch (*p_++) {
          case '"': out += '"'; break;
          case '\\': out += '\\'; break;
          case '/': out += '/'; break;
          case 'b': out += '\b'; break;
          case 'f': out += '\f'; break;
          case 'n': out += '\n'; break;
          case 'r': out += '\r'; break;
          case 't': out += '\t'; break;
          case 'u': {
            unsigned cp = hex4();
            if (cp >= 0xD800 && cp < 0xDC00 && p_[0] == '\\' && p_[1] == 'u') {
              p_ += 2;
              unsigned lo = hex4();
              cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
            }
            utf8(out, cp);
            break;
          }
          default: fail("bad escape");
        }
      } else {
        out += *p_++;
      }
    }
    ++p_;
    return out;
  }
  NodePtr value() {
    Nest nest(*this);
    ws();
    auto n = std::make_shared<Node>();
    if (*p_ == '{') {
      ++p_;
      n->type = Node::Obj;
      ws();
      if (*p_ == '}') { ++p_; return n; }
      while (true) {
        ws();
        std::string k = string();
        ws();
        if (*p_++ != ':') fail("expected ':'");
        n->obj.emplace_back(std::move(k), value());
        ws();
        if (*p_ == ',') { ++p_; continue; }
        if (*p_ == '}') { ++p_; return n; }
        fail("expected ',' or '}'");
      }
    }
    if (*p_ == '[') {
      ++p_;
      n->type = Node::Arr;
      ws();
      if (*p_ == ']') { ++p_; return n; }
      while (true) {
        n->arr.push_back(value());
        ws();
        if (*p_ == ',') { ++p_; continue; }
        if (*p_ == ']') { ++p_; return n; }
        fail("expected ',' or ']'");
      }
    }
    if (*p_ == '"') { n->type = Node::Str; n->text = string(); return n; }
    if (eat("true")) { n->type = Node::Bool; n->b = true; return n; }
    if (eat("false")) { n->type = Node::Bool; return n; }
    if (eat("null")) return n;
    const char* s = p_;
    if (*p_ == '-' || *p_ == '+') ++p_;
    while ((*p_ >= '0' && *p_ <= '9') || *p_ == '.' || *p_ == 'e' || *p_ == 'E' || *p_ == '+' || *p_ == '-') ++p_;
    if (p_ == s) fail("unexpected character");
    n->type = Node::Num;
    n->text.assign(s, p_);
    return n;
  }
};

inline NodePtr parse(const char* s) {
  if (!s) throw std::runtime_error("json: null input");
  return Parser(s).parse();
}

// ---- writer: a string builder with just enough structure -----------------------------------------
class Writer {
 public:
  std::string out;
  Writer& begin_obj() { sep(); out += '{'; first_ = true; return *this; }
  Writer& end_obj() { out += '}'; first_ = false; return *this; }
  Writer& begin_arr() { sep(); out += '['; first_ = true; return *this; }
  Writer& end_arr() { out += ']'; first_ = false; return *this; }
  Writer& key(const std::string& k) { sep(); quote(k); out += ':'; first_ = true; return *this; }
  Writer& str(const std::string& v) { sep(); quote(v); return *this; }
  Writer& num(long long v) { sep(); out += std::to_string(v); return *this; }
  Writer& boolean(bool v) { sep(); out += v ? "true" : "false"; return *this; }
  Writer& raw(const std::string& v) { sep(); out += v; return *this; }

 private:
  bool first_ = true;
  void sep() {
    if (!first_) out += ',';
    first_ = false;
  }
  void quote(const std::string& s) {
    out += '"';
    for (unsigned char c : s) {
      switch (c) {
        case '"': out += "\\\""; break;
        case '\\': out += "\\\\"; break;
        case '\n': out += "\\n"; break;
        case '\r': out += "\\r"; break;
        case '\t': out += "\\t"; break;
        default:
          if (c < 0x20) { char buf[8]; std::snprintf(buf, sizeof buf, "\\u%04x", c); out += buf; }
          else out += (char)c;
      }
    }
    out += '"';
  }
};

}  // namespace ktjson
