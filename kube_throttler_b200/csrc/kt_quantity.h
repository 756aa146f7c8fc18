// kt_quantity.h -- exact decimal resource quantities for the packer (host side).
//
// The reference does all of its arithmetic on k8s.io/apimachinery v0.26.4 resource.Quantity (go.mod:14),
// an arbitrary-precision decimal.  The device computes on int64 columns, so the packer's job is:
//   1. parse a quantity string exactly (suffixes of deploy/crd.yaml:181: Ki..Ei, n u m k M G T P E, e/E exponent),
//      rounding anything finer than 1n away from zero and capping BinarySI values at 2^63-1 as
//      resource.ParseQuantity does;
//   2. choose per resource column a power-of-ten scale at which every value of the column is an integer;
//   3. prove the column cannot overflow int64 when summed.
// A quantity is held as mant * 10^exp with mant not divisible by ten (so equal values are identical).
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>

namespace kt {

using i128 = __int128;

struct Quantity {
  i128 mant = 0;
  int exp = 0;  // value = mant * 10^exp; canonical: mant % 10 != 0, or mant == 0 && exp == 0
  enum Format : uint8_t { DecimalSI, BinarySI, DecimalExponent } format = DecimalSI;

  bool is_zero() const { return mant == 0; }
  void canon() {
    if (mant == 0) { exp = 0; return; }
    while (mant % 10 == 0) { mant /= 10; ++exp; }
  }
};

namespace qdetail {
inline bool mul_overflow(i128 a, i128 b, i128* out) { return __builtin_mul_overflow(a, b, out); }
inline i128 pow10(int k) {
  i128 v = 1;
  while (k-- > 0) v *= 10;
  return v;
}
}  // namespace qdetail

// resource.ParseQuantity.  Throws std::runtime_error with the reference's two error texts.
inline Quantity parse_quantity(const std::string& str) {
  static const char* kFormatWrong = "quantities must match the regular expression '^([+-]?[0-9.]+)([eEinumkKMGTP]*[-+]?[0-9]*)$'";
  static const char* kNumeric = "unable to parse numeric part of quantity";
  static const char* kSuffix = "unable to parse quantity's suffix";
  if (str.empty()) throw std::runtime_error(kFormatWrong);
  const size_t end = str.size();
  size_t pos = 0;
  bool positive = true;
  if (str[0] == '-') { positive = false; ++pos; }
  else if (str[0] == '+') ++pos;
  while (pos < end && str[pos] == '0') ++pos;  // leading zeros
  size_t b = pos;
  while (pos < end && str[pos] >= '0' && str[pos] <= '9') ++pos;
  std::string num = str.substr(b, pos - b), denom, suffix;
  if (pos < end && str[pos] == '.') {
    ++pos;
    b = pos;
    while (pos < end && str[pos] >= '0' && str[pos] <= '9') ++pos;
    denom = str.substr(b, pos - b);
  }
  if (pos < end) {  // suffix letters, then an optional signed exponent
    const size_t s0 = pos;
    while (pos < end && std::string("eEinumkKMGTP").find(str[pos]) != std::string::npos) ++pos;
    if (pos < end && (str[pos] == '-' || str[pos] == '+')) ++pos;
    while (pos < end && str[pos] >= '0' && str[pos] <= '9') ++pos;
    if (pos != end) throw std::runtime_error(kFormatWrong);
    suffix = str.substr(s0);
  }
  // (like parseQuantityString, "." and "+" parse as zero: only a malformed suffix is an error)
  // ---- suffix -> (base, exponent, format) ----
  int base2 = 0, base10 = 0;
  Quantity::Format fmt = Quantity::DecimalSI;
  if (suffix.empty()) {
  } else if (suffix == "Ki") { base2 = 10; fmt = Quantity::BinarySI; }
  else if (suffix == "Mi") { base2 = 20; fmt = Quantity::BinarySI; }
  else if (suffix == "Gi") { base2 = 30; fmt = Quantity::BinarySI; }
  else if (suffix == "Ti") { base2 = 40; fmt = Quantity::BinarySI; }
  else if (suffix == "Pi") { base2 = 50; fmt = Quantity::BinarySI; }
  else if (suffix == "Ei") { base2 = 60; fmt = Quantity::BinarySI; }
  else if (suffix == "n") base10 = -9;
  else if (suffix == "u") base10 = -6;
  else if (suffix == "m") base10 = -3;
  else if (suffix == "k") base10 = 3;
  else if (suffix == "M") base10 = 6;
  else if (suffix == "G") base10 = 9;
  else if (suffix == "T") base10 = 12;
  else if (suffix == "P") base10 = 15;
  else if (suffix == "E") base10 = 18;
  else if ((suffix[0] == 'e' || suffix[0] == 'E') && suffix.size() > 1) {
    // decimal exponent: strconv.ParseInt(suffix[1:], 10, 64)
    size_t i = 1;
    bool neg = false;
    if (suffix[i] == '-') { neg = true; ++i; }
    else if (suffix[i] == '+') ++i;
    if (i >= suffix.size()) throw std::runtime_error(kSuffix);
    long long e = 0;
    for (; i < suffix.size(); ++i) {
      if (suffix[i] < '0' || suffix[i] > '9') throw std::runtime_error(kSuffix);
      e = e * 10 + (suffix[i] - '0');
      if (e > 100000) throw std::runtime_error(kSuffix);
    }
    base10 = (int)(neg ? -e : e);
    fmt = Quantity::DecimalExponent;
  } else {
    throw std::runtime_error(kSuffix);
  }
  // ---- mantissa ----
  const std::string digits = num + denom;
  if (digits.size() > 36) throw std::runtime_error(kNumeric);
  Quantity q;
  q.format = fmt;
  for (char c : digits) q.mant = q.mant * 10 + (c - '0');
  q.exp = -(int)denom.size() + base10;
  if (base2) {
    i128 r;
    if (qdetail::mul_overflow(q.mant, (i128)1 << base2, &r)) throw std::runtime_error(kNumeric);
    q.mant = r;
  }
  if (!positive) q.mant = -q.mant;
  q.canon();
  // anything finer than nano is rounded AWAY from zero (inf.RoundUp): a request for some is never a request for none
  if (q.exp < -9) {
    const int drop = -9 - q.exp;
    i128 m = q.mant < 0 ? -q.mant : q.mant;
    bool inexact = false;
    for (int i = 0; i < drop; ++i) {
      if (m % 10) inexact = true;
      m /= 10;
    }
    if (inexact) ++m;
    q.mant = q.mant < 0 ? -m : m;
    q.exp = -9;
    q.canon();
  }
  // BinarySI values are capped at MaxInt64 ("The max is just a simple cap")
  if (fmt == Quantity::BinarySI && q.mant > 0) {
    const i128 cap = ((i128)1 << 63) - 1;
    if (q.exp > 19) { q.mant = cap; q.exp = 0; q.canon(); }
    else if (q.exp >= 0) {
      i128 v;
      if (qdetail::mul_overflow(q.mant, qdetail::pow10(q.exp), &v) || v > cap) { q.mant = cap; q.exp = 0; q.canon(); }
    }
  }
  return q;
}

// Exact compare / add (resource.Quantity.Cmp / Add): align the exponents in 128-bit arithmetic.
inline int quantity_cmp(const Quantity& a, const Quantity& b) {
  if (a.mant == 0 || b.mant == 0) {
    const i128 x = a.mant, y = b.mant;
    return x < y ? -1 : (x > y ? 1 : 0);
  }
  const int e = a.exp < b.exp ? a.exp : b.exp;
  if (a.exp - e > 36 || b.exp - e > 36) {  // magnitudes too far apart to align: the larger exponent dominates
    if ((a.mant < 0) != (b.mant < 0)) return a.mant < 0 ? -1 : 1;
    const bool a_big = a.exp > b.exp;
    return (a.mant < 0) ? (a_big ? -1 : 1) : (a_big ? 1 : -1);
  }
  i128 x, y;
  if (qdetail::mul_overflow(a.mant, qdetail::pow10(a.exp - e), &x) || qdetail::mul_overflow(b.mant, qdetail::pow10(b.exp - e), &y))
    throw std::runtime_error("quantity compare overflows 128 bits");
  return x < y ? -1 : (x > y ? 1 : 0);
}
inline Quantity quantity_add(const Quantity& a, const Quantity& b) {
  if (a.mant == 0) { Quantity r = b; r.format = a.format; return r; }
  if (b.mant == 0) return a;
  const int e = a.exp < b.exp ? a.exp : b.exp;
  i128 x, y, s;
  if (a.exp - e > 36 || b.exp - e > 36 || qdetail::mul_overflow(a.mant, qdetail::pow10(a.exp - e), &x) ||
      qdetail::mul_overflow(b.mant, qdetail::pow10(b.exp - e), &y) || __builtin_add_overflow(x, y, &s))
    throw std::runtime_error("quantity add overflows 128 bits");
  Quantity r;
  r.mant = s;
  r.exp = e;
  r.format = a.format;
  r.canon();
  return r;
}

// Smallest power-of-ten exponent at which q is an integer (<= 0 means it needs fractional digits).
inline int quantity_min_exp(const Quantity& q) { return q.is_zero() ? 0 : (q.exp < 0 ? q.exp : 0); }

// q as an integer count of 10^scale_exp units.  ok=false when it does not fit 62 bits (the packer reports it).
inline int64_t quantity_at_scale(const Quantity& q, int scale_exp, bool* ok) {
  *ok = true;
  if (q.is_zero()) return 0;
  const int up = q.exp - scale_exp;
  if (up < 0) { *ok = false; return 0; }  // would need a finer column scale
  if (up > 30) { *ok = false; return 0; }
  i128 v;
  if (qdetail::mul_overflow(q.mant, qdetail::pow10(up), &v)) { *ok = false; return 0; }
  const i128 lim = (i128)1 << 62;
  if (v >= lim || v <= -lim) { *ok = false; return 0; }
  return (int64_t)v;
}

// Quantity.ScaledValue(scale) (MilliValue = scale -3, Value = scale 0): the value in units of 10^scale, rounded AWAY from
// zero when it is not a whole number of them (apimachinery amount.go "rounding up"); saturates at the int64 range, where the
// reference's result is unspecified.
inline int64_t quantity_scaled_value(const Quantity& q, int scale) {
  if (q.is_zero()) return 0;
  const int up = q.exp - scale;
  i128 v = q.mant;
  if (up >= 0) {
    for (int i = 0; i < up; ++i) {
      if (v > ((i128)INT64_MAX) || v < -((i128)INT64_MAX)) break;  // already beyond int64: saturates below
      v *= 10;
    }
  } else {
    bool inexact = false;
    for (int i = 0; i < -up && v != 0; ++i) {
      if (v % 10 != 0) inexact = true;
      v /= 10;
    }
    if (inexact || (v == 0 && q.mant != 0)) v += q.mant > 0 ? 1 : -1;
  }
  if (v > (i128)INT64_MAX) return INT64_MAX;
  if (v < (i128)INT64_MIN) return INT64_MIN;
  return (int64_t)v;
}

// Plain decimal spelling of v * 10^scale_exp ("0.5", "1", "536870912"): what the status JSON carries.
inline std::string decimal_string(i128 v, int scale_exp) {
  const bool neg = v < 0;
  if (neg) v = -v;
  std::string digits;
  if (v == 0) digits = "0";
  while (v > 0) { digits.insert(digits.begin(), (char)('0' + (int)(v % 10))); v /= 10; }
  std::string s;
  if (scale_exp >= 0) {
    s = digits == "0" ? "0" : digits + std::string((size_t)scale_exp, '0');
  } else {
    const size_t frac = (size_t)(-scale_exp);
    if (digits.size() <= frac) digits.insert(0, frac - digits.size() + 1, '0');
    s = digits.substr(0, digits.size() - frac);
    std::string f = digits.substr(digits.size() - frac);
    while (!f.empty() && f.back() == '0') f.pop_back();
    if (!f.empty()) s += "." + f;
  }
  return neg && s != "0" ? "-" + s : s;
}
inline std::string decimal_string(const Quantity& q) { return decimal_string(q.mant, q.exp); }

// Canonical Kubernetes spelling (resource.Quantity.String) for the two formats the status fields use:
// DecimalSI picks the largest of n/u/m/""/k/M/G/T/P/E that keeps the mantissa an integer;
// BinarySI uses Ki..Ei when the value is a whole multiple of 1024^k, else falls back to DecimalSI.
inline std::string canonical_string(i128 v, int scale_exp, Quantity::Format fmt) {
  if (v == 0) return "0";
  Quantity q;
  q.mant = v;
  q.exp = scale_exp;
  q.canon();
  if (fmt == Quantity::BinarySI && q.exp >= 0 && q.exp <= 18) {
    const i128 whole = q.mant * qdetail::pow10(q.exp);
    static const char* suf[] = {"", "Ki", "Mi", "Gi", "Ti", "Pi", "Ei"};
    int k = 0;
    i128 m = whole;
    while (k < 6 && m % 1024 == 0) { m /= 1024; ++k; }
    return decimal_string(m, 0) + suf[k];
  }
  // DecimalSI: exponent rounded DOWN to a multiple of three, mantissa scaled to match
  int e3 = q.exp >= 0 ? (q.exp / 3) * 3 : -(((-q.exp) + 2) / 3) * 3;
  if (e3 < -9) e3 = -9;
  if (e3 > 18) e3 = 18;
  i128 m = q.mant * qdetail::pow10(q.exp - e3);
  const char* s = "";
  switch (e3) {
    case -9: s = "n"; break;
    case -6: s = "u"; break;
    case -3: s = "m"; break;
    case 0: s = ""; break;
    case 3: s = "k"; break;
    case 6: s = "M"; break;
    case 9: s = "G"; break;
    case 12: s = "T"; break;
    case 15: s = "P"; break;
    default: s = "E"; break;
  }
  return decimal_string(m, 0) + s;
}

}  // namespace kt
