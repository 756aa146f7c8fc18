// ko_quantity.h -- ORACLE (test infrastructure, never shipped, never on the product path).
//
// CPU restatement of k8s.io/apimachinery v0.26.4 pkg/api/resource.Quantity as far as the
// kube-throttler hot path uses it.  The dependency is NOT vendored under /root/reference
// (go.mod:14 pins k8s.io/apimachinery v0.26.4), so this follows the published algorithm of
// resource.ParseQuantity / Quantity.Add / Sub / Cmp / IsZero and is anchored on the reference's
// own call sites and tests:
//   call sites: pkg/resourcelist/resourcelist.go:51 (Add), :59 (Sub), :68,103,114,126 (Cmp),
//               pkg/apis/schedule/v1alpha1/resource_amount.go:52 (IsZero), :130,132 (Cmp)
//   pinned by : pkg/resourcelist/resourcelist_test.go:119-421 ({-2..2} algebra),
//               v1alpha1/resource_amount_test.go:27-210, integration milli-cpu sums
//               (test/integration/throttle_test.go:60,94,116,138,189).
// PARITY UNPINNED (no reference test covers them): binary suffixes beyond the README's 512Mi,
// decimal-exponent form, sub-nano rounding, the BinarySI 2^63-1 cap.
//
// Representation: exact integer count of nano-units (10^-9) in __int128.  ParseQuantity never keeps
// more than nano precision (it rounds the magnitude UP to 1n granularity), and Add/Sub/Cmp are exact
// at any scale, so an integer number of nanos is an exact model for |value| < ~1.7e29.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>

namespace ko {

typedef __int128 i128;

struct Quantity {
  i128 nano = 0;
  enum Format : uint8_t { DecimalSI = 0, BinarySI = 1, DecimalExponent = 2 } format = DecimalSI;

  bool IsZero() const { return nano == 0; }
  int Cmp(const Quantity& y) const { return nano < y.nano ? -1 : (nano > y.nano ? 1 : 0); }
  // Quantity.Add: "if q.IsZero() { q.Format = y.Format }" then exact add.
  void Add(const Quantity& y) {
    if (IsZero()) format = y.format;
    nano += y.nano;
  }
  void Sub(const Quantity& y) {
    if (IsZero()) format = y.format;
    nano -= y.nano;
  }
};

struct QuantityError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

namespace detail {
inline i128 pow10_i128(int e) {
  i128 r = 1;
  for (int i = 0; i < e; ++i) r *= 10;
  return r;
}
// Largest magnitude we accept: 10^37 nano keeps every intermediate below the i128 limit (~1.7e38).
inline bool fits(i128 v) {
  static const i128 lim = pow10_i128(37);
  return v < lim && v > -lim;
}
}  // namespace detail

// resource.ParseQuantity.  Grammar (also deploy/crd.yaml:181):
//   <sign>? (digits ('.' digits*)? | '.' digits) <suffix>?
//   suffix = Ki Mi Gi Ti Pi Ei | n u m k M G T P E | (e|E) <sign>? digits
inline Quantity ParseQuantity(const std::string& str) {
  if (str.empty()) throw QuantityError("quantities must match the regular expression (empty)");
  if (str == "0") return Quantity{};
  size_t pos = 0, end = str.size();
  bool positive = true;
  if (str[pos] == '+') ++pos;
  else if (str[pos] == '-') { positive = false; ++pos; }
  // strip leading zeros of the integer part like parseQuantityString does
  size_t num_start = pos;
  while (pos < end && str[pos] >= '0' && str[pos] <= '9') ++pos;
  std::string num = str.substr(num_start, pos - num_start);
  std::string denom;
  if (pos < end && str[pos] == '.') {
    ++pos;
    size_t ds = pos;
    while (pos < end && str[pos] >= '0' && str[pos] <= '9') ++pos;
    denom = str.substr(ds, pos - ds);
  }
  // parseQuantityString turns an empty numerator into "0" and accepts an empty denominator ("1." and "." are both legal,
  // see its "we currently allow 1.G" note): ".", "-" and "+" are all zero.
  std::string suffix = str.substr(pos);

  int base = 10, exponent = 0;
  Quantity::Format format = Quantity::DecimalSI;
  if (suffix.empty()) { exponent = 0; }
  else if (suffix == "Ki") { base = 2; exponent = 10; format = Quantity::BinarySI; }
  else if (suffix == "Mi") { base = 2; exponent = 20; format = Quantity::BinarySI; }
  else if (suffix == "Gi") { base = 2; exponent = 30; format = Quantity::BinarySI; }
  else if (suffix == "Ti") { base = 2; exponent = 40; format = Quantity::BinarySI; }
  else if (suffix == "Pi") { base = 2; exponent = 50; format = Quantity::BinarySI; }
  else if (suffix == "Ei") { base = 2; exponent = 60; format = Quantity::BinarySI; }
  else if (suffix == "n") exponent = -9;
  else if (suffix == "u") exponent = -6;
  else if (suffix == "m") exponent = -3;
  else if (suffix == "k") exponent = 3;
  else if (suffix == "M") exponent = 6;
  else if (suffix == "G") exponent = 9;
  else if (suffix == "T") exponent = 12;
  else if (suffix == "P") exponent = 15;
  else if (suffix == "E") exponent = 18;
  else if (suffix[0] == 'e' || suffix[0] == 'E') {
    // decimal exponent: strconv.ParseInt(suffix[1:], 10, 64)
    size_t p = 1;
    bool eneg = false;
    if (p < suffix.size() && (suffix[p] == '+' || suffix[p] == '-')) { eneg = suffix[p] == '-'; ++p; }
    if (p >= suffix.size()) throw QuantityError("unable to parse quantity's suffix: " + str);
    long e = 0;
    for (; p < suffix.size(); ++p) {
      if (suffix[p] < '0' || suffix[p] > '9') throw QuantityError("unable to parse quantity's suffix: " + str);
      e = e * 10 + (suffix[p] - '0');
      if (e > 100) throw QuantityError("oracle: exponent out of supported range: " + str);
    }
    exponent = eneg ? -(int)e : (int)e;
    format = Quantity::DecimalExponent;
  } else {
    throw QuantityError("unable to parse quantity's suffix: " + str);
  }

  // mantissa digits (num . denom) as one integer M with `fd` fractional digits: value = M * 10^-fd * base^exponent
  std::string digits = num + denom;
  size_t nz = 0;
  while (nz + 1 < digits.size() && digits[nz] == '0') ++nz;
  digits = digits.substr(nz);
  if (digits.size() > 36) throw QuantityError("oracle: mantissa too long: " + str);
  i128 M = 0;
  for (char c : digits) M = M * 10 + (c - '0');
  int fd = (int)denom.size();

  // value in nano = M * 10^(9 - fd) * base^exponent   -> numerator / 10^k with exact rounding-up of the magnitude
  i128 numer = M;
  int p10 = 9 - fd;  // power of ten still to apply
  if (base == 10) p10 += exponent;
  else {
    for (int i = 0; i < exponent; ++i) {
      numer *= 2;
      if (!detail::fits(numer)) throw QuantityError("oracle: magnitude out of supported range: " + str);
    }
  }
  i128 nano;
  if (p10 >= 0) {
    if (p10 > 37) throw QuantityError("oracle: magnitude out of supported range: " + str);
    i128 mul = detail::pow10_i128(p10);
    if (numer != 0 && !detail::fits(numer) ) throw QuantityError("oracle: magnitude out of supported range: " + str);
    // overflow check: numer * mul < 10^37
    if (numer != 0) {
      i128 lim = detail::pow10_i128(37) / mul;
      if (numer >= lim) throw QuantityError("oracle: magnitude out of supported range: " + str);
    }
    nano = numer * mul;
  } else {
    int k = -p10;
    if (k > 37) nano = (numer != 0) ? 1 : 0;  // anything non-zero rounds up to 1n
    else {
      i128 div = detail::pow10_i128(k);
      nano = numer / div;
      if (numer % div != 0) nano += 1;  // inf.RoundUp on the magnitude: "if you want some resources, you should get some"
    }
  }
  // "The max is just a simple cap" -- BinarySI only: min(amount, 2^63-1)
  if (format == Quantity::BinarySI) {
    i128 cap = (i128)INT64_MAX * detail::pow10_i128(9);
    if (nano > cap) nano = cap;
  }
  Quantity q;
  q.nano = positive ? nano : -nano;
  q.format = format;
  return q;
}

// Exact decimal text of a Quantity ("0.5", "536870912", "-2").  NOT Quantity.String()'s canonical
// suffix form -- tests compare values numerically.
inline std::string QuantityDecimalString(const Quantity& q) {
  i128 v = q.nano;
  bool neg = v < 0;
  if (neg) v = -v;
  i128 ip = v / 1000000000, fp = v % 1000000000;
  std::string s;
  if (ip == 0) s = "0";
  while (ip > 0) { s.insert(s.begin(), (char)('0' + (int)(ip % 10))); ip /= 10; }
  if (fp != 0) {
    std::string f(9, '0');
    for (int i = 8; i >= 0; --i) { f[i] = (char)('0' + (int)(fp % 10)); fp /= 10; }
    while (!f.empty() && f.back() == '0') f.pop_back();
    s += "." + f;
  }
  if (neg) s.insert(s.begin(), '-');
  return s;
}

}  // namespace ko
