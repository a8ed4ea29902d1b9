// TEST INFRASTRUCTURE ONLY -- part of the CPU oracle (see oracle/README.md).
// Nothing under genericarpack.jl_b200/ may include, link or execute this file.
//
// Double-double ("Double64") scalar used by the oracle's Double64 instantiation.
// The error-free transformations follow the ones the reference itself vendors in
// src/arpack-blas-nrm2.jl:51-145 (two_sum, two_hilo_sum, two_prod, add_dddd_dd =
// "Algorithm 6", mul_dddd_dd = "Algorithm 12" of Joldes/Muller/Popescu, "Tight and
// rigourous error bounds for basic building blocks of double-word arithmetic").
// Division and square root are the standard Newton refinements; DoubleFloats.jl 1.2.1
// (the user-side package the reference is run with, test/Manifest.toml) is NOT under
// /root/reference, so Double64 parity is to ~1e-30 relative, not bitwise.
#pragma once
#include <cmath>
#include <cstdint>

namespace oracle {

struct dd {
  double hi, lo;
  dd() : hi(0.0), lo(0.0) {}
  dd(double h) : hi(h), lo(0.0) {}
  dd(int h) : hi((double)h), lo(0.0) {}
  dd(double h, double l) : hi(h), lo(l) {}
  explicit operator double() const { return hi; }
};

static inline void two_sum(double a, double b, double& hi, double& lo) {
  hi = a + b;
  double a1 = hi - b;
  double b1 = hi - a1;
  lo = (a - a1) + (b - b1);
}
static inline void two_hilo_sum(double a, double b, double& s, double& e) {
  s = a + b;
  e = b - (s - a);
}
static inline void two_prod(double a, double b, double& s, double& t) {
  s = a * b;
  t = std::fma(a, b, -s);
}

static inline dd operator+(dd x, dd y) {  // arpack-blas-nrm2.jl:78-88
  double hi, lo, thi, tlo, c;
  two_sum(x.hi, y.hi, hi, lo);
  two_sum(x.lo, y.lo, thi, tlo);
  c = lo + thi;
  two_hilo_sum(hi, c, hi, lo);
  c = tlo + lo;
  two_hilo_sum(hi, c, hi, lo);
  return dd(hi, lo);
}
static inline dd operator-(dd x) { return dd(-x.hi, -x.lo); }
static inline dd operator-(dd x, dd y) { return x + (-y); }
static inline dd operator*(dd x, dd y) {  // arpack-blas-nrm2.jl:91-101
  double hi, lo, t;
  two_prod(x.hi, y.hi, hi, lo);
  t = x.lo * y.lo;
  t = std::fma(x.hi, y.lo, t);
  t = std::fma(x.lo, y.hi, t);
  t = lo + t;
  two_hilo_sum(hi, t, hi, lo);
  return dd(hi, lo);
}
static inline dd operator/(dd x, dd y) {
  // long division with two correction steps
  double q1 = x.hi / y.hi;
  dd r = x - y * dd(q1);
  double q2 = r.hi / y.hi;
  r = r - y * dd(q2);
  double q3 = r.hi / y.hi;
  double s, e;
  two_hilo_sum(q1, q2, s, e);
  dd q = dd(s, e) + dd(q3);
  return q;
}
static inline dd& operator+=(dd& x, dd y) { x = x + y; return x; }
static inline dd& operator-=(dd& x, dd y) { x = x - y; return x; }
static inline dd& operator*=(dd& x, dd y) { x = x * y; return x; }
static inline dd& operator/=(dd& x, dd y) { x = x / y; return x; }

static inline bool operator==(dd x, dd y) { return x.hi == y.hi && x.lo == y.lo; }
static inline bool operator!=(dd x, dd y) { return !(x == y); }
static inline bool operator<(dd x, dd y) { return x.hi < y.hi || (x.hi == y.hi && x.lo < y.lo); }
static inline bool operator>(dd x, dd y) { return y < x; }
static inline bool operator<=(dd x, dd y) { return !(y < x); }
static inline bool operator>=(dd x, dd y) { return !(x < y); }

static inline dd abs(dd x) { return (x.hi < 0.0 || (x.hi == 0.0 && x.lo < 0.0)) ? -x : x; }
static inline dd sqrt(dd x) {
  if (x.hi <= 0.0) return dd(x.hi == 0.0 ? 0.0 : std::nan(""));
  // Newton on the inverse square root, as the reference inlines for its dd norm
  // (arpack-blas-nrm2.jl:196-212).
  double rf = 1.0 / std::sqrt(x.hi);
  dd h(x.hi * 0.5, x.lo * 0.5);
  double s, t;
  two_prod(rf, rf, s, t);
  dd r2(s, t);
  dd radj = dd(0.5) - h * r2;
  radj = radj * dd(rf);
  dd rdd = dd(rf) + radj;
  r2 = rdd * rdd;
  radj = dd(0.5) - h * r2;
  radj = radj * rdd;
  rdd = rdd + radj;
  return rdd * x;
}
static inline bool isnan(dd x) { return std::isnan(x.hi) || std::isnan(x.lo); }
static inline dd copysign(dd x, dd y) {
  bool xneg = (x.hi < 0.0 || (x.hi == 0.0 && std::signbit(x.hi)));
  bool yneg = (y.hi < 0.0 || (y.hi == 0.0 && std::signbit(y.hi)));
  return xneg == yneg ? x : -x;
}
static inline dd max(dd a, dd b) { return a < b ? b : a; }
static inline dd min(dd a, dd b) { return b < a ? b : a; }

}  // namespace oracle
