// Element types of the B200 Lanczos library and their arithmetic.
//
//   F64  double            GenericArpack eigs(Float64, ...)
//   C64  cdbl {re, im}     eigs(ComplexF64, Hermitian(A), ...)   (memory layout of Julia ComplexF64)
//   DD   ddbl {hi, lo}     eigs(Double64, ...)                   (memory layout of DoubleFloats.Double64:
//                                                                  two adjacent Float64, hi first)
//   F32  float             eigs(Float32, Float64, ...): Float32 vectors, Float64 factorisation (mixed
//                           precision, src/interface.jl:26-34); element arithmetic in binary32 like the
//                           reference's sgemv / SparseArrays products, dot-product and norm accumulators
//                           in binary64 (every float product is exact in a double)
// The reference is generic over the element type T and its real type TR
// (/root/reference/src/dsaitr.jl:962-981); here T in {double, cdbl, ddbl, float}, TR = Elem<T>::real_t.
//
// Double-double arithmetic uses FMA-based TwoProd and branch-free TwoSum (the same error-free
// transformations the reference vendors in src/arpack-blas-nrm2.jl:51-145).  All functions are
// __host__ __device__ so the host-side tridiagonal work in Double64 uses the same scalar.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#define GA_HD __host__ __device__ __forceinline__

namespace ga {

typedef long long i64;

struct __align__(16) cdbl { double re, im; };
struct __align__(16) ddbl {
  double hi, lo;
  GA_HD ddbl() {}
  GA_HD ddbl(double h) : hi(h), lo(0.0) {}
  GA_HD ddbl(double h, double l) : hi(h), lo(l) {}
};

// ------------------------------------------------------------------ error-free transformations
GA_HD void two_sum(double a, double b, double& s, double& e) {
  s = a + b;
  double a1 = s - b;
  double b1 = s - a1;
  e = (a - a1) + (b - b1);
}
GA_HD void fast_two_sum(double a, double b, double& s, double& e) {  // |a| >= |b|
  s = a + b;
  e = b - (s - a);
}
GA_HD void two_prod(double a, double b, double& p, double& e) {
  p = a * b;
  e = fma(a, b, -p);
}

// accurate double-word addition (Joldes-Muller-Popescu Alg. 6), error < 3u^2
GA_HD ddbl dd_add(ddbl x, ddbl y) {
  double sh, sl, th, tl, c, vh, vl, w;
  two_sum(x.hi, y.hi, sh, sl);
  two_sum(x.lo, y.lo, th, tl);
  c = sl + th;
  fast_two_sum(sh, c, vh, vl);
  w = tl + vl;
  double zh, zl;
  fast_two_sum(vh, w, zh, zl);
  return ddbl(zh, zl);
}
GA_HD ddbl dd_neg(ddbl x) { return ddbl(-x.hi, -x.lo); }
GA_HD ddbl dd_sub(ddbl x, ddbl y) { return dd_add(x, dd_neg(y)); }
// double-word x double-word (Alg. 12), error <= 5u^2
GA_HD ddbl dd_mul(ddbl x, ddbl y) {
  double ch, cl1, tl;
  two_prod(x.hi, y.hi, ch, cl1);
  tl = x.lo * y.lo;
  tl = fma(x.hi, y.lo, tl);
  tl = fma(x.lo, y.hi, tl);
  tl = cl1 + tl;
  double zh, zl;
  fast_two_sum(ch, tl, zh, zl);
  return ddbl(zh, zl);
}
GA_HD ddbl dd_mul_d(ddbl x, double y) {
  double ch, cl1, cl3;
  two_prod(x.hi, y, ch, cl1);
  cl3 = fma(x.lo, y, cl1);
  double zh, zl;
  fast_two_sum(ch, cl3, zh, zl);
  return ddbl(zh, zl);
}
GA_HD ddbl dd_add_d(ddbl x, double y) {
  double sh, sl, v;
  two_sum(x.hi, y, sh, sl);
  v = x.lo + sl;
  double zh, zl;
  fast_two_sum(sh, v, zh, zl);
  return ddbl(zh, zl);
}
GA_HD ddbl dd_div(ddbl x, ddbl y) {
  double q1 = x.hi / y.hi;
  ddbl r = dd_sub(x, dd_mul_d(y, q1));
  double q2 = r.hi / y.hi;
  r = dd_sub(r, dd_mul_d(y, q2));
  double q3 = r.hi / y.hi;
  double s, e;
  fast_two_sum(q1, q2, s, e);
  return dd_add_d(ddbl(s, e), q3);
}
GA_HD ddbl dd_sqrt(ddbl x) {
  if (!(x.hi > 0.0)) return ddbl(x.hi == 0.0 ? 0.0 : NAN, 0.0);
  // Karp-Markstein: s = x*r + r*(x - (x*r)^2)/2 with r ~ 1/sqrt(x)
  double r = 1.0 / sqrt(x.hi);
  double ax = x.hi * r;
  double ph, pl;
  two_prod(ax, ax, ph, pl);
  ddbl d = dd_sub(x, ddbl(ph, pl));
  double corr = d.hi * (r * 0.5);
  double zh, zl;
  fast_two_sum(ax, corr, zh, zl);
  // one more Newton step in double-double to reach ~2^-104
  ddbl s(zh, zl);
  ddbl s2 = dd_mul(s, s);
  ddbl dl = dd_sub(x, s2);
  double c2 = dl.hi * (r * 0.5);
  return dd_add_d(s, c2);
}

GA_HD ddbl operator+(ddbl a, ddbl b) { return dd_add(a, b); }
GA_HD ddbl operator-(ddbl a, ddbl b) { return dd_sub(a, b); }
GA_HD ddbl operator-(ddbl a) { return dd_neg(a); }
GA_HD ddbl operator*(ddbl a, ddbl b) { return dd_mul(a, b); }
GA_HD ddbl operator/(ddbl a, ddbl b) { return dd_div(a, b); }
GA_HD ddbl& operator+=(ddbl& a, ddbl b) { a = dd_add(a, b); return a; }
GA_HD ddbl& operator-=(ddbl& a, ddbl b) { a = dd_sub(a, b); return a; }
GA_HD ddbl& operator*=(ddbl& a, ddbl b) { a = dd_mul(a, b); return a; }
GA_HD bool operator==(ddbl a, ddbl b) { return a.hi == b.hi && a.lo == b.lo; }
GA_HD bool operator!=(ddbl a, ddbl b) { return !(a == b); }
GA_HD bool operator<(ddbl a, ddbl b) { return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo); }
GA_HD bool operator>(ddbl a, ddbl b) { return b < a; }
GA_HD bool operator<=(ddbl a, ddbl b) { return !(b < a); }
GA_HD bool operator>=(ddbl a, ddbl b) { return !(a < b); }

// ------------------------------------------------------------------ real-type helpers (TR)
GA_HD double r_abs(double x) { return fabs(x); }
GA_HD ddbl r_abs(ddbl x) { return (x.hi < 0.0 || (x.hi == 0.0 && x.lo < 0.0)) ? dd_neg(x) : x; }
GA_HD double r_sqrt(double x) { return sqrt(x); }
GA_HD ddbl r_sqrt(ddbl x) { return dd_sqrt(x); }
GA_HD double r_hi(double x) { return x; }
GA_HD double r_hi(ddbl x) { return x.hi; }
GA_HD double r_max(double a, double b) { return a < b ? b : a; }
GA_HD ddbl r_max(ddbl a, ddbl b) { return a < b ? b : a; }
GA_HD double r_min(double a, double b) { return b < a ? b : a; }
GA_HD ddbl r_min(ddbl a, ddbl b) { return b < a ? b : a; }
GA_HD bool r_isnan(double x) { return x != x; }
GA_HD bool r_isnan(ddbl x) { return x.hi != x.hi || x.lo != x.lo; }
GA_HD double r_copysign(double x, double y) { return copysign(x, y); }
GA_HD ddbl r_copysign(ddbl x, ddbl y) {
  bool xn = signbit(x.hi), yn = signbit(y.hi);
  return xn == yn ? x : dd_neg(x);
}

template <class TR> struct Real;
template <> struct Real<double> {
  static GA_HD double eps() { return 2.220446049250313e-16; }        // eps(Float64)
  static GA_HD double floatmin() { return 2.2250738585072014e-308; } // floatmin(Float64)
  static GA_HD double floatmin2() { return 1.0010415475915505e-146; } // 2^-485, LinearAlgebra.floatmin2
  static GA_HD double from(double x) { return x; }
};
template <> struct Real<ddbl> {
  static GA_HD ddbl eps() { return ddbl(4.930380657631324e-32); }       // 2^-104 = eps(Double64)
  static GA_HD ddbl floatmin() { return ddbl(2.0041683600089728e-292); } // 2^-969 = floatmin(Double64)
  static GA_HD ddbl floatmin2() { return ddbl(1.3435752215134178e-138); } // 2^-458, reference src/fixes.jl:5
  static GA_HD ddbl from(double x) { return ddbl(x); }
};

// ------------------------------------------------------------------ element traits
template <class T> struct Elem;

// a product that the compiler may NOT contract with a neighbouring add into an FMA: element
// products are then the same bits in every kernel that forms them (and in the CPU oracle, which
// is built with -ffp-contract=off)
GA_HD double mul_rn(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}

// Dot-product accumulator of the binary64 sweeps: Ogita-Rump-Oishi "Dot2".  Every product is formed
// exactly (TwoProd), the running sum is kept as an unevaluated pair by TwoSum, the rounding errors
// are collected in `lo`.  hi + lo is the dot product as if computed in twice the working precision,
// so the coefficients h = V'w do not depend (to ~1e-30 |v|'|w|) on how rows are split over threads,
// CTAs or GPUs: restart counts are reproducible across 1/2/4/8 GPUs and against the oracle, whose
// gemv_adj accumulates the same way (ARITH_EXACT).  10 FP64 instructions per product instead of 1; the
// sweeps are HBM-bound with FP64-pipe headroom, which is what pays for it (profiles/r02).
// GA_DOT_PLAIN (ga_set_option) selects one fused multiply-add per product instead, the arithmetic of an
// ordinary BLAS dgemv: faster, but the last bits of h then depend on the row partition.
struct dot2 {
  double hi, lo;
  GA_HD dot2() {}
  GA_HD explicit dot2(double h) : hi(h), lo(0.0) {}
};

template <> struct Elem<double> {
  typedef double real_t;
  typedef dot2 acc_t;
  static constexpr int dtype = 0;
  static constexpr bool is_complex = false;
  static GA_HD double zero() { return 0.0; }
  static GA_HD double re(double x) { return x; }
  static GA_HD double from_real(double r) { return r; }
  static GA_HD void add(double& a, double b) { a += b; }
  static GA_HD void acc_conj_mul(double& acc, double v, double w) { acc = fma(v, w, acc); }    // plain: acc += v w
  static GA_HD void acc_conj_mul(dot2& acc, double v, double w) {                              // acc += conj(v) w
    double p, e, s, t;
    two_prod(v, w, p, e);
    two_sum(acc.hi, p, s, t);
    acc.hi = s;
    acc.lo += (e + t);
  }
  static GA_HD void sub_mul(double& r, double v, double h) { r = fma(-v, h, r); }             // r -= v h
  static GA_HD void add_mul_real(double& y, double v, double q) { y = fma(v, q, y); }         // y += v q (q real)
  static GA_HD double mul_real(double v, double a) { return v * a; }
  static GA_HD double mul(double a, double b) { return mul_rn(a, b); }
};
template <> struct Elem<cdbl> {
  typedef double real_t;
  typedef cdbl acc_t;
  static constexpr int dtype = 1;
  static constexpr bool is_complex = true;
  static GA_HD cdbl zero() { cdbl z; z.re = 0.0; z.im = 0.0; return z; }
  static GA_HD double re(cdbl x) { return x.re; }
  static GA_HD cdbl from_real(double r) { cdbl z; z.re = r; z.im = 0.0; return z; }
  static GA_HD void add(cdbl& a, cdbl b) { a.re += b.re; a.im += b.im; }
  static GA_HD void acc_conj_mul(cdbl& acc, cdbl v, cdbl w) {
    acc.re = fma(v.re, w.re, acc.re); acc.re = fma(v.im, w.im, acc.re);
    acc.im = fma(v.re, w.im, acc.im); acc.im = fma(-v.im, w.re, acc.im);
  }
  static GA_HD void sub_mul(cdbl& r, cdbl v, cdbl h) {
    r.re = fma(-v.re, h.re, r.re); r.re = fma(v.im, h.im, r.re);
    r.im = fma(-v.re, h.im, r.im); r.im = fma(-v.im, h.re, r.im);
  }
  static GA_HD void add_mul_real(cdbl& y, cdbl v, double q) { y.re = fma(v.re, q, y.re); y.im = fma(v.im, q, y.im); }
  static GA_HD cdbl mul_real(cdbl v, double a) { cdbl z; z.re = v.re * a; z.im = v.im * a; return z; }
  static GA_HD cdbl mul(cdbl a, cdbl b) {
    cdbl z;
    z.re = fma(a.re, b.re, -mul_rn(a.im, b.im));
    z.im = fma(a.re, b.im, mul_rn(a.im, b.re));
    return z;
  }
};
template <> struct Elem<ddbl> {
  typedef ddbl real_t;
  typedef ddbl acc_t;
  static constexpr int dtype = 2;
  static constexpr bool is_complex = false;
  static GA_HD ddbl zero() { return ddbl(0.0, 0.0); }
  static GA_HD ddbl re(ddbl x) { return x; }
  static GA_HD ddbl from_real(ddbl r) { return r; }
  static GA_HD void add(ddbl& a, ddbl b) { a = dd_add(a, b); }
  static GA_HD void acc_conj_mul(ddbl& acc, ddbl v, ddbl w) { acc = dd_add(acc, dd_mul(v, w)); }
  static GA_HD void sub_mul(ddbl& r, ddbl v, ddbl h) { r = dd_sub(r, dd_mul(v, h)); }
  static GA_HD void add_mul_real(ddbl& y, ddbl v, ddbl q) { y = dd_add(y, dd_mul(v, q)); }
  static GA_HD ddbl mul_real(ddbl v, ddbl a) { return dd_mul(v, a); }
  static GA_HD ddbl mul(ddbl a, ddbl b) { return dd_mul(a, b); }
};

GA_HD float fmul_rn(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fmul_rn(a, b);
#else
  return a * b;
#endif
}
template <> struct Elem<float> {
  typedef double real_t;                 // TF = Float64: H, Q, norms, tolerances
  typedef double acc_t;                  // dot-product accumulators
  static constexpr int dtype = 3;
  static constexpr bool is_complex = false;
  static GA_HD float zero() { return 0.0f; }
  static GA_HD double re(float x) { return (double)x; }
  static GA_HD float from_real(double r) { return (float)r; }
  static GA_HD void add(float& a, float b) { a += b; }
  static GA_HD void acc_conj_mul(double& acc, float v, float w) { acc = fma((double)v, (double)w, acc); }
  static GA_HD void sub_mul(float& r, float v, float h) { r = fmaf(-v, h, r); }
  static GA_HD void add_mul_real(float& y, float v, double q) { y = fmaf(v, (float)q, y); }
  static GA_HD float mul_real(float v, double a) { return (float)((double)v * a); }   // Float32 * Float64, rounded on store
  static GA_HD float mul(float a, float b) { return fmul_rn(a, b); }
};

}  // namespace ga
