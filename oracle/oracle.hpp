// =====================================================================================
// TEST INFRASTRUCTURE ONLY.  CPU restatement ("oracle") of GenericArpack.jl's implicitly
// restarted Lanczos path.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may build, load or call this.  Nothing under
// genericarpack.jl_b200/ (the product) includes, links or executes it.
//
// What it restates (all citations relative to /root/reference, GenericArpack.jl v0.2.1):
//   src/arpack-blas.jl:25-76,457-544   _dlascl_factor/_scale_from_to/_dscal!, _dlarnv_idist_2!
//   src/arpack-blas-nrm2.jl:272-299    Float80.mynorm  (x87 80-bit == `long double` on x86-64)
//   src/arpack-blas-nrm2.jl:333-365    _dlassq + generic norm2 (used for Double64)
//   src/dgetv0.jl:504-753, src/dsaitr.jl:962-1535, src/dsapps.jl:577-869,
//   src/dsaup2.jl:962-1465, src/dsaupd.jl:921-1088, src/simple.jl (dsconv/dsortr/dsgets/dseigt!),
//   src/dstqrb.jl:607-1052, src/dseupd.jl:986-1532, src/arpack-blas-qr.jl, src/interface.jl:215-337,
//   src/idonow_ops.jl:216-333 (normal operator, singular vectors).
// The reverse-communication state machines are restated in their `idonow` (single call)
// form; the reference's own test/dsaupd_idonow.jl requires both forms to be identical.
//
// Parity pinning: see oracle/README.md.  Bit-level pin: the dlarnv start vector
// (test/dgetv0_simple.jl:6-35).  Eigenvalue/invariant-level pins: the KATs of
// test/interface_simple.jl, test/dsaupd_*.jl, test/dstqrb.jl, test/dsaitr_simple.jl.
// BLAS reduction order (OpenBLAS dgemv, SparseArrays.mul!) is outside /root/reference and
// unspecified, so V/H bits are "parity unpinned" (SURVEY.md section 8c).
// =====================================================================================
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "dd.hpp"

namespace oracle {

using i64 = long long;
using cplx = std::complex<double>;

// ------------------------------------------------------------------ scalar helpers
static inline double abs(double x) { return std::fabs(x); }
static inline double abs(float x) { return std::fabs((double)x); }   // abs feeding a Float64 accumulator
static inline double abs(std::complex<double> x) { return std::abs(x); }
static inline double sqrt(double x) { return std::sqrt(x); }
static inline bool isnan(double x) { return std::isnan(x); }
static inline double copysign(double x, double y) { return std::copysign(x, y); }
static inline double max(double a, double b) { return a < b ? b : a; }
static inline double min(double a, double b) { return b < a ? b : a; }

template <class TR> struct rt;
template <> struct rt<double> {
  static double eps() { return 2.220446049250313e-16; }        // eps(Float64)
  static double floatmin() { return 2.2250738585072014e-308; } // floatmin(Float64)
  static double floatmin2() { return std::ldexp(1.0, -485); }  // LinearAlgebra.floatmin2
  static int dstqrb_maxit() { return 30; }                     // src/dstqrb.jl:659
  static double eps23() { return std::pow(eps() / 2, 2.0 / 3.0); }  // src/macros.jl:11-13
};
template <> struct rt<dd> {
  static dd eps() { return dd(std::ldexp(1.0, -104)); }        // eps(Double64)
  static dd floatmin() { return dd(std::ldexp(1.0, -969)); }   // floatmin(Double64)
  static dd floatmin2() { return dd(std::ldexp(1.0, -458)); }  // src/fixes.jl:5
  static int dstqrb_maxit() { return 80; }                     // src/fixes.jl:6
  static dd eps23() { return dd(std::ldexp(1.0, -70)); }       // (2^-105)^(2/3) exactly
};

template <class T> struct vt;
template <> struct vt<double> {
  using real_t = double;
  static constexpr bool is_complex = false;
  static double re(double x) { return x; }
  static double conj(double x) { return x; }
};
// Float32 vectors with a Float64 factorisation: symeigs(Float32, Float64, op, nev) (src/interface.jl:26-34).  Element
// arithmetic (dot products, axpy, products of the sparse matrix) runs in binary32 like the reference's
// mul!(::Vector{Float32}, ...) calls; H, norms and every host quantity are Float64 (TR = TF).
template <> struct vt<float> {
  using real_t = double;
  static constexpr bool is_complex = false;
  static double re(float x) { return (double)x; }
  static float conj(float x) { return x; }
};
template <> struct vt<cplx> {
  using real_t = double;
  static constexpr bool is_complex = true;
  static double re(cplx x) { return x.real(); }
  static cplx conj(cplx x) { return std::conj(x); }
};
template <> struct vt<dd> {
  using real_t = dd;
  static constexpr bool is_complex = false;
  static dd re(dd x) { return x; }
  static dd conj(dd x) { return x; }
};

static inline double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct Stats {  // src/macros.jl:173-198
  i64 nopx = 0, nbx = 0, nrorth = 0, nitref = 0, nrstrt = 0;
  double taupd = 0, taup2 = 0, taitr = 0, teigt = 0, tgets = 0, tapps = 0, tconv = 0, tmvopx = 0,
         tgetv0 = 0, titref = 0, trvec = 0;
};

// ------------------------------------------------------------------ dlarnv LCG
// src/arpack-blas.jl:281-408 is a 128 x 4 table of 12-bit limbs; row i holds a^i mod 2^48 for
// a = row 1 = (494,322,2508,2549).  We regenerate the table from row 1 (checked against rows
// 2, 3, 128 given literally in tests/test_oracle_lcg.py).
struct LcgTable {
  int mm[128][4];
  LcgTable() {
    const unsigned long long mask = (1ULL << 48) - 1;
    const unsigned long long a = (494ULL << 36) | (322ULL << 24) | (2508ULL << 12) | 2549ULL;
    unsigned long long p = 1;
    for (int i = 0; i < 128; ++i) {
      p = (p * a) & mask;
      mm[i][0] = (int)((p >> 36) & 4095);
      mm[i][1] = (int)((p >> 24) & 4095);
      mm[i][2] = (int)((p >> 12) & 4095);
      mm[i][3] = (int)(p & 4095);
    }
  }
};
static inline const LcgTable& lcg_table() {
  static LcgTable t;
  return t;
}

// src/arpack-blas.jl:457-544.  x gets n values uniform(-1,1); complex takes two draws per
// element (re first).  iseed is 4 limbs of 12 bits, most significant first, updated on exit.
template <class T>
static void dlarnv_idist_2(i64 iseed[4], i64 n, T* x) {
  using TR = typename vt<T>::real_t;
  const LcgTable& tab = lcg_table();
  const bool iscomplex = vt<T>::is_complex;
  const i64 maxloop = iscomplex ? 2 * n : n;
  const i64 IPW2 = 4096;
  const double R = 1.0 / 4096.0;
  i64 i1 = iseed[0], i2 = iseed[1], i3 = iseed[2], i4 = iseed[3];
  i64 IT1 = i1, IT2 = i2, IT3 = i3, IT4 = i4;
  int nrv = 0;
  double lastrv = 0.0;
  for (i64 i = 1; i <= maxloop; ++i) {
    double rv = 1.0;
    nrv += 1;
    if (nrv > 128) {
      nrv = 1;
      i1 = IT1; i2 = IT2; i3 = IT3; i4 = IT4;
    }
    const int* m = tab.mm[nrv - 1];
    while (true) {
      IT4 = i4 * m[3];
      IT3 = IT4 / IPW2;
      IT4 = IT4 - IPW2 * IT3;
      IT3 = IT3 + i3 * m[3] + i4 * m[2];
      IT2 = IT3 / IPW2;
      IT3 = IT3 - IPW2 * IT2;
      IT2 = IT2 + i2 * m[3] + i3 * m[2] + i4 * m[1];
      IT1 = IT2 / IPW2;
      IT2 = IT2 - IPW2 * IT1;
      IT1 = IT1 + i1 * m[3] + i2 * m[2] + i3 * m[1] + i4 * m[0];
      IT1 = IT1 % IPW2;
      rv = R * ((double)IT1 + R * ((double)IT2 + R * ((double)IT3 + R * (double)IT4)));
      if (rv == 1.0) {  // unreachable in 53-bit arithmetic; kept for fidelity
        i1 += 2; i2 += 2; i3 += 2; i4 += 2;
      } else {
        break;
      }
    }
    if constexpr (vt<T>::is_complex) {
      if (i % 2 == 0) x[i / 2 - 1] = T(2 * lastrv - 1, 2 * rv - 1);
    } else {
      x[i - 1] = T(TR(2 * rv - 1));
    }
    lastrv = rv;
  }
  iseed[0] = IT1; iseed[1] = IT2; iseed[2] = IT3; iseed[3] = IT4;
}

// The same routine for RT = Float32 (mixed precision / Float32 vectors, src/interface.jl:26-34): the conversion
// R*(RT(IT1)+R*(RT(IT2)+R*(RT(IT3)+R*RT(IT4)))) is evaluated in binary32, where it CAN round up to exactly 1
// (X >= 2^48 - 2^23, about one draw in 2^25); the reference then adds 2 to every limb of the current block's base
// seed and redraws (src/arpack-blas.jl:520-525), which changes the rest of the stream.  Built with
// -ffp-contract=off, so these are plain IEEE binary32 operations.  Also counts the reseeds taken.
static i64 dlarnv_idist_2_f32(i64 iseed[4], i64 n, float* x) {
  const LcgTable& tab = lcg_table();
  const i64 IPW2 = 4096;
  const float R = 1.0f / 4096.0f;
  i64 i1 = iseed[0], i2 = iseed[1], i3 = iseed[2], i4 = iseed[3];
  i64 IT1 = i1, IT2 = i2, IT3 = i3, IT4 = i4;
  int nrv = 0;
  i64 reseeds = 0;
  for (i64 i = 1; i <= n; ++i) {
    volatile float rv = 1.0f;
    nrv += 1;
    if (nrv > 128) {
      nrv = 1;
      i1 = IT1; i2 = IT2; i3 = IT3; i4 = IT4;
    }
    const int* m = tab.mm[nrv - 1];
    while (true) {
      IT4 = i4 * m[3];
      IT3 = IT4 / IPW2;
      IT4 = IT4 - IPW2 * IT3;
      IT3 = IT3 + i3 * m[3] + i4 * m[2];
      IT2 = IT3 / IPW2;
      IT3 = IT3 - IPW2 * IT2;
      IT2 = IT2 + i2 * m[3] + i3 * m[2] + i4 * m[1];
      IT1 = IT2 / IPW2;
      IT2 = IT2 - IPW2 * IT1;
      IT1 = IT1 + i1 * m[3] + i2 * m[2] + i3 * m[1] + i4 * m[0];
      IT1 = IT1 % IPW2;
      volatile float a4 = R * (float)IT4;
      volatile float a3 = (float)IT3 + a4;
      volatile float b3 = R * a3;
      volatile float a2 = (float)IT2 + b3;
      volatile float b2 = R * a2;
      volatile float a1 = (float)IT1 + b2;
      rv = R * a1;
      if (rv == 1.0f) {
        i1 += 2; i2 += 2; i3 += 2; i4 += 2;
        reseeds += 1;
      } else {
        break;
      }
    }
    volatile float two_rv = 2.0f * rv;
    x[i - 1] = two_rv - 1.0f;
  }
  iseed[0] = IT1; iseed[1] = IT2; iseed[2] = IT3; iseed[3] = IT4;
  return reseeds;
}

// ------------------------------------------------------------------ arithmetic mode
// The reference's Float64 V-sweeps are OpenBLAS dgemv calls whose summation order is unspecified and its
// norm2 has two code paths: x87 80-bit accumulators on x86 (src/arpack-blas-nrm2.jl:272-299) and a
// double-double variant everywhere else (:157-215).  The oracle restates both:
//   ARITH_PLAIN (default)  plain binary64 dots/updates in column order + the x87 norm (what an x86 run does;
//                          also what the CPU timing legs of bench.py use)
//   ARITH_EXACT            dots accumulated with error-free TwoProd/TwoSum ("Dot2": the result is the exact
//                          dot product rounded once, to ~1e-30), the reference's own double-double norm
//                          (:157-215), and r -= V h summed in four column-interleaved partial sums.  Every
//                          reduction is then within one rounding of its exact value, so the trajectory (0.717
//                          tests, deflation, restart counts) does not depend on thread count or blocking -- and the
//                          B200 kernels, which accumulate the same way, reproduce it bit for bit in practice
//                          (tests/test_gpu_parity.py).  Only the binary64 real instantiation has this mode.
enum { ARITH_PLAIN = 0, ARITH_EXACT = 1 };
static inline int& arith_mode() { static int m = ARITH_PLAIN; return m; }

// ------------------------------------------------------------------ norm2
// src/arpack-blas-nrm2.jl:157-215 (_dnrm2_unroll_ext_dd, the non-x86 path): exact squares (two_prod)
// accumulated in double-double in four interleaved accumulators, Newton-refined double-double square root.
static inline double mynorm_dd(const double* a, i64 len) {
  double mx = 0.0;
  for (i64 i = 0; i < len; ++i) mx = std::fabs(a[i]) > mx ? std::fabs(a[i]) : mx;       // :159
  const double scale = (mx >= std::sqrt(std::nextafter(INFINITY, 0.0))) ? std::ldexp(1.0, -511) : 1.0;  // :163-168
  auto add_sum_sq = [](dd acc, double b) { double p, e; two_prod(b, b, p, e); return acc + dd(p, e); };  // :139-141
  dd ss1, ss2, ss3, ss4;
  i64 off = 0;
  while (off + 7 < len) {                                                             // :175-185
    ss1 = add_sum_sq(ss1, a[off + 0] * scale); ss2 = add_sum_sq(ss2, a[off + 1] * scale);
    ss3 = add_sum_sq(ss3, a[off + 2] * scale); ss4 = add_sum_sq(ss4, a[off + 3] * scale);
    ss1 = add_sum_sq(ss1, a[off + 4] * scale); ss2 = add_sum_sq(ss2, a[off + 5] * scale);
    ss3 = add_sum_sq(ss3, a[off + 6] * scale); ss4 = add_sum_sq(ss4, a[off + 7] * scale);
    off += 8;
  }
  for (i64 i = off; i < len; ++i) ss1 = add_sum_sq(ss1, a[i] * scale);                // :186-188
  ss1 = ss1 + ss3; ss1 = ss1 + ss2; ss1 = ss1 + ss4;                                  // :189-191
  if (ss1.hi == 0.0) return 0.0;                                                      // :193-195
  // :199-212 (sqrt_dd_dd of DoubleFloats.jl, two Newton steps on 1/sqrt)
  auto sub_fpdd = [](double x, dd y) {                                                // two_diff(a,b,c) :118-124
    double s, t, xx, u, yy;
    { double a0 = -y.hi, b0 = y.lo; s = a0 - b0; double a1 = s + b0, b1 = s - a1; t = (a0 - a1) - (b0 + b1); }
    two_sum(x, s, xx, u);
    yy = u + t;
    double h, l; two_sum(xx, yy, h, l);
    return dd(h, l);
  };
  auto add_fpdd = [](double x, dd y) {                                                // two_sum(a,b,c) :64-70
    double t0, t1, hi, t2; two_sum(x, y.hi, t0, t1); two_sum(t0, y.lo, hi, t2);
    double lo = t2 + t1, h, l; two_hilo_sum(hi, lo, h, l);
    return dd(h, l);
  };
  const double rf = 1.0 / std::sqrt(ss1.hi);
  const dd h(ss1.hi * 0.5, ss1.lo * 0.5);
  double p, e; two_prod(rf, rf, p, e);
  dd r2(p, e);
  dd hr2 = h * r2;
  dd radj = sub_fpdd(0.5, hr2);
  radj = radj * dd(rf, 0.0);
  dd rdd = add_fpdd(rf, radj);
  r2 = rdd * rdd;
  hr2 = h * r2;
  radj = sub_fpdd(0.5, hr2);
  radj = radj * rdd;
  rdd = rdd + radj;
  rdd = rdd * ss1;
  rdd = rdd * dd(1.0 / scale, 0.0);
  return rdd.hi;
}
// src/arpack-blas-nrm2.jl:272-299 (Float80.mynorm): four x87 accumulators, unrolled by 8.
static inline double mynorm_x87(const double* a, i64 len) {
  long double ss1 = 0, ss2 = 0, ss3 = 0, ss4 = 0;
  i64 off = 0;
  while (off + 7 < len) {
    ss1 += (long double)a[off + 0] * (long double)a[off + 0];
    ss2 += (long double)a[off + 1] * (long double)a[off + 1];
    ss3 += (long double)a[off + 2] * (long double)a[off + 2];
    ss4 += (long double)a[off + 3] * (long double)a[off + 3];
    ss1 += (long double)a[off + 4] * (long double)a[off + 4];
    ss2 += (long double)a[off + 5] * (long double)a[off + 5];
    ss3 += (long double)a[off + 6] * (long double)a[off + 6];
    ss4 += (long double)a[off + 7] * (long double)a[off + 7];
    off += 8;
  }
  for (i64 i = off; i < len; ++i) ss1 += (long double)a[i] * (long double)a[i];
  ss1 += ss3;
  ss1 += ss2;
  ss1 += ss4;
  return (double)sqrtl(ss1);
}
// src/arpack-blas-nrm2.jl:333-365 (generic _dlassq path: Double64)
template <class TA>
static TA norm2_lassq(const TA* v, i64 n) {
  if (n == 1) return abs(v[0]);
  TA scale = TA(0.0), ssq = TA(1.0);
  for (i64 i = 0; i < n; ++i) {
    if (v[i] != TA(0.0)) {
      TA absxi = abs(v[i]);
      if (scale < absxi) {
        TA val = scale / absxi;
        ssq = TA(1.0) + ssq * val * val;
        scale = absxi;
      } else {
        TA val = absxi / scale;
        ssq += val * val;
      }
    }
  }
  return scale * sqrt(ssq);
}
static inline double norm2(const double* v, i64 n) {                                   // :368 with :309-326
  return arith_mode() == ARITH_EXACT ? mynorm_dd(v, n) : mynorm_x87(v, n);
}
static inline double norm2(const cplx* v, i64 n) { return mynorm_x87((const double*)v, 2 * n); }  // :369
static inline dd norm2(const dd* v, i64 n) { return norm2_lassq<dd>(v, n); }           // :355-365
static inline double norm2(const float* v, i64 n) {   // norm2(Float64, ::Vector{Float32}): _dlassq with TA = Float64 (:333-365)
  if (n == 1) return std::fabs((double)v[0]);
  double scale = 0.0, ssq = 1.0;
  for (i64 i = 0; i < n; ++i) {
    if (v[i] != 0.0f) {
      const double absxi = std::fabs((double)v[i]);
      if (scale < absxi) { const double val = scale / absxi; ssq = 1.0 + ssq * val * val; scale = absxi; }
      else { const double val = absxi / scale; ssq += val * val; }
    }
  }
  return scale * std::sqrt(ssq);
}

// ------------------------------------------------------------------ dlascl / dscal
// src/arpack-blas.jl:25-59
template <class TR>
static void dlascl_factor(TR cfrom, TR cto, TR& mul, TR& cfromc, TR& ctoc, bool& done) {
  TR smlnum = rt<TR>::floatmin();
  TR bignum = TR(1.0) / smlnum;
  cfromc = cfrom;
  ctoc = cto;
  TR cfrom1 = cfromc * smlnum;
  if (cfrom1 == cfromc) {
    mul = ctoc / cfromc;
    done = true;
  } else {
    TR cto1 = ctoc / bignum;
    if (cto1 == ctoc) {
      mul = ctoc;
      done = true;
      cfromc = TR(1.0);
    } else if (abs(cfrom1) > abs(ctoc) && ctoc != TR(0.0)) {
      mul = smlnum;
      done = false;
      cfromc = cfrom1;
    } else if (abs(cto1) > abs(cfromc)) {
      mul = bignum;
      done = false;
      ctoc = cto1;
    } else {
      mul = ctoc / cfromc;
      done = true;
    }
  }
}
// src/arpack-blas.jl:61-72
// v .*= a as the reference's broadcast!(*, v, v, a) evaluates it: for Float32 vectors and a Float64 factor the
// product is formed in Float64 and rounded on store
template <class T, class TS>
static inline T scale_elem(T v, TS a) {
  if constexpr (std::is_same<T, float>::value) return (float)((double)v * (double)a);
  else return v * T(a);
}
template <class T, class TR>
static void scale_from_to(TR from, TR to, T* v, i64 n) {
  if (from == to) return;
  TR mul, cfromc, ctoc;
  bool done;
  dlascl_factor(from, to, mul, cfromc, ctoc, done);
  for (i64 i = 0; i < n; ++i) v[i] = scale_elem(v[i], mul);
  while (!done) {
    TR f = cfromc, t = ctoc;
    dlascl_factor(f, t, mul, cfromc, ctoc, done);
    for (i64 i = 0; i < n; ++i) v[i] = scale_elem(v[i], mul);
  }
}
template <class T, class TS>
static void dscal(TS a, T* v, i64 n) {  // src/arpack-blas.jl:74-76
#pragma omp parallel for schedule(static) if (n > 100000)
  for (i64 i = 0; i < n; ++i) v[i] = scale_elem(v[i], a);
}
template <class T>
static void vcopy(T* dst, const T* src, i64 n) {
#pragma omp parallel for schedule(static) if (n > 100000)
  for (i64 i = 0; i < n; ++i) dst[i] = src[i];
}

// ------------------------------------------------------------------ plane rotation
// src/arpack-blas.jl:8-11 -> LinearAlgebra.givens(f,g,1,2) -> givensAlgorithm == LAPACK 3.x
// dlartg (scaling loops with safmn2 = floatmin2).  Julia stdlib, not under /root/reference.
template <class TR>
static void plane_rotation(TR f, TR g, TR& cs, TR& sn, TR& r) {
  const TR one = TR(1.0), zero = TR(0.0);
  TR safmn2 = rt<TR>::floatmin2();
  TR safmx2 = one / safmn2;
  if (g == zero) {
    cs = one; sn = zero; r = f;
  } else if (f == zero) {
    cs = zero; sn = one; r = g;
  } else {
    TR f1 = f, g1 = g;
    TR scale = max(abs(f1), abs(g1));
    if (scale >= safmx2) {
      int count = 0;
      while (true) {
        count++;
        f1 *= safmn2; g1 *= safmn2;
        scale = max(abs(f1), abs(g1));
        if (scale < safmx2 || count >= 20) break;
      }
      r = sqrt(f1 * f1 + g1 * g1);
      cs = f1 / r; sn = g1 / r;
      for (int i = 0; i < count; ++i) r *= safmx2;
    } else if (scale <= safmn2) {
      int count = 0;
      while (true) {
        count++;
        f1 *= safmx2; g1 *= safmx2;
        scale = max(abs(f1), abs(g1));
        if (scale > safmn2) break;
      }
      r = sqrt(f1 * f1 + g1 * g1);
      cs = f1 / r; sn = g1 / r;
      for (int i = 0; i < count; ++i) r *= safmn2;
    } else {
      r = sqrt(f1 * f1 + g1 * g1);
      cs = f1 / r; sn = g1 / r;
    }
    if (abs(f) > abs(g) && cs < zero) {
      cs = -cs; sn = -sn; r = -r;
    }
  }
}

// src/arpack-blas.jl:91-156
template <class TR>
static void dlaev2(TR a, TR b, TR c, TR& rt1, TR& rt2, TR& cs1, TR& sn1) {
  const TR one = TR(1.0), half = TR(0.5), zero = TR(0.0);
  TR sm = a + c, df = a - c, adf = abs(df), tb = b + b, ab = abs(tb);
  TR acmx, acmn;
  if (abs(a) > abs(c)) { acmx = a; acmn = c; } else { acmx = c; acmn = a; }
  TR rtv;
  if (adf > ab) { TR q = ab / adf; rtv = adf * sqrt(one + q * q); }
  else if (adf < ab) { TR q = adf / ab; rtv = ab * sqrt(one + q * q); }
  else rtv = ab * sqrt(TR(2.0));
  int sgn1, sgn2;
  if (sm < zero) { rt1 = half * (sm - rtv); sgn1 = -1; rt2 = (acmx / rt1) * acmn - (b / rt1) * b; }
  else if (sm > zero) { rt1 = half * (sm + rtv); sgn1 = 1; rt2 = (acmx / rt1) * acmn - (b / rt1) * b; }
  else { rt1 = half * rtv; rt2 = -half * rtv; sgn1 = 1; }
  TR cs;
  if (df >= zero) { cs = df + rtv; sgn2 = 1; } else { cs = df - rtv; sgn2 = -1; }
  TR acs = abs(cs);
  if (acs > ab) { TR ct = -tb / cs; sn1 = one / sqrt(one + ct * ct); cs1 = ct * sn1; }
  else {
    if (ab == zero) { cs1 = one; sn1 = zero; }
    else { TR tn = -cs / tb; cs1 = one / sqrt(one + tn * tn); sn1 = tn * cs1; }
  }
  if (sgn1 == sgn2) { TR tn = cs1; cs1 = -sn1; sn1 = tn; }
}
template <class TR>
static TR dlapy2(TR x, TR y) {  // src/arpack-blas.jl:159-169
  if (isnan(x)) return x;
  if (isnan(y)) return y;
  TR w = max(abs(x), abs(y)), z = min(abs(x), abs(y));
  if (z == TR(0.0)) return w;
  TR q = z / w;
  return w * sqrt(TR(1.0) + q * q);
}

// ------------------------------------------------------------------ tridiagonal QL/QR
// src/dstqrb.jl:670-1052.  z is either a vector (last row of the eigenvector matrix,
// dstqrb!) or an ldz x n matrix (simple_dsteqr!).  `zrows` = 1 -> vector mode.
template <class TR>
struct Steqr {
  int n_block;            // length of current block
  TR *d, *e;              // block views
  TR* z; int ldz; int zrows; bool isvector;  // z: block view (vector: z[i]; matrix: column i at z+i*ldz)
  TR* work;

  // apply rotation (c,s) to columns (j, j+1) of z : src/arpack-blas.jl:193-212
  void rot_cols(int j, TR ctemp, TR stemp) {
    if (ctemp != TR(1.0) || stemp != TR(0.0)) {
      if (isvector) {
        TR temp = z[j + 1];
        z[j + 1] = ctemp * temp - stemp * z[j];
        z[j] = stemp * temp + ctemp * z[j];
      } else {
        TR* a0 = z + (i64)j * ldz;
        TR* a1 = z + (i64)(j + 1) * ldz;
        for (int i = 0; i < zrows; ++i) {
          TR temp = a1[i];
          a1[i] = ctemp * temp - stemp * a0[i];
          a0[i] = stemp * temp + ctemp * a0[i];
        }
      }
    }
  }

  int ql_iteration(int maxit) {  // src/dstqrb.jl:793-924 (0-based)
    const TR eps = rt<TR>::eps() / TR(2.0), eps2 = eps * eps, safmin = rt<TR>::floatmin();
    const TR two = TR(2.0), one = TR(1.0), zero = TR(0.0);
    int n = n_block, iend = n - 1, is = 0, jtot = 0;
    while (is < iend) {
      int lastm = iend;
      for (int m = is; m <= iend - 1; ++m) {
        TR tst = abs(e[m]); tst = tst * tst;
        if (tst <= (eps2 * abs(d[m])) * abs(d[m + 1]) + safmin) { lastm = m; break; }
      }
      if (lastm < iend) e[lastm] = zero;
      if (lastm == is) {
        is += 1;
      } else if (lastm == is + 1) {
        TR rt1, rt2, c, s;
        dlaev2(d[is], e[is], d[is + 1], rt1, rt2, c, s);
        work[is] = c; work[is + n - 1] = s;
        rot_cols(is, c, s);
        d[is] = rt1; d[is + 1] = rt2; e[is] = zero;
        is += 2;
      } else {
        if (jtot == maxit) break;
        jtot += 1;
        TR p = d[is];
        TR g = (d[is + 1] - p) / (two * e[is]);
        TR r = dlapy2(g, one);
        g = d[lastm] - p + (e[is] / (g + copysign(r, g)));
        TR s = one, c = one;
        p = zero;
        for (int i = lastm - 1; i >= is; --i) {
          TR f = s * e[i], b = c * e[i];
          plane_rotation(g, f, c, s, r);
          if (i != lastm - 1) e[i + 1] = r;
          g = d[i + 1] - p;
          r = (d[i] - g) * s + two * c * b;
          p = s * r;
          d[i + 1] = g + p;
          g = c * r - b;
          work[i] = c; work[n - 1 + i] = -s;
        }
        // _apply_plane_rotations_right!(...; rev=true): columns lastm-1 .. is
        for (int j = lastm - 1; j >= is; --j) rot_cols(j, work[j], work[n - 1 + j]);
        d[is] -= p;
        e[is] = g;
      }
    }
    if (jtot == maxit) throw std::runtime_error("dstqrb: failed to converge eigenvalues (QL)");
    return jtot;
  }

  int qr_iteration(int maxit) {  // src/dstqrb.jl:926-1052 (0-based)
    const TR eps = rt<TR>::eps() / TR(2.0), eps2 = eps * eps, safmin = rt<TR>::floatmin();
    const TR two = TR(2.0), one = TR(1.0), zero = TR(0.0);
    int n = n_block, is = n - 1, iend = 0, jtot = 0;
    while (iend < is) {
      int lastm = iend;
      for (int m = is; m >= iend + 1; --m) {
        TR tst = abs(e[m - 1]); tst = tst * tst;
        if (tst <= (eps2 * abs(d[m])) * abs(d[m - 1]) + safmin) { lastm = m; break; }
      }
      if (lastm > iend) e[lastm - 1] = zero;
      if (lastm == is) {
        is -= 1;
      } else if (lastm == is - 1) {
        TR rt1, rt2, c, s;
        dlaev2(d[is - 1], e[is - 1], d[is], rt1, rt2, c, s);
        rot_cols(is - 1, c, s);
        d[is - 1] = rt1; d[is] = rt2; e[is - 1] = zero;
        is -= 2;
      } else {
        if (jtot == maxit) break;
        jtot += 1;
        TR p = d[is];
        TR g = (d[is - 1] - p) / (two * e[is - 1]);
        TR r = dlapy2(g, one);
        g = d[lastm] - p + (e[is - 1] / (g + copysign(r, g)));
        TR s = one, c = one;
        p = zero;
        for (int i = lastm; i <= is - 1; ++i) {
          TR f = s * e[i], b = c * e[i];
          plane_rotation(g, f, c, s, r);
          if (i != lastm) e[i - 1] = r;
          g = d[i] - p;
          r = (d[i + 1] - g) * s + two * c * b;
          p = s * r;
          d[i] = g + p;
          g = c * r - b;
          work[i] = c; work[n - 1 + i] = s;
        }
        for (int j = lastm; j <= is - 1; ++j) rot_cols(j, work[j], work[n - 1 + j]);
        d[is] -= p;
        e[is - 1] = g;
      }
    }
    if (jtot == maxit) throw std::runtime_error("dstqrb: failed to converge eigenvalues (QR)");
    return jtot;
  }
};

// src/dstqrb.jl:670-745.  isvector: z has n entries (init e_n); else z is n x n (ldz), init I.
template <class TR>
static void flexible_dsteqr(int n, TR* d, TR* e, TR* z, int ldz, bool isvector, TR* work) {
  const TR zero = TR(0.0), one = TR(1.0);
  if (isvector) {
    for (int i = 0; i < n; ++i) z[i] = zero;
    z[n - 1] = one;
  } else {
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) z[i + (i64)j * ldz] = (i == j) ? one : zero;
  }
  int jtot = 0;
  const int nmaxit = rt<TR>::dstqrb_maxit() * n;
  int l1 = 0;
  while (l1 < n) {
    if (l1 > 0) e[l1 - 1] = zero;
    // _find_next_block! (src/dstqrb.jl:747-759)
    int l = l1, lend = n - 1;
    {
      const TR eps = rt<TR>::eps() / TR(2.0);
      for (int m = l; m < n - 1; ++m) {
        TR tst = abs(e[m]);
        if (tst == zero) { lend = m; break; }
        else if (tst <= sqrt(abs(d[m])) * sqrt(abs(d[m + 1])) * eps) { e[m] = zero; lend = m; break; }
      }
    }
    if (lend > l) {
      // _process_block (src/dstqrb.jl:761-791)
      int nb = lend - l + 1;
      TR* db = d + l; TR* eb = e + l;
      const TR eps = rt<TR>::eps() / TR(2.0), eps2 = eps * eps, safmin = rt<TR>::floatmin();
      const TR safmax = one / safmin, ssfmax = sqrt(safmax) / TR(3.0), ssfmin = sqrt(safmin) / eps2;
      // opnorm1 of the block (src/dstqrb.jl:640-657)
      TR anorm = max(abs(db[0]) + abs(eb[0]), abs(db[nb - 1]) + abs(eb[nb - 2]));
      for (int j = 1; j < nb - 1; ++j) anorm = max(anorm, abs(db[j]) + abs(eb[j]) + abs(eb[j - 1]));
      TR scaleto = anorm > ssfmax ? ssfmax : (anorm < ssfmin ? ssfmin : anorm);
      scale_from_to(anorm, scaleto, db, (i64)nb);
      scale_from_to(anorm, scaleto, eb, (i64)nb - 1);
      Steqr<TR> s;
      s.n_block = nb; s.d = db; s.e = eb; s.work = work; s.isvector = isvector;
      s.ldz = ldz; s.zrows = isvector ? 1 : n;
      s.z = isvector ? (z + l) : (z + (i64)l * ldz);
      int niter;
      if (abs(db[nb - 1]) < abs(db[0])) niter = s.qr_iteration(nmaxit - jtot);
      else niter = s.ql_iteration(nmaxit - jtot);
      scale_from_to(scaleto, anorm, db, (i64)nb);
      scale_from_to(scaleto, anorm, eb, (i64)nb - 1);
      jtot += niter;
    }
    l1 = lend + 1;
  }
  // selection sort (src/dstqrb.jl:716-743)
  for (int ii = 1; ii < n; ++ii) {
    int i = ii - 1, k = i;
    TR p = d[i];
    for (int j = ii; j < n; ++j)
      if (d[j] < p) { k = j; p = d[j]; }
    if (k != i) {
      d[k] = d[i]; d[i] = p;
      if (isvector) { TR t = z[k]; z[k] = z[i]; z[i] = t; }
      else
        for (int r = 0; r < n; ++r) { TR t = z[r + (i64)k * ldz]; z[r + (i64)k * ldz] = z[r + (i64)i * ldz]; z[r + (i64)i * ldz] = t; }
    }
  }
}

// ------------------------------------------------------------------ simple.jl
enum Which { LM = 0, SM = 1, LA = 2, SA = 3, BE = 4 };

template <class TR>
static int dsconv(int n, const TR* ritz, const TR* bounds, TR tol) {  // src/simple.jl:37-80
  TR eps23 = rt<TR>::eps23();
  int nconv = 0;
  for (int i = 0; i < n; ++i) {
    TR tmp = max(eps23, abs(ritz[i]));
    if (bounds[i] <= tol * tmp) nconv++;
  }
  return nconv;
}
template <class TR>
static bool sort_cmp(Which w, TR x, TR y) {
  switch (w) {
    case SA: return x < y;
    case SM: return abs(x) < abs(y);
    case LA: return x > y;
    case LM: return abs(x) > abs(y);
    default: return false;
  }
}
// src/simple.jl:255-308; x2 either a vector (x2cols == 0) or a matrix whose columns are swapped.
template <class TR, class T2>
static void dsortr(Which which, bool apply, int n, TR* x1, T2* x2, i64 ld2 = 0, i64 rows2 = 0) {
  int igap = n / 2;
  while (igap != 0) {
    for (int i = igap; i <= n - 1; ++i) {
      int j = i - igap;
      while (j >= 0) {
        if (sort_cmp(which, x1[j], x1[j + igap])) {
          std::swap(x1[j], x1[j + igap]);
          if (apply) {
            if (rows2 == 0) std::swap(x2[j], x2[j + igap]);
            else for (i64 r = 0; r < rows2; ++r) std::swap(x2[r + ld2 * j], x2[r + ld2 * (j + igap)]);
          }
        } else break;
        j -= igap;
      }
    }
    igap /= 2;
  }
}
template <class TR>
static void swap_within_array(int n, TR* v, int offset1, int ioffset1 = 1) {  // 1-based offsets, src/simple.jl:534-542
  for (int i = 1; i <= n; ++i) std::swap(v[offset1 - 1 + i - 1], v[i + ioffset1 - 1 - 1]);
}
// src/simple.jl:603-702
template <class TR>
static void dsgets(int ishift, Which which, int kev, int np, TR* ritz, TR* bounds, TR* shifts) {
  if (which == BE) {
    dsortr<TR, TR>(LA, true, kev + np, ritz, bounds);
    int kevd2 = kev / 2;
    if (kev > 1) {
      swap_within_array(std::min(kevd2, np), ritz, std::max(kevd2, np) + 1);
      swap_within_array(std::min(kevd2, np), bounds, std::max(kevd2, np) + 1);
    }
  } else {
    dsortr<TR, TR>(which, true, kev + np, ritz, bounds);
  }
  if (ishift == 1 && np > 0) {
    dsortr<TR, TR>(SM, true, np, bounds, ritz);
    for (int i = 0; i < np; ++i) shifts[i] = ritz[i];
  }
}
// src/simple.jl:968-1035.  H is ldh x 2 (col 0 = subdiag from row 1, col 1 = diag).
template <class TR>
static int dseigt(TR rnorm, int n, const TR* H, int ldh, TR* eig, TR* bounds, TR* workl) {
  for (int i = 0; i < n; ++i) eig[i] = H[i + ldh];
  for (int i = 0; i < n - 1; ++i) workl[i] = H[i + 1];
  flexible_dsteqr<TR>(n, eig, workl, bounds, 1, true, workl + n);
  for (int k = 0; k < n; ++k) bounds[k] = rnorm * abs(bounds[k]);
  return 0;
}

// ------------------------------------------------------------------ operators
template <class T>
struct Op {  // src/idonow_ops.jl:50-88 : size, arpack_mode, bmat, opx!, genopx!, bx!
  i64 n = 0;
  int mode = 1;          // arpack_mode(op)
  bool bmat_g = false;   // bmat(op) == Val(:G)
  virtual ~Op() {}
  virtual void opx(T* y, const T* x) = 0;
  // mode 2 (src/idonow_ops.jl:471-487): y = inv(B) A x AND x <- A x (ARPACK remark 5)
  virtual void genopx(T* y, T* x) { opx(y, x); }
  virtual void bx(T* y, const T* x) { for (i64 i = 0; i < n; ++i) y[i] = x[i]; }
};
template <class T>
struct CsrOp : Op<T> {  // ArpackSimpleOp over a (full, both triangles) CSR matrix
  const i64* rowptr; const int* col; const T* val;
  CsrOp(i64 n_, const i64* rp, const int* c, const T* v) : rowptr(rp), col(c), val(v) { this->n = n_; }
  void opx(T* y, const T* x) override {
    const i64 n = this->n;
#pragma omp parallel for schedule(static) if (n > 20000)
    for (i64 i = 0; i < n; ++i) {
      T s = T(0.0);
      for (i64 k = rowptr[i]; k < rowptr[i + 1]; ++k) s += val[k] * x[col[k]];
      y[i] = s;
    }
  }
};
// ArpackNormalOp (src/idonow_ops.jl:216-277): A is m x n CSR; At_side = :left when m > n
// (OP = A^H A, size n), else :right (OP = A A^H, size m).
template <class T>
struct NormalOp : Op<T> {
  i64 m, ncols; const i64* rowptr; const int* col; const T* val; bool left; std::vector<T> extra;
  NormalOp(i64 m_, i64 n_, const i64* rp, const int* c, const T* v) : m(m_), ncols(n_), rowptr(rp), col(c), val(v) {
    left = !(m <= ncols);
    this->n = left ? ncols : m;
    extra.assign(left ? m : ncols, T(0.0));
  }
  void Ax(T* y, const T* x) {
#pragma omp parallel for schedule(static) if (m > 20000)
    for (i64 i = 0; i < m; ++i) {
      T s = T(0.0);
      for (i64 k = rowptr[i]; k < rowptr[i + 1]; ++k) s += val[k] * x[col[k]];
      y[i] = s;
    }
  }
  void Ahx(T* y, const T* x) {
    for (i64 j = 0; j < ncols; ++j) y[j] = T(0.0);
    for (i64 i = 0; i < m; ++i)
      for (i64 k = rowptr[i]; k < rowptr[i + 1]; ++k) y[col[k]] += vt<T>::conj(val[k]) * x[i];
  }
  void opx(T* y, const T* x) override {
    if (left) { Ax(extra.data(), x); Ahx(y, extra.data()); }
    else { Ahx(extra.data(), x); Ax(y, extra.data()); }
  }
};

// ArpackSymmetricGeneralizedOp(A, S, B) (src/idonow_ops.jl:462-492): mode 2, bmat = :G.  A and B are
// full (both triangles) CSR matrices, B Hermitian positive definite.  The reference's S is whatever
// factorisation the caller passes (factorize(B): LDLt / Cholesky / LU, outside /root/reference); it is
// restated here as a banded Cholesky B = L L^H (band = the bandwidth of B's pattern).
template <class T>
struct GeneralizedOp : Op<T> {
  using TR = typename vt<T>::real_t;
  const i64* arp; const int* acol; const T* aval;
  const i64* brp; const int* bcol; const T* bval;
  i64 kd = 0;            // half bandwidth of B
  std::vector<T> L;      // L(i, j) for i - kd <= j <= i at L[i * (kd + 1) + (i - j)]
  bool spd = true;
  GeneralizedOp(i64 n_, const i64* arp_, const int* acol_, const T* aval_, const i64* brp_, const int* bcol_,
                const T* bval_) : arp(arp_), acol(acol_), aval(aval_), brp(brp_), bcol(bcol_), bval(bval_) {
    this->n = n_; this->mode = 2; this->bmat_g = true;
    const i64 n = n_;
    for (i64 i = 0; i < n; ++i)
      for (i64 k = brp[i]; k < brp[i + 1]; ++k) kd = std::max<i64>(kd, i - (i64)bcol[k]);
    const i64 w = kd + 1;
    L.assign((size_t)n * w, T(0.0));
    for (i64 i = 0; i < n; ++i)
      for (i64 k = brp[i]; k < brp[i + 1]; ++k)
        if (bcol[k] <= i) L[(size_t)i * w + (i - bcol[k])] = bval[k];      // lower triangle of B
    for (i64 j = 0; j < n; ++j) {                                           // column-by-column Cholesky
      const i64 iend = std::min(n - 1, j + kd);
      for (i64 i = j; i <= iend; ++i) {
        T sum = L[(size_t)i * w + (i - j)];
        for (i64 k = std::max<i64>(0, i - kd); k < j; ++k) {
          if (j - k > kd) continue;
          sum -= L[(size_t)i * w + (i - k)] * vt<T>::conj(L[(size_t)j * w + (j - k)]);
        }
        if (i == j) {
          const TR d = vt<T>::re(sum);
          if (!(d > TR(0.0))) { spd = false; L[(size_t)j * w] = T(1.0); }
          else L[(size_t)j * w] = T(sqrt(d));
        } else {
          L[(size_t)i * w + (i - j)] = sum / L[(size_t)j * w];
        }
      }
    }
  }
  static void spmv(i64 n, const i64* rp, const int* col, const T* val, T* y, const T* x) {
#pragma omp parallel for schedule(static) if (n > 20000)
    for (i64 i = 0; i < n; ++i) {
      T s = T(0.0);
      for (i64 k = rp[i]; k < rp[i + 1]; ++k) s += val[k] * x[col[k]];
      y[i] = s;
    }
  }
  void solve(T* y) const {  // y <- inv(B) y
    const i64 n = this->n, w = kd + 1;
    for (i64 i = 0; i < n; ++i) {                                           // L z = y
      T sum = y[i];
      for (i64 k = std::max<i64>(0, i - kd); k < i; ++k) sum -= L[(size_t)i * w + (i - k)] * y[k];
      y[i] = sum / L[(size_t)i * w];
    }
    for (i64 i = n - 1; i >= 0; --i) {                                      // L^H x = z
      T sum = y[i];
      const i64 kend = std::min(n - 1, i + kd);
      for (i64 k = i + 1; k <= kend; ++k) sum -= vt<T>::conj(L[(size_t)k * w + (k - i)]) * y[k];
      y[i] = sum / L[(size_t)i * w];
    }
  }
  void genopx(T* y, T* x) override {                                        // :471-487
    spmv(this->n, arp, acol, aval, y, x);
    for (i64 i = 0; i < this->n; ++i) x[i] = y[i];
    solve(y);
  }
  void opx(T* y, const T* x) override {                                     // :488 (opx! = genopx!; x kept here)
    spmv(this->n, arp, acol, aval, y, x);
    solve(y);
  }
  void bx(T* y, const T* x) override { spmv(this->n, brp, bcol, bval, y, x); }  // :490-492
};

// dot(a, b) = sum conj(a_i) b_i (LinearAlgebra.dot -> BLAS ddot / zdotc, plain accumulation; the B-norms of
// bmat = :G are sqrt(abs(dot(r, B r))), src/dsaitr.jl:1195-1204)
template <class T>
static typename vt<T>::real_t bnorm_from_dot(const T* a, const T* b, i64 n) {
  T s = T(0.0);
  for (i64 i = 0; i < n; ++i) s += vt<T>::conj(a[i]) * b[i];
  return sqrt(typename vt<T>::real_t(abs(s)));
}

// ------------------------------------------------------------------ tall-skinny BLAS2
// The reference calls LinearAlgebra.mul! -> OpenBLAS dgemv/zgemv (Float64/ComplexF64) or
// generic_matvecmul! (Double64); neither is under /root/reference.  Restated as plain
// column dots / row updates; OpenMP splits rows statically so results are run-to-run
// deterministic for a fixed thread count.
// ARITH_EXACT: Ogita-Rump-Oishi Dot2 -- exact products, TwoSum accumulation, errors collected separately; the
// chunk results are combined in double-double, the final pair is rounded once.
static void gemv_adj_exact(i64 n, int j, const double* V, i64 ldv, const double* w, double* h) {
  const int NB = n > 50000 ? 64 : 1;
  std::vector<dd> part((size_t)NB * j);
#pragma omp parallel for schedule(static) if (NB > 1)
  for (int b = 0; b < NB; ++b) {
    const i64 lo = n * b / NB, hi = n * (b + 1) / NB;
    for (int c = 0; c < j; ++c) {
      const double* v = V + (i64)c * ldv;
      double sh = 0.0, sl = 0.0;
      for (i64 i = lo; i < hi; ++i) {
        double p, e, s, t;
        two_prod(v[i], w[i], p, e);
        two_sum(sh, p, s, t);
        sh = s;
        sl += (e + t);
      }
      double a, b2;
      two_sum(sh, sl, a, b2);
      part[(size_t)b * j + c] = dd(a, b2);
    }
  }
  for (int c = 0; c < j; ++c) {
    dd s;
    for (int b = 0; b < NB; ++b) s = s + part[(size_t)b * j + c];
    h[c] = s.hi + s.lo;
  }
}
// ARITH_EXACT: r_i = w_i + (p_0 + p_1 + p_2 + p_3), p_g = sum over columns c = g, g+4, ... of fma(-v_ic, h_c, .):
// four column-interleaved partial sums, the blocking the B200 sweep kernel uses for 8-byte elements
// (genericarpack.jl_b200/csrc/ga_gs_cfg.cuh: TG = 4).  Any fixed order is as legitimate as OpenBLAS's.
static void gemv_sub_exact(i64 n, int j, const double* V, i64 ldv, const double* h, double* r) {
#pragma omp parallel for schedule(static) if (n > 50000)
  for (i64 i = 0; i < n; ++i) {
    double p[4] = {0.0, 0.0, 0.0, 0.0};
    for (int c = 0; c < j; ++c) p[c & 3] = std::fma(-V[(i64)c * ldv + i], h[c], p[c & 3]);
    double x = r[i];
    x += p[0]; x += p[1]; x += p[2]; x += p[3];
    r[i] = x;
  }
}

template <class T>
static void gemv_adj(i64 n, int j, const T* V, i64 ldv, const T* w, T* h) {  // h = V[:,0:j]^H w
  if constexpr (std::is_same<T, double>::value) {
    if (arith_mode() == ARITH_EXACT) { gemv_adj_exact(n, j, V, ldv, w, h); return; }
  }
  if (n <= 50000) {
    for (int c = 0; c < j; ++c) {
      T s = T(0.0);
      const T* v = V + (i64)c * ldv;
      for (i64 i = 0; i < n; ++i) s += vt<T>::conj(v[i]) * w[i];
      h[c] = s;
    }
    return;
  }
  const int NB = 64;  // fixed number of row chunks -> result independent of thread count
  std::vector<T> part((size_t)NB * j);
#pragma omp parallel for schedule(static)
  for (int b = 0; b < NB; ++b) {
    i64 lo = n * b / NB, hi = n * (b + 1) / NB;
    for (int c = 0; c < j; ++c) {
      T s = T(0.0);
      const T* v = V + (i64)c * ldv;
      for (i64 i = lo; i < hi; ++i) s += vt<T>::conj(v[i]) * w[i];
      part[(size_t)b * j + c] = s;
    }
  }
  for (int c = 0; c < j; ++c) {
    T s = T(0.0);
    for (int b = 0; b < NB; ++b) s += part[(size_t)b * j + c];
    h[c] = s;
  }
}
template <class T>
static void gemv_sub(i64 n, int j, const T* V, i64 ldv, const T* h, T* r) {  // r -= V[:,0:j] h
  if constexpr (std::is_same<T, double>::value) {
    if (arith_mode() == ARITH_EXACT) { gemv_sub_exact(n, j, V, ldv, h, r); return; }
  }
#pragma omp parallel for schedule(static) if (n > 50000)
  for (i64 b = 0; b < (n + 1023) / 1024; ++b) {
    i64 lo = b * 1024, hi = std::min(n, lo + 1024);
    for (int c = 0; c < j; ++c) {
      const T* v = V + (i64)c * ldv;
      const T hc = h[c];
      for (i64 i = lo; i < hi; ++i) r[i] -= v[i] * hc;
    }
  }
}
template <class T, class TQ>
static void gemv_n(i64 n, int j, const T* V, i64 ldv, const TQ* q, T* out) {  // out = V[:,0:j] q
  if constexpr (std::is_same<T, double>::value && std::is_same<TQ, double>::value) {
    if (arith_mode() == ARITH_EXACT) {   // one fused multiply-add per column, ascending (the device kernel's order)
#pragma omp parallel for schedule(static) if (n > 50000)
      for (i64 i = 0; i < n; ++i) {
        double y = 0.0;
        for (int c = 0; c < j; ++c) y = std::fma(V[(i64)c * ldv + i], q[c], y);
        out[i] = y;
      }
      return;
    }
  }
#pragma omp parallel for schedule(static) if (n > 50000)
  for (i64 b = 0; b < (n + 1023) / 1024; ++b) {
    i64 lo = b * 1024, hi = std::min(n, lo + 1024);
    for (i64 i = lo; i < hi; ++i) out[i] = T(0.0);
    for (int c = 0; c < j; ++c) {
      const T* v = V + (i64)c * ldv;
      const T qc = T(q[c]);
      for (i64 i = lo; i < hi; ++i) out[i] += v[i] * qc;
    }
  }
}

// ------------------------------------------------------------------ solver state
template <class T>
struct Problem {
  using TR = typename vt<T>::real_t;
  i64 n; int ncv;
  Op<T>* op;
  std::vector<T> V, resid, workd;
  std::vector<TR> workl;  // ncv^2 + 8 ncv, laid out as src/dsaupd.jl:1016-1023
  i64 iseed[4] = {1, 3, 5, 7};  // src/macros.jl:291
  TR rnorm = TR(0.0);
  int nev0 = 0, np = 0, mxiter = 0;  // aupd_nev0 / aupd_np / aupd_mxiter
  i64 iparam[11] = {0};
  Stats stats;
  Problem(i64 n_, int ncv_, Op<T>* op_) : n(n_), ncv(ncv_), op(op_) {
    V.assign((size_t)n * ncv, T(0.0));
    resid.assign((size_t)n, T(0.0));
    workd.assign((size_t)3 * n, T(0.0));
    workl.assign((size_t)ncv * ncv + 8 * ncv, TR(0.0));
  }
};

// ------------------------------------------------------------------ dgetv0  (bmat = :I and :G)
// src/dgetv0.jl:504-753, idonow flow.  j is the 1-based column being generated.
template <class T>
static int dgetv0(Problem<T>& P, int itry, bool initv, int j) {
  using TR = typename vt<T>::real_t;
  (void)itry;
  const i64 n = P.n;
  const bool G = P.op->bmat_g;
  double t0 = now_s();
  T* resid = P.resid.data(); T* workd = P.workd.data(); const T* V = P.V.data();
  int ierr = 0, iter = 0;
  if (!initv) {                                                     // :551-553
    if constexpr (std::is_same<T, float>::value) dlarnv_idist_2_f32(P.iseed, n, resid);
    else dlarnv_idist_2<T>(P.iseed, n, resid);
  }
  TR rnorm0;
  if (G) {
    P.stats.nopx++;                                                 // :566-576: force the vector into the range of OP
    vcopy(workd, resid, n);
    P.op->genopx(workd + n, workd);
    P.stats.nbx++;                                                  // :596-606
    vcopy(resid, workd + n, n);
    P.op->bx(workd, workd + n);
    rnorm0 = bnorm_from_dot<T>(resid, workd, n);                    // :625-627
  } else {
    vcopy(workd, resid, n);                                         // :612
    rnorm0 = norm2(resid, n);                                       // :630
  }
  P.rnorm = rnorm0;
  if (j == 1) { P.stats.tgetv0 += now_s() - t0; return 0; }         // :633-637
  while (true) {
    gemv_adj<T>(n, j - 1, V, n, workd, workd + n);                  // :662
    gemv_sub<T>(n, j - 1, V, n, workd + n, resid);                  // :665
    if (G) {
      P.stats.nbx++;                                                // :669-679
      vcopy(workd + n, resid, n);
      P.op->bx(workd, workd + n);
      P.rnorm = bnorm_from_dot<T>(resid, workd, n);                 // :696-698
    } else {
      vcopy(workd, resid, n);                                       // :681
      P.rnorm = norm2(resid, n);                                    // :700
    }
    if (P.rnorm > TR(0.717) * rnorm0) break;                        // :711
    iter += 1;
    if (iter <= 5) { rnorm0 = P.rnorm; }                            // :716-720
    else {
      for (i64 i = 0; i < n; ++i) resid[i] = T(0.0);                // :726-728
      P.rnorm = TR(0.0);
      ierr = -1;
      break;
    }
  }
  P.stats.tgetv0 += now_s() - t0;
  return ierr;
}

// ------------------------------------------------------------------ dsaitr  (mode 1 bmat = :I; mode 2 bmat = :G)
// src/dsaitr.jl:962-1535, idonow flow.  H is ldh x 2.  Returns info (0 or invariant-subspace size).
template <class T>
static int dsaitr(Problem<T>& P, int k, int np, typename vt<T>::real_t* H, int ldh) {
  using TR = typename vt<T>::real_t;
  const i64 n = P.n;
  const TR safmin = rt<TR>::floatmin();
  double t0 = now_s();
  T* V = P.V.data(); T* resid = P.resid.data();
  T* wp = P.workd.data();          // ipj
  T* wr = P.workd.data() + n;      // irj
  T* wv = P.workd.data() + 2 * n;  // ivj
  int info = 0;
  const bool G = P.op->bmat_g;     // only with mode 2 (dsaupd rejects the other combinations)
  for (int j = k + 1; j <= k + np; ++j) {  // 1-based j as in the reference
    bool rstart = false;
    if (P.rnorm <= TR(0.0)) {                                        // :1061-1087, :1497-1525
      P.stats.nrstrt++;
      int itry = 1;
      rstart = true;
      while (true) {
        int ierr = dgetv0<T>(P, itry, false, j);
        if (ierr < 0) {
          itry++;
          if (itry <= 3) continue;
          info = j - 1;
          P.stats.taitr += now_s() - t0;
          return info;
        }
        break;
      }
    }
    T* vj = V + (i64)(j - 1) * n;
    vcopy(vj, resid, n);                                             // :1097
    if (P.rnorm >= safmin) {                                         // :1098-1105
      TR temp1 = TR(1.0) / P.rnorm;
      dscal(temp1, vj, n);
      dscal(temp1, wp, n);
    } else {
      scale_from_to(P.rnorm, TR(1.0), vj, n);
      scale_from_to(P.rnorm, TR(1.0), wp, n);
    }
    P.stats.nopx++;                                                  // :1113
    vcopy(wv, vj, n);                                                // :1115
    double t2 = now_s();
    if (G) P.op->genopx(wr, wv);                                     // :1128 (mode 2: wv <- A v_j)
    else P.op->opx(wr, wv);                                          // :1126
    P.stats.tmvopx += now_s() - t2;
    vcopy(resid, wr, n);                                             // :1142
    TR wnorm;
    if (G) {
      wnorm = bnorm_from_dot<T>(resid, wv, n);                       // :1195-1200: inv(B)-norm of A v_j
      gemv_adj<T>(n, j, V, n, wv, wr);                               // :1232
    } else {
      vcopy(wp, resid, n);                                           // :1173
      wnorm = norm2(resid, n);                                       // :1208
      gemv_adj<T>(n, j, V, n, wp, wr);                               // :1227
    }
    gemv_sub<T>(n, j, V, n, wr, resid);                              // :1240
    H[(j - 1) + ldh] = vt<T>::re(wr[j - 1]);                         // :1243
    if (j == 1 || rstart) H[j - 1] = TR(0.0); else H[j - 1] = P.rnorm;  // :1259-1263
    double t4 = now_s();
    if (G) {
      P.stats.nbx++;                                                 // :1272-1282
      vcopy(wr, resid, n);
      P.op->bx(wp, wr);
      P.rnorm = bnorm_from_dot<T>(resid, wp, n);                     // :1298-1300
    } else {
      vcopy(wp, resid, n);                                           // :1284
      P.rnorm = norm2(resid, n);                                     // :1304
    }
    if (!(P.rnorm > TR(0.717) * wnorm)) {                            // :1324
      int iter = 0;
      P.stats.nrorth++;                                              // :1335 (net of :1439)
      while (true) {
        gemv_adj<T>(n, j, V, n, wp, wr);                             // :1352
        gemv_sub<T>(n, j, V, n, wr, resid);                          // :1363
        if (j == 1 || rstart) H[j - 1] = TR(0.0);                    // :1364-1366
        H[(j - 1) + ldh] = H[(j - 1) + ldh] + vt<T>::re(wr[j - 1]);  // :1367
        TR rnorm1;
        if (G) {
          P.stats.nbx++;                                             // :1373-1384
          vcopy(wr, resid, n);
          P.op->bx(wp, wr);
          rnorm1 = bnorm_from_dot<T>(resid, wp, n);                  // :1405-1407
        } else {
          vcopy(wp, resid, n);                                       // :1386
          rnorm1 = norm2(resid, n);                                  // :1410
        }
        if (rnorm1 > TR(0.717) * P.rnorm) { P.rnorm = rnorm1; break; }  // :1424-1426
        P.stats.nitref++;                                            // :1430
        P.rnorm = rnorm1;
        iter += 1;
        if (iter <= 1) continue;                                     // :1433-1439
        for (i64 i = 0; i < n; ++i) resid[i] = T(0.0);               // :1442-1443
        P.rnorm = TR(0.0);
        break;
      }
    }
    P.stats.titref += now_s() - t4;
    if (H[j - 1] < TR(0.0)) {                                        // :1466-1473
      H[j - 1] = -H[j - 1];
      if (j < k + np) dscal(TR(-1.0), V + (i64)j * n, n);
      else dscal(TR(-1.0), resid, n);
    }
  }
  P.stats.taitr += now_s() - t0;
  return info;
}

// ------------------------------------------------------------------ dsapps
// src/dsapps.jl:577-869.  H ldh x 2, Q ldq x kplusp (real), shifts in `shift[0:np)`.
template <class T>
static void dsapps(Problem<T>& P, int kev, int np, const typename vt<T>::real_t* shift,
                   typename vt<T>::real_t* H, int ldh, typename vt<T>::real_t* Q, int ldq) {
  using TR = typename vt<T>::real_t;
  const i64 n = P.n;
  double t0 = now_s();
  T* V = P.V.data(); T* resid = P.resid.data(); T* workd = P.workd.data();
  const TR epsmch = rt<TR>::eps() / TR(2.0);
  const TR zero = TR(0.0), one = TR(1.0);
  int itop = 1;
  const int kplusp = kev + np;
#define Hs(i) H[(i)-1]            /* H[i,1] subdiag, 1-based */
#define Hd(i) H[(i)-1 + ldh]      /* H[i,2] diag */
#define Qm(i, j) Q[(i)-1 + (i64)((j)-1) * ldq]
  for (int j = 1; j <= kplusp; ++j)
    for (int i = 1; i <= kplusp; ++i) Qm(i, j) = (i == j) ? one : zero;   // :623
  for (int jj = 1; jj <= np; ++jj) {
    int istart = itop;
    while (istart <= kplusp) {
      int iend = kplusp;
      for (int i = istart; i <= kplusp - 1; ++i) {                       // :661-673
        TR big = abs(Hd(i)) + abs(Hd(i + 1));
        if (Hs(i + 1) <= epsmch * big) { Hs(i + 1) = zero; iend = i; break; }
      }
      if (istart < iend) {
        TR f = Hd(istart) - shift[jj - 1], g = Hs(istart + 1), c, s, r;   // :681-683
        plane_rotation(f, g, c, s, r);
        TR a1 = c * Hd(istart) + s * Hs(istart + 1);
        TR a2 = c * Hs(istart + 1) + s * Hd(istart + 1);
        TR a4 = c * Hd(istart + 1) - s * Hs(istart + 1);
        TR a3 = c * Hs(istart + 1) - s * Hd(istart);
        Hd(istart) = c * a1 + s * a2;
        Hd(istart + 1) = c * a4 - s * a3;
        Hs(istart + 1) = c * a3 + s * a4;
        for (int j = 1; j <= std::min(istart + jj, kplusp); ++j) {       // :698-702
          a1 = c * Qm(j, istart) + s * Qm(j, istart + 1);
          Qm(j, istart + 1) = -s * Qm(j, istart) + c * Qm(j, istart + 1);
          Qm(j, istart) = a1;
        }
        for (int i = istart + 1; i <= iend - 1; ++i) {                   // :713-756
          f = Hs(i);
          g = s * Hs(i + 1);
          Hs(i + 1) = c * Hs(i + 1);
          plane_rotation(f, g, c, s, r);
          if (r < zero) { r = -r; c = -c; s = -s; }
          Hs(i) = r;
          a1 = c * Hd(i) + s * Hs(i + 1);
          a2 = c * Hs(i + 1) + s * Hd(i + 1);
          a3 = c * Hs(i + 1) - s * Hd(i);
          a4 = c * Hd(i + 1) - s * Hs(i + 1);
          Hd(i) = c * a1 + s * a2;
          Hd(i + 1) = c * a4 - s * a3;
          Hs(i + 1) = c * a3 + s * a4;
          for (int j = 1; j <= std::min(i + jj, kplusp); ++j) {
            a1 = c * Qm(j, i) + s * Qm(j, i + 1);
            Qm(j, i + 1) = -s * Qm(j, i) + c * Qm(j, i + 1);
            Qm(j, i) = a1;
          }
        }
      }
      if (Hs(iend) < zero) {                                             // :766-769
        Hs(iend) = -Hs(iend);
        for (int j = 1; j <= kplusp; ++j) Qm(j, iend) = -Qm(j, iend);
      }
      istart = iend + 1;
    }
    for (int i = itop; i <= kplusp - 1; ++i) {                           // :781-787
      if (Hs(i + 1) > zero) break;
      itop += 1;
    }
  }
  for (int i = itop; i <= kplusp - 1; ++i) {                             // :796-805
    TR big = abs(Hd(i)) + abs(Hd(i + 1));
    if (Hs(i + 1) <= epsmch * big) Hs(i + 1) = zero;
  }
  if (Hs(kev + 1) > zero)                                                // :813-817
    gemv_n<T, TR>(n, kplusp, V, n, &Qm(1, kev + 1), workd + n);
  for (int i = 1; i <= kev; ++i) {                                       // :825-830
    int band = kplusp - i + 1;
    gemv_n<T, TR>(n, band, V, n, &Qm(1, kev - i + 1), workd);
    vcopy(V + (i64)(band - 1) * n, workd, n);
  }
  for (int i = 1; i <= kev; ++i) vcopy(V + (i64)(i - 1) * n, V + (i64)(np + i - 1) * n, n);  // :833-835
  if (Hs(kev + 1) > zero) vcopy(V + (i64)kev * n, workd + n, n);         // :839-841
  dscal(Qm(kplusp, kev), resid, n);                                      // :850
  if (Hs(kev + 1) > zero) {                                              // :851-854
    const T beta = T(Hs(kev + 1));
    const T* vk = V + (i64)kev * n;
    bool fused = false;
    if constexpr (std::is_same<T, double>::value) fused = arith_mode() == ARITH_EXACT;
    if (fused) {
      if constexpr (std::is_same<T, double>::value)
        for (i64 i = 0; i < n; ++i) resid[i] = std::fma(vk[i], beta, resid[i]);
    } else {
      for (i64 i = 0; i < n; ++i) resid[i] += vk[i] * beta;
    }
  }
#undef Hs
#undef Hd
#undef Qm
  P.stats.tapps += now_s() - t0;
}

// ------------------------------------------------------------------ dsaup2
// src/dsaup2.jl:962-1465, idonow flow, ishift = 1.  Returns info.
template <class T>
static int dsaup2(Problem<T>& P, Which which, int& nev, int& np, typename vt<T>::real_t tol, int ishift,
                  int& mxiter, typename vt<T>::real_t* H, int ldh, typename vt<T>::real_t* ritz,
                  typename vt<T>::real_t* bounds, typename vt<T>::real_t* Q, int ldq,
                  typename vt<T>::real_t* workl, bool initv) {
  using TR = typename vt<T>::real_t;
  const i64 n = P.n;
  double t0 = now_s();
  const TR eps23 = rt<TR>::eps23();
  const int nev0 = nev, np0 = np, kplusp = nev0 + np0;
  int nconv = 0, iter = 0, info = 0;
  bool ok = false;
  info = dgetv0<T>(P, 1, initv, 1);                                      // :1064
  if (P.rnorm == TR(0.0)) { info = -9; goto done; }                      // :1071-1076
  info = dsaitr<T>(P, 0, nev0, H, ldh);                                  // :1430
  if (info > 0) { np = info; mxiter = iter; info = -9999; goto done; }   // :1437-1448
  while (true) {
    iter += 1;                                                           // :1084
    info = dsaitr<T>(P, nev, np, H, ldh);                                // :1097
    if (info > 0) { np = info; mxiter = iter; info = -9999; goto done; } // :1104-1113
    {
      double te = now_s();
      dseigt<TR>(P.rnorm, kplusp, H, ldh, ritz, bounds, workl);          // :1122
      P.stats.teigt += now_s() - te;
    }
    for (int i = 0; i < kplusp; ++i) { workl[kplusp + i] = ritz[i]; workl[2 * kplusp + i] = bounds[i]; }  // :1133-1134
    nev = nev0; np = np0;                                                // :1145-1146
    dsgets<TR>(ishift, which, nev0, np0, ritz, bounds, workl);           // :1148
    for (int i = 0; i < nev0; ++i) workl[np0 + i] = bounds[np0 + i];     // :1155
    nconv = dsconv<TR>(nev0, ritz + np0, workl + np0, tol);              // :1156
    {
      int nptemp = np;                                                   // :1177-1183
      for (int j = 0; j < nptemp; ++j)
        if (bounds[j] == TR(0.0)) { np -= 1; nev += 1; }
    }
    if (nconv >= nev0 || iter > mxiter || np == 0) {                     // :1185
      if (which == BE) {                                                 // :1194-1211
        dsortr<TR, TR>(SA, true, kplusp, ritz, bounds);
        int nevd2 = nev0 / 2, nevm2 = nev0 - nevd2;
        if (nev > 1) {
          np = kplusp - nev0;
          swap_within_array(std::min(nevd2, np), ritz, std::max(kplusp - nevd2 + 1, kplusp - np + 1), nevm2 + 1);
          swap_within_array(std::min(nevd2, np), bounds, std::max(kplusp - nevd2 + 1, kplusp - np + 1), nevm2 + 1);
        }
      } else {                                                           // :1221-1232
        Which wprime = which == LM ? SM : which == SM ? LM : which == LA ? SA : LA;
        dsortr<TR, TR>(wprime, true, kplusp, ritz, bounds);
      }
      for (int j = 0; j < nev0; ++j) { TR temp = max(eps23, abs(ritz[j])); bounds[j] = bounds[j] / temp; }  // :1237-1240
      dsortr<TR, TR>(LA, true, nev0, bounds, ritz);                      // :1248-1249
      for (int j = 0; j < nev0; ++j) { TR temp = max(eps23, abs(ritz[j])); bounds[j] = bounds[j] * temp; }  // :1253-1256
      if (which == BE) dsortr<TR, TR>(LA, true, nconv, ritz, bounds);    // :1264-1276
      else dsortr<TR, TR>(which, true, nconv, ritz, bounds);
      H[0] = P.rnorm;                                                    // :1280
      if (iter > mxiter && nconv < nev) info = 1;                        // :1288-1290
      if (np == 0 && nconv < nev0) info = 2;                             // :1293-1295
      np = nconv;                                                        // :1297
      ok = true;
      break;
    } else if (nconv < nev && ishift == 1) {                             // :1300-1322
      int nevbef = nev;
      nev = nev + std::min(nconv, np / 2);
      if (nev == 1 && kplusp >= 6) nev = kplusp / 2;
      else if (nev == 1 && kplusp > 2) nev = 2;
      np = kplusp - nev;
      if (nevbef < nev) dsgets<TR>(ishift, which, nev, np, ritz, bounds, workl);
    }
    dsapps<T>(P, nev, np, ritz, H, ldh, Q, ldq);                         // :1374
    if (P.op->bmat_g) {
      P.stats.nbx++;                                                     // :1383-1394
      vcopy(P.workd.data() + n, P.resid.data(), n);
      P.op->bx(P.workd.data(), P.workd.data() + n);
      P.rnorm = bnorm_from_dot<T>(P.resid.data(), P.workd.data(), n);    // :1409-1411
    } else {
      vcopy(P.workd.data(), P.resid.data(), n);                          // :1397
      P.rnorm = norm2(P.resid.data(), n);                                // :1414
    }
  }
done:
  if (ok) { mxiter = iter; nev = nconv; }                                // :1455-1458
  P.stats.taup2 += now_s() - t0;
  return info;
}

// ------------------------------------------------------------------ dsaupd
// src/dsaupd.jl:921-1148 (idonow).  iparam/ipntr 0-based arrays holding the reference's
// 1-based slots shifted by one (iparam[0] = ishift, iparam[2] = mxiter, iparam[6] = mode ...).
template <class T>
static int dsaupd(Problem<T>& P, Which which, int nev, typename vt<T>::real_t tol, int mode, bool initv) {
  using TR = typename vt<T>::real_t;
  const int ncv = P.ncv;
  const i64 n = P.n;
  double t0 = now_s();
  int ierr = 0;
  int ishift = (int)P.iparam[0];
  P.mxiter = (int)P.iparam[2];
  if (n <= 0) ierr = -1;                                                 // :964-996
  else if (nev <= 0) ierr = -2;
  else if (ncv <= nev || ncv > n) ierr = -3;
  if (P.mxiter < 0) ierr = -4;
  if ((int)which < 0 || (int)which > 4) ierr = -5;
  if ((i64)P.workl.size() < (i64)ncv * ncv + 8 * ncv) ierr = -7;
  if (mode < 1 || mode > 5) ierr = -10;
  else if (ishift < 0 || ishift > 1) ierr = -12;
  else if (nev == 1 && which == BE) ierr = -13;
  else if (mode == 1 && P.op->bmat_g) ierr = -11;                        // :990-991
  if (mode > 2 || (mode == 2) != P.op->bmat_g) ierr = ierr ? ierr : -10; // oracle restates modes 1 (:I) and 2 (:G)
  if (ierr != 0) return ierr;
  if (tol <= TR(0.0)) tol = rt<TR>::eps() / TR(2.0);                     // :1009-1011
  const int ldh = ncv, ldq = ncv;
  const int ih = 0, ritz = ih + 2 * ldh, bounds = ritz + ncv, iq = bounds + ncv, iw = iq + ncv * ncv;  // :1016-1023
  P.np = ncv - nev;                                                      // :1032-1033
  P.nev0 = nev;
  std::fill(P.workl.begin(), P.workl.end(), TR(0.0));                    // :1036
  TR* wl = P.workl.data();
  ierr = dsaup2<T>(P, which, P.nev0, P.np, tol, ishift, P.mxiter, wl + ih, ldh, wl + ritz, wl + bounds,
                   wl + iq, ldq, wl + iw, initv);                        // :1056
  P.iparam[2] = P.mxiter;                                                // :1081-1087
  P.iparam[4] = P.np;
  P.iparam[8] = P.stats.nopx; P.iparam[9] = P.stats.nbx; P.iparam[10] = P.stats.nrorth;
  if (ierr == 2) ierr = 3;                                               // :1092-1094
  P.stats.taupd += now_s() - t0;
  return ierr;
}

// ------------------------------------------------------------------ Householder QR pieces
// src/arpack-blas-qr.jl:137-185 (_dlarfg!) for real TR
template <class TR>
static TR dlarfg(TR& alpha, TR* x, int len) {
  if (len <= 0) return TR(0.0);
  TR xnorm = norm2(x, (i64)len);
  TR tau = TR(0.0), beta = alpha;
  if (!(xnorm == TR(0.0))) {
    beta = -copysign(dlapy2(alpha, xnorm), alpha);
    TR safmin = rt<TR>::floatmin() / (rt<TR>::eps() / TR(2.0));
    int knt = 0;
    if (abs(beta) < safmin) {
      TR rsafmn = TR(1.0) / safmin;
      while (knt < 20) {
        knt++;
        for (int i = 0; i < len; ++i) x[i] = x[i] * rsafmn;
        beta = beta * rsafmn; alpha = alpha * rsafmn;
        if (abs(beta) >= safmin) break;
      }
      xnorm = norm2(x, (i64)len);
      beta = -copysign(dlapy2(alpha, xnorm), alpha);
    }
    tau = (beta - alpha) / beta;
    TR sc = TR(1.0) / (alpha - beta);
    for (int i = 0; i < len; ++i) x[i] = x[i] * sc;
    for (int j = 0; j < knt; ++j) beta = beta * safmin;
  }
  alpha = beta;
  return tau;
}
// _dgeqr2! (src/arpack-blas-qr.jl:48-74) on an m x ncol real matrix A (lda), small (ncv x nconv)
template <class TR>
static void dgeqr2(int m, int ncol, TR* A, int lda, TR* tau) {
  int k = std::min(m, ncol);
  std::vector<TR> work(ncol);
  for (int i = 0; i < k; ++i) {
    tau[i] = dlarfg<TR>(A[i + (i64)i * lda], A + i + 1 + (i64)i * lda, m - i - 1);
    if (i < ncol - 1) {
      TR aii = A[i + (i64)i * lda];
      A[i + (i64)i * lda] = TR(1.0);
      // _dlarf_left!: C = A[i:m, i+1:ncol];  w = C^T v ; C -= tau v w^T
      const TR* v = A + i + (i64)i * lda;
      int mv = m - i;
      if (tau[i] != TR(0.0)) {
        for (int c = i + 1; c < ncol; ++c) {
          TR* C = A + i + (i64)c * lda;
          TR s = TR(0.0);
          for (int r = 0; r < mv; ++r) s += C[r] * v[r];
          work[c] = s;
        }
        for (int c = i + 1; c < ncol; ++c) {
          TR* C = A + i + (i64)c * lda;
          if (work[c] != TR(0.0)) {
            TR temp = -tau[i] * work[c];
            for (int r = 0; r < mv; ++r) C[r] += v[r] * temp;
          }
        }
      }
      A[i + (i64)i * lda] = aii;
    }
  }
}
// dorm2r!(Val(:R), Val(:N), m, ncolC, k, A, tau, C, work): C <- C * H(1) ... H(k)
// (src/arpack-blas-qr.jl:409-476 with _dlarf_right! :293-322).  C is m x ncolC of T, A real.
template <class T, class TR>
static void dorm2r_right(i64 m, int ncolC, int k, TR* A, int lda, const TR* tau, T* C, i64 ldc, T* work) {
  for (int i = 0; i < k; ++i) {
    int ni = ncolC - i;
    TR aii = A[i + (i64)i * lda];
    A[i + (i64)i * lda] = TR(1.0);
    const TR* v = A + i + (i64)i * lda;
    if (tau[i] != TR(0.0)) {
      int lastv = ni;
      while (lastv > 0 && v[lastv - 1] == TR(0.0)) lastv--;
      if (lastv > 0) {
        T* Ci = C + (i64)i * ldc;
        gemv_n<T, TR>(m, lastv, Ci, ldc, v, work);                       // w = C(:,i:i+lastv) v
        for (int c = 0; c < lastv; ++c) {                                // C -= tau w v^T  (_dger!)
          if (v[c] != TR(0.0)) {
            T temp = T(-tau[i] * v[c]);
            T* Cc = Ci + (i64)c * ldc;
#pragma omp parallel for schedule(static) if (m > 100000)
            for (i64 r = 0; r < m; ++r) Cc[r] += work[r] * temp;
          }
        }
      }
    }
    A[i + (i64)i * lda] = aii;
  }
}
// dorm2r!(Val(:L), Val(:T), ...) on a single real column (src/dseupd.jl:1451-1453)
template <class TR>
static void dorm2r_left_t_vec(int m, int k, TR* A, int lda, const TR* tau, TR* c) {
  for (int i = 0; i < k; ++i) {
    TR aii = A[i + (i64)i * lda];
    A[i + (i64)i * lda] = TR(1.0);
    const TR* v = A + i + (i64)i * lda;
    if (tau[i] != TR(0.0)) {
      TR s = TR(0.0);
      for (int r = 0; r < m - i; ++r) s += c[i + r] * v[r];
      if (s != TR(0.0)) { TR temp = -tau[i] * s; for (int r = 0; r < m - i; ++r) c[i + r] += v[r] * temp; }
    }
    A[i + (i64)i * lda] = aii;
  }
}

// ------------------------------------------------------------------ simple_dseupd! (REGULR)
// src/dseupd.jl:986-1532 for mode 1.  d: nconv values out (ascending); Z: n x nconv out when rvec.
// Returns info.  Also leaves V <- V*Q like the reference (:1431).
template <class T>
static int dseupd(Problem<T>& P, bool rvec, typename vt<T>::real_t* d, T* Z, Which which, int nev,
                  typename vt<T>::real_t tol) {
  using TR = typename vt<T>::real_t;
  const int ncv = P.ncv;
  const i64 n = P.n;
  const int nconv = (int)P.iparam[4];
  if (tol <= TR(0.0)) tol = rt<TR>::eps() / TR(2.0);
  if (nconv == 0) return 0;                                              // :1029-1032
  TR* workl = P.workl.data();
  const int ldh = ncv, ldq = ncv;
  const int ih = 0, ritz = 2 * ncv, bounds = 3 * ncv;                    // ipntr[5..7] - 1
  const int ihd = bounds + ldh, ihb = ihd + ldh, iq = ihb + ldh, iw = iq + ldh * ncv;  // :1131-1135
  const int ipntr11 = 4 * ncv + ncv * ncv;                               // iw of dsaupd (0-based)
  const int irz = ipntr11 + ncv, ibd = irz + ncv;                        // :1163-1164
  const TR eps23 = rt<TR>::eps23();
  TR* Q = workl + iq;
  const TR rnorm = workl[ih];                                            // :1175
  std::vector<int> select(ncv, 0);
  if (rvec) {
    bool reord = false;
    for (int j = 0; j < ncv; ++j) { workl[bounds + j] = TR((double)(j + 1)); select[j] = 0; }  // :1195-1198
    int np = ncv - nev;
    dsgets<TR>(0, which, nev, np, workl + irz, workl + bounds, workl);   // :1212
    int numcnv = 0;
    for (int j = 1; j <= ncv; ++j) {                                     // :1228-1238
      TR temp1 = max(eps23, abs(workl[irz + ncv - j]));
      int jj = (int)(double)workl[bounds + ncv - j];
      if (numcnv < nconv && workl[ibd + jj - 1] <= tol * temp1) {
        select[jj - 1] = 1;
        numcnv++;
        if (jj > nconv) reord = true;
      }
    }
    if (numcnv != nconv) return -17;                                     // :1251-1254
    for (int i = 0; i < ncv - 1; ++i) workl[ihb + i] = workl[ih + 1 + i];  // :1261
    for (int i = 0; i < ncv; ++i) workl[ihd + i] = workl[ih + ldh + i];  // :1262
    flexible_dsteqr<TR>(ncv, workl + ihd, workl + ihb, Q, ldq, false, workl + iw);  // :1264
    if (reord) {                                                         // :1282-1320
      int leftptr = 1, rghtptr = ncv;
      while (leftptr < rghtptr) {
        if (select[leftptr - 1]) leftptr++;
        else if (!select[rghtptr - 1]) rghtptr--;
        else {
          std::swap(workl[ihd + leftptr - 1], workl[ihd + rghtptr - 1]);
          for (int r = 0; r < ncv; ++r) std::swap(Q[r + (i64)ncv * (leftptr - 1)], Q[r + (i64)ncv * (rghtptr - 1)]);
          leftptr++; rghtptr--;
        }
      }
    }
    for (int i = 0; i < nconv; ++i) d[i] = workl[ihd + i];               // :1325
  } else {
    for (int i = 0; i < nconv; ++i) d[i] = workl[ritz + i];              // :1328
    for (int i = 0; i < ncv; ++i) workl[ihd + i] = workl[ritz + i];      // :1329
  }
  if (rvec) dsortr<TR, TR>(LA, true, nconv, d, Q, ldq, ncv);             // :1346
  else for (int i = 0; i < ncv; ++i) workl[ihb + i] = workl[bounds + i]; // :1348
  if (rvec) {
    double tr = now_s();
    TR* tau = workl + iw + ncv;                                          // :1419
    dgeqr2<TR>(ncv, nconv, Q, ldq, tau);                                 // :1421
    dorm2r_right<T, TR>(n, ncv, nconv, Q, ldq, tau, P.V.data(), n, P.workd.data() + n);  // :1431
    for (int c = 0; c < nconv; ++c) vcopy(Z + (i64)c * n, P.V.data() + (i64)c * n, n);   // :1434
    for (int j = 0; j < ncv; ++j) workl[ihb + j] = TR(0.0);              // :1441-1444
    workl[ihb + ncv - 1] = TR(1.0);
    dorm2r_left_t_vec<TR>(ncv, nconv, Q, ldq, tau, workl + ihb);         // :1451
    for (int j = 0; j < nconv; ++j) workl[iw + ncv + j] = workl[ihb + j];  // :1461-1463
    for (int j = 0; j < ncv; ++j) workl[ihb + j] = rnorm * abs(workl[ihb + j]);  // :1468-1471
    P.stats.trvec += now_s() - tr;
  }
  return 0;
}

// ------------------------------------------------------------------ symeigs (src/interface.jl:215-290)
template <class T>
struct EigsResult {
  using TR = typename vt<T>::real_t;
  int info = 0, ierr = 0, nconv = 0, niter = 0;
  std::vector<TR> values, bounds;
  std::vector<T> vectors;  // n x nvectors
  int nvectors = 0;
};
template <class T>
static EigsResult<T> symeigs(Problem<T>& P, int nev, Which which, int maxiter, typename vt<T>::real_t tol,
                             const T* v0, bool ritzvec) {
  using TR = typename vt<T>::real_t;
  EigsResult<T> R;
  P.iparam[0] = 1; P.iparam[3] = 1; P.iparam[2] = maxiter; P.iparam[6] = P.op->mode;  // :253-256
  bool initv = false;
  if (v0) { vcopy(P.resid.data(), v0, P.n); initv = true; }              // :259-262
  R.info = dsaupd<T>(P, which, nev, tol, P.op->mode, initv);             // :265
  R.niter = (int)P.iparam[2];
  R.nconv = (int)P.iparam[4];
  if (R.info != 0 && R.info != 1) return R;                              // :269-273 (caller raises)
  int nconv = R.nconv;
  R.nvectors = ritzvec ? std::min(nev, nconv) : 0;                       // :277
  R.values.assign(nconv, TR(0.0));
  R.vectors.assign((size_t)P.n * std::max(R.nvectors, ritzvec ? nconv : 0), T(0.0));
  R.ierr = dseupd<T>(P, ritzvec, R.values.data(), R.vectors.data(), which, nev, tol);  // :281
  R.bounds.assign(nconv, TR(0.0));
  for (int i = 0; i < nconv; ++i) R.bounds[i] = P.workl[5 * P.ncv + i];  // :191-196
  return R;
}

}  // namespace oracle
