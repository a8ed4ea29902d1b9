// Host-side small dense kernels of the IRL driver (ncv x ncv work; stays on the host by design,
// BASELINE.json north_star).  Mirrors, for TR in {double, ga::ddbl}:
//   plane_rotation  src/arpack-blas.jl:8-11  (LinearAlgebra.givens == LAPACK dlartg, classic variant)
//   _dlaev2         src/arpack-blas.jl:91-156
//   _dlapy2_julia   src/arpack-blas.jl:159-169
//   _scale_from_to  src/arpack-blas.jl:25-72
//   dstqrb! / simple_dsteqr! = flexible_dsteqr!  src/dstqrb.jl:670-1052
// (all citations relative to /root/reference).  Written for the product; the test oracle under
// oracle/ is a separate restatement and is never included from here.
#pragma once
#include <cmath>
#include <stdexcept>
#include <vector>

#include "../ga_types.cuh"

namespace ga {
namespace host {

template <class TR> inline TR rzero() { return Real<TR>::from(0.0); }
template <class TR> inline TR rone() { return Real<TR>::from(1.0); }

template <class TR> struct HostConst;
template <> struct HostConst<double> {
  static int steqr_maxit() { return 30; }                                  // src/dstqrb.jl:659
  static double eps23() { return std::pow(Real<double>::eps() / 2, 2.0 / 3.0); }  // src/macros.jl:11-13
};
template <> struct HostConst<ddbl> {
  static int steqr_maxit() { return 80; }                                  // src/fixes.jl:6
  static ddbl eps23() { return ddbl(std::ldexp(1.0, -70)); }               // (2^-105)^(2/3)
};

// Givens rotation [cs sn; -sn cs] [f; g] = [r; 0]
template <class TR>
inline void givens(TR f, TR g, TR& cs, TR& sn, TR& r) {
  const TR zero = rzero<TR>(), one = rone<TR>();
  const TR tiny2 = Real<TR>::floatmin2(), huge2 = one / tiny2;
  if (g == zero) { cs = one; sn = zero; r = f; return; }
  if (f == zero) { cs = zero; sn = one; r = g; return; }
  TR f1 = f, g1 = g;
  TR scale = r_max(r_abs(f1), r_abs(g1));
  int count = 0;
  if (scale >= huge2) {
    do { ++count; f1 = f1 * tiny2; g1 = g1 * tiny2; scale = r_max(r_abs(f1), r_abs(g1)); } while (scale >= huge2 && count < 20);
    r = r_sqrt(f1 * f1 + g1 * g1);
    cs = f1 / r; sn = g1 / r;
    for (int i = 0; i < count; ++i) r = r * huge2;
  } else if (scale <= tiny2) {
    do { ++count; f1 = f1 * huge2; g1 = g1 * huge2; scale = r_max(r_abs(f1), r_abs(g1)); } while (scale <= tiny2);
    r = r_sqrt(f1 * f1 + g1 * g1);
    cs = f1 / r; sn = g1 / r;
    for (int i = 0; i < count; ++i) r = r * tiny2;
  } else {
    r = r_sqrt(f1 * f1 + g1 * g1);
    cs = f1 / r; sn = g1 / r;
  }
  if (r_abs(f) > r_abs(g) && cs < zero) { cs = -cs; sn = -sn; r = -r; }
}

template <class TR>
inline TR hypot2(TR x, TR y) {
  if (r_isnan(x)) return x;
  if (r_isnan(y)) return y;
  const TR w = r_max(r_abs(x), r_abs(y)), z = r_min(r_abs(x), r_abs(y));
  if (z == rzero<TR>()) return w;
  const TR q = z / w;
  return w * r_sqrt(rone<TR>() + q * q);
}

// eigen-decomposition of [a b; b c]: rt1 (larger |.|), rt2, and the unit eigenvector (cs1, sn1) of rt1
template <class TR>
inline void sym2x2(TR a, TR b, TR c, TR& rt1, TR& rt2, TR& cs1, TR& sn1) {
  const TR zero = rzero<TR>(), one = rone<TR>(), half = Real<TR>::from(0.5);
  const TR sm = a + c, df = a - c, adf = r_abs(df), tb = b + b, ab = r_abs(tb);
  const bool a_big = r_abs(a) > r_abs(c);
  const TR acmx = a_big ? a : c, acmn = a_big ? c : a;
  TR rt;
  if (adf > ab) { const TR q = ab / adf; rt = adf * r_sqrt(one + q * q); }
  else if (adf < ab) { const TR q = adf / ab; rt = ab * r_sqrt(one + q * q); }
  else rt = ab * r_sqrt(Real<TR>::from(2.0));
  int sgn1;
  if (sm < zero) { rt1 = half * (sm - rt); sgn1 = -1; rt2 = (acmx / rt1) * acmn - (b / rt1) * b; }
  else if (sm > zero) { rt1 = half * (sm + rt); sgn1 = 1; rt2 = (acmx / rt1) * acmn - (b / rt1) * b; }
  else { rt1 = half * rt; rt2 = -(half * rt); sgn1 = 1; }
  int sgn2;
  TR cs;
  if (df >= zero) { cs = df + rt; sgn2 = 1; } else { cs = df - rt; sgn2 = -1; }
  if (r_abs(cs) > ab) { const TR ct = -tb / cs; sn1 = one / r_sqrt(one + ct * ct); cs1 = ct * sn1; }
  else if (ab == zero) { cs1 = one; sn1 = zero; }
  else { const TR tn = -cs / tb; cs1 = one / r_sqrt(one + tn * tn); sn1 = tn * cs1; }
  if (sgn1 == sgn2) { const TR tn = cs1; cs1 = -sn1; sn1 = tn; }
}

// dlascl-style safe rescaling of v by to/from
template <class TR>
inline void rescale(TR from, TR to, TR* v, int n) {
  if (from == to) return;
  const TR small = Real<TR>::floatmin(), big = rone<TR>() / small;
  TR cf = from, ct = to;
  bool done = false;
  while (!done) {
    TR mul;
    const TR cf1 = cf * small;
    if (cf1 == cf) { mul = ct / cf; done = true; }
    else {
      const TR ct1 = ct / big;
      if (ct1 == ct) { mul = ct; done = true; cf = rone<TR>(); }
      else if (r_abs(cf1) > r_abs(ct) && ct != rzero<TR>()) { mul = small; cf = cf1; }
      else if (r_abs(ct1) > r_abs(cf)) { mul = big; ct = ct1; }
      else { mul = ct / cf; done = true; }
    }
    for (int i = 0; i < n; ++i) v[i] = v[i] * mul;
  }
}

// Symmetric tridiagonal eigen-solver (implicit QL/QR, LAPACK dsteqr lineage as modified by
// ARPACK's dstqrb).  d[n], e[n-1] are overwritten (d <- ascending eigenvalues).
//   want = LastRow : z[n]      <- last row of the eigenvector matrix   (dstqrb!)
//   want = Full    : z[n*ldz]  <- eigenvector matrix, column-major      (simple_dsteqr!)
enum class Vecs { LastRow, Full };

template <class TR>
class TridiagEig {
 public:
  TridiagEig(int n, TR* d, TR* e, TR* z, int ldz, Vecs want) : n_(n), d_(d), e_(e), z_(z), ldz_(ldz), want_(want), cs_(n), sn_(n) {}

  void solve() {
    const TR zero = rzero<TR>(), one = rone<TR>();
    if (want_ == Vecs::LastRow) { for (int i = 0; i < n_; ++i) z_[i] = zero; z_[n_ - 1] = one; }
    else for (int j = 0; j < n_; ++j) for (int i = 0; i < n_; ++i) z_[i + (size_t)j * ldz_] = i == j ? one : zero;
    budget_ = HostConst<TR>::steqr_maxit() * n_;
    const TR epsm = Real<TR>::eps() * Real<TR>::from(0.5);
    int start = 0;
    while (start < n_) {
      if (start > 0) e_[start - 1] = zero;
      int stop = n_ - 1;  // block = [start, stop]
      for (int m = start; m < n_ - 1; ++m) {
        const TR t = r_abs(e_[m]);
        if (t == zero) { stop = m; break; }
        if (t <= r_sqrt(r_abs(d_[m])) * r_sqrt(r_abs(d_[m + 1])) * epsm) { e_[m] = zero; stop = m; break; }
      }
      if (stop > start) block(start, stop);
      start = stop + 1;
    }
    // selection sort into ascending order, carrying z along
    for (int i = 0; i + 1 < n_; ++i) {
      int k = i;
      TR p = d_[i];
      for (int j = i + 1; j < n_; ++j) if (d_[j] < p) { k = j; p = d_[j]; }
      if (k != i) {
        d_[k] = d_[i]; d_[i] = p;
        if (want_ == Vecs::LastRow) { TR t = z_[k]; z_[k] = z_[i]; z_[i] = t; }
        else for (int r = 0; r < n_; ++r) { TR t = z_[r + (size_t)k * ldz_]; z_[r + (size_t)k * ldz_] = z_[r + (size_t)i * ldz_]; z_[r + (size_t)i * ldz_] = t; }
      }
    }
  }

 private:
  int n_; TR *d_, *e_, *z_; int ldz_; Vecs want_;
  std::vector<TR> cs_, sn_;
  int budget_ = 0;

  // columns (j, j+1) of z <- rotated
  void rot(int j, TR c, TR s) {
    if (c == rone<TR>() && s == rzero<TR>()) return;
    if (want_ == Vecs::LastRow) {
      const TR t = z_[j + 1];
      z_[j + 1] = c * t - s * z_[j];
      z_[j] = s * t + c * z_[j];
    } else {
      TR* a = z_ + (size_t)j * ldz_;
      TR* b = a + ldz_;
      for (int i = 0; i < n_; ++i) { const TR t = b[i]; b[i] = c * t - s * a[i]; a[i] = s * t + c * a[i]; }
    }
  }

  void block(int lo, int hi) {
    const TR one = rone<TR>();
    const TR epsm = Real<TR>::eps() * Real<TR>::from(0.5), eps2 = epsm * epsm, safmin = Real<TR>::floatmin();
    const TR ssfmax = r_sqrt(one / safmin) / Real<TR>::from(3.0), ssfmin = r_sqrt(safmin) / eps2;
    const int nb = hi - lo + 1;
    TR* d = d_ + lo; TR* e = e_ + lo;
    TR anorm = r_max(r_abs(d[0]) + r_abs(e[0]), r_abs(d[nb - 1]) + r_abs(e[nb - 2]));
    for (int j = 1; j < nb - 1; ++j) anorm = r_max(anorm, r_abs(d[j]) + r_abs(e[j]) + r_abs(e[j - 1]));
    const TR target = anorm > ssfmax ? ssfmax : (anorm < ssfmin ? ssfmin : anorm);
    rescale(anorm, target, d, nb);
    rescale(anorm, target, e, nb - 1);
    if (r_abs(d[nb - 1]) < r_abs(d[0])) sweep_qr(lo, hi); else sweep_ql(lo, hi);
    rescale(target, anorm, d, nb);
    rescale(target, anorm, e, nb - 1);
  }

  bool negligible(TR off, TR a, TR b) const {
    const TR epsm = Real<TR>::eps() * Real<TR>::from(0.5), eps2 = epsm * epsm;
    const TR t = r_abs(off);
    return t * t <= (eps2 * r_abs(a)) * r_abs(b) + Real<TR>::floatmin();
  }

  void sweep_ql(int lo, int hi) {  // src/dstqrb.jl:793-924
    const TR zero = rzero<TR>(), one = rone<TR>(), two = Real<TR>::from(2.0);
    int l = lo;
    while (l < hi) {
      int m = hi;
      for (int q = l; q < hi; ++q) if (negligible(e_[q], d_[q], d_[q + 1])) { m = q; break; }
      if (m < hi) e_[m] = zero;
      if (m == l) { ++l; continue; }
      if (m == l + 1) {
        TR rt1, rt2, c, s;
        sym2x2(d_[l], e_[l], d_[l + 1], rt1, rt2, c, s);
        rot(l, c, s);
        d_[l] = rt1; d_[l + 1] = rt2; e_[l] = zero;
        l += 2;
        continue;
      }
      if (budget_ == 0) throw std::runtime_error("tridiagonal QL iteration did not converge");
      --budget_;
      TR p = d_[l];
      TR g = (d_[l + 1] - p) / (two * e_[l]);
      TR r = hypot2(g, one);
      g = d_[m] - p + (e_[l] / (g + r_copysign(r, g)));
      TR s = one, c = one;
      p = zero;
      for (int i = m - 1; i >= l; --i) {
        const TR f = s * e_[i], b = c * e_[i];
        givens(g, f, c, s, r);
        if (i != m - 1) e_[i + 1] = r;
        g = d_[i + 1] - p;
        r = (d_[i] - g) * s + two * c * b;
        p = s * r;
        d_[i + 1] = g + p;
        g = c * r - b;
        cs_[i] = c; sn_[i] = -s;
      }
      for (int i = m - 1; i >= l; --i) rot(i, cs_[i], sn_[i]);
      d_[l] = d_[l] - p;
      e_[l] = g;
    }
  }

  void sweep_qr(int lo, int hi) {  // src/dstqrb.jl:926-1052
    const TR zero = rzero<TR>(), one = rone<TR>(), two = Real<TR>::from(2.0);
    int l = hi;
    while (l > lo) {
      int m = lo;
      for (int q = l; q > lo; --q) if (negligible(e_[q - 1], d_[q], d_[q - 1])) { m = q; break; }
      if (m > lo) e_[m - 1] = zero;
      if (m == l) { --l; continue; }
      if (m == l - 1) {
        TR rt1, rt2, c, s;
        sym2x2(d_[l - 1], e_[l - 1], d_[l], rt1, rt2, c, s);
        rot(l - 1, c, s);
        d_[l - 1] = rt1; d_[l] = rt2; e_[l - 1] = zero;
        l -= 2;
        continue;
      }
      if (budget_ == 0) throw std::runtime_error("tridiagonal QR iteration did not converge");
      --budget_;
      TR p = d_[l];
      TR g = (d_[l - 1] - p) / (two * e_[l - 1]);
      TR r = hypot2(g, one);
      g = d_[m] - p + (e_[l - 1] / (g + r_copysign(r, g)));
      TR s = one, c = one;
      p = zero;
      for (int i = m; i <= l - 1; ++i) {
        const TR f = s * e_[i], b = c * e_[i];
        givens(g, f, c, s, r);
        if (i != m) e_[i - 1] = r;
        g = d_[i] - p;
        r = (d_[i + 1] - g) * s + two * c * b;
        p = s * r;
        d_[i] = g + p;
        g = c * r - b;
        cs_[i] = c; sn_[i] = s;
      }
      for (int i = m; i <= l - 1; ++i) rot(i, cs_[i], sn_[i]);
      d_[l] = d_[l] - p;
      e_[l - 1] = g;
    }
  }
};

}  // namespace host
}  // namespace ga
