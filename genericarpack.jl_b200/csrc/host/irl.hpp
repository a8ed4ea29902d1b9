// Host control of the implicitly restarted Lanczos iteration: the O(ncv)/O(ncv^2) scalar work
// that BASELINE.json's north_star keeps on the host.  With Julia available this is the
// reference's own code (dsaup2!, dsgets, dsconv, dseigt!, dsapps!'s bulge chase, simple_dseupd!'s
// small eigenproblem) calling the C ABI; here it is mirrored in C++ for TR in {double, ddbl}.
// Citations relative to /root/reference.
#pragma once
#include <algorithm>
#include <vector>

#include "tridiag.hpp"

namespace ga {
namespace host {

// ---- dsortr (src/simple.jl:255-308): shell sort of x1 in the order OPPOSITE to `which`
// (so that the wanted end lands in the LAST positions), optionally applied to a companion.
template <class TR>
inline bool out_of_order(int which, TR a, TR b) {
  switch (which) {
    case GA_SA: return a < b;
    case GA_SM: return r_abs(a) < r_abs(b);
    case GA_LA: return a > b;
    case GA_LM: return r_abs(a) > r_abs(b);
  }
  return false;
}
template <class TR, class Swap>
inline void shell_sort(int which, int n, TR* x1, Swap&& swap_companion) {
  for (int gap = n / 2; gap != 0; gap /= 2)
    for (int i = gap; i < n; ++i)
      for (int j = i - gap; j >= 0 && out_of_order(which, x1[j], x1[j + gap]); j -= gap) {
        std::swap(x1[j], x1[j + gap]);
        swap_companion(j, j + gap);
      }
}
template <class TR>
inline void sort_pair(int which, int n, TR* x1, TR* x2) {
  shell_sort(which, n, x1, [&](int a, int b) { std::swap(x2[a], x2[b]); });
}

// ---- dsgets (src/simple.jl:603-702)
template <class TR>
inline void select_shifts(int ishift, int which, int kev, int np, TR* ritz, TR* bounds, TR* shifts) {
  if (which == GA_BE) {
    sort_pair(GA_LA, kev + np, ritz, bounds);
    const int kevd2 = kev / 2;
    if (kev > 1) {
      const int cnt = std::min(kevd2, np), off = std::max(kevd2, np);
      for (int i = 0; i < cnt; ++i) { std::swap(ritz[i], ritz[off + i]); std::swap(bounds[i], bounds[off + i]); }
    }
  } else {
    sort_pair(which, kev + np, ritz, bounds);
  }
  if (ishift == 1 && np > 0) {
    sort_pair(GA_SM, np, bounds, ritz);
    for (int i = 0; i < np; ++i) shifts[i] = ritz[i];
  }
}

// ---- dsconv (src/simple.jl:37-80)
template <class TR>
inline int count_converged(int n, const TR* ritz, const TR* bounds, TR tol) {
  const TR e23 = HostConst<TR>::eps23();
  int nconv = 0;
  for (int i = 0; i < n; ++i)
    if (bounds[i] <= tol * r_max(e23, r_abs(ritz[i]))) ++nconv;
  return nconv;
}

// ---- dseigt! (src/simple.jl:968-1035): Ritz values of the tridiagonal H and their error bounds
template <class TR>
inline void ritz_and_bounds(TR rnorm, int n, const TR* H, int ldh, TR* eig, TR* bounds, TR* work) {
  for (int i = 0; i < n; ++i) eig[i] = H[ldh + i];
  for (int i = 0; i + 1 < n; ++i) work[i] = H[i + 1];
  TridiagEig<TR> te(n, eig, work, bounds, 1, Vecs::LastRow);
  te.solve();
  for (int i = 0; i < n; ++i) bounds[i] = rnorm * r_abs(bounds[i]);
}

// ---- host part of dsapps! (src/dsapps.jl:609-805): np implicit shifts by bulge chasing on the
// tridiagonal H (ldh x 2), accumulating the rotations in Q (ldq x kplusp).
template <class TR>
inline void chase_shifts(int kev, int np, const TR* shift, TR* H, int ldh, TR* Q, int ldq) {
  const TR zero = rzero<TR>(), one = rone<TR>();
  const TR epsmch = Real<TR>::eps() * Real<TR>::from(0.5);
  const int kplusp = kev + np;
  auto sub = [&](int i) -> TR& { return H[i]; };          // H[i+1,1] for 0-based row i
  auto dia = [&](int i) -> TR& { return H[ldh + i]; };    // H[i+1,2]
  auto q = [&](int i, int j) -> TR& { return Q[i + (size_t)j * ldq]; };
  for (int j = 0; j < kplusp; ++j) for (int i = 0; i < kplusp; ++i) q(i, j) = i == j ? one : zero;
  int itop = 0;
  for (int jj = 0; jj < np; ++jj) {
    int istart = itop;
    while (istart < kplusp) {
      int iend = kplusp - 1;
      for (int i = istart; i < kplusp - 1; ++i) {
        const TR big = r_abs(dia(i)) + r_abs(dia(i + 1));
        if (sub(i + 1) <= epsmch * big) { sub(i + 1) = zero; iend = i; break; }
      }
      if (istart < iend) {
        TR f = dia(istart) - shift[jj], g = sub(istart + 1), c, s, r;
        givens(f, g, c, s, r);
        TR a1 = c * dia(istart) + s * sub(istart + 1);
        TR a2 = c * sub(istart + 1) + s * dia(istart + 1);
        TR a4 = c * dia(istart + 1) - s * sub(istart + 1);
        TR a3 = c * sub(istart + 1) - s * dia(istart);
        dia(istart) = c * a1 + s * a2;
        dia(istart + 1) = c * a4 - s * a3;
        sub(istart + 1) = c * a3 + s * a4;
        const int rows0 = std::min(istart + jj + 2, kplusp);  // min(istart+jj, kplusp) in 1-based terms
        for (int j = 0; j < rows0; ++j) {
          a1 = c * q(j, istart) + s * q(j, istart + 1);
          q(j, istart + 1) = c * q(j, istart + 1) - s * q(j, istart);
          q(j, istart) = a1;
        }
        for (int i = istart + 1; i <= iend - 1; ++i) {
          f = sub(i);
          g = s * sub(i + 1);
          sub(i + 1) = c * sub(i + 1);
          givens(f, g, c, s, r);
          if (r < zero) { r = -r; c = -c; s = -s; }
          sub(i) = r;
          a1 = c * dia(i) + s * sub(i + 1);
          a2 = c * sub(i + 1) + s * dia(i + 1);
          a3 = c * sub(i + 1) - s * dia(i);
          a4 = c * dia(i + 1) - s * sub(i + 1);
          dia(i) = c * a1 + s * a2;
          dia(i + 1) = c * a4 - s * a3;
          sub(i + 1) = c * a3 + s * a4;
          const int rows = std::min(i + jj + 2, kplusp);
          for (int j = 0; j < rows; ++j) {
            a1 = c * q(j, i) + s * q(j, i + 1);
            q(j, i + 1) = c * q(j, i + 1) - s * q(j, i);
            q(j, i) = a1;
          }
        }
      }
      if (sub(iend) < zero) {
        sub(iend) = -sub(iend);
        for (int j = 0; j < kplusp; ++j) q(j, iend) = -q(j, iend);
      }
      istart = iend + 1;
    }
    for (int i = itop; i < kplusp - 1; ++i) {
      if (sub(i + 1) > zero) break;
      ++itop;
    }
  }
  for (int i = itop; i < kplusp - 1; ++i) {
    const TR big = r_abs(dia(i)) + r_abs(dia(i + 1));
    if (sub(i + 1) <= epsmch * big) sub(i + 1) = zero;
  }
}

// ---- Householder pieces used by simple_dseupd! on the ncv x nconv matrix of wanted tridiagonal
// eigenvectors (src/arpack-blas-qr.jl:48-185; norm2 of a short real vector -> plain scaled norm).
template <class TR>
inline TR small_norm2(const TR* x, int n) {
  if (n == 1) return r_abs(x[0]);
  TR scale = rzero<TR>(), ssq = rone<TR>();
  for (int i = 0; i < n; ++i) {
    if (x[i] != rzero<TR>()) {
      const TR a = r_abs(x[i]);
      if (scale < a) { const TR v = scale / a; ssq = rone<TR>() + ssq * v * v; scale = a; }
      else { const TR v = a / scale; ssq = ssq + v * v; }
    }
  }
  return scale * r_sqrt(ssq);
}
template <class TR>
inline TR make_reflector(TR& alpha, TR* x, int len) {  // _dlarfg!
  if (len <= 0) return rzero<TR>();
  TR xnorm = small_norm2(x, len);
  if (xnorm == rzero<TR>()) return rzero<TR>();
  TR beta = -r_copysign(hypot2(alpha, xnorm), alpha);
  const TR safmin = Real<TR>::floatmin() / (Real<TR>::eps() * Real<TR>::from(0.5));
  int knt = 0;
  if (r_abs(beta) < safmin) {
    const TR rsafmn = rone<TR>() / safmin;
    while (knt < 20) {
      ++knt;
      for (int i = 0; i < len; ++i) x[i] = x[i] * rsafmn;
      beta = beta * rsafmn; alpha = alpha * rsafmn;
      if (r_abs(beta) >= safmin) break;
    }
    xnorm = small_norm2(x, len);
    beta = -r_copysign(hypot2(alpha, xnorm), alpha);
  }
  const TR tau = (beta - alpha) / beta;
  const TR sc = rone<TR>() / (alpha - beta);
  for (int i = 0; i < len; ++i) x[i] = x[i] * sc;
  for (int i = 0; i < knt; ++i) beta = beta * safmin;
  alpha = beta;
  return tau;
}
// A (m x ncol, lda) <- Householder QR factors; tau[ncol]
template <class TR>
inline void householder_qr(int m, int ncol, TR* A, int lda, TR* tau) {
  std::vector<TR> w(ncol);
  const int k = std::min(m, ncol);
  for (int i = 0; i < k; ++i) {
    TR* v = A + i + (size_t)i * lda;
    tau[i] = make_reflector(v[0], v + 1, m - i - 1);
    if (i + 1 < ncol && tau[i] != rzero<TR>()) {
      const TR aii = v[0];
      v[0] = rone<TR>();
      for (int c = i + 1; c < ncol; ++c) {
        const TR* C = A + i + (size_t)c * lda;
        TR s = rzero<TR>();
        for (int r = 0; r < m - i; ++r) s = s + C[r] * v[r];
        w[c] = s;
      }
      for (int c = i + 1; c < ncol; ++c) {
        if (w[c] == rzero<TR>()) continue;
        TR* C = A + i + (size_t)c * lda;
        const TR t = -(tau[i] * w[c]);
        for (int r = 0; r < m - i; ++r) C[r] = C[r] + v[r] * t;
      }
      v[0] = aii;
    }
  }
}
// C (rows x m, ldc) <- C * H(1) H(2) ... H(k), reflectors stored in A (m x k, lda)   [dorm2r R,N]
template <class TR>
inline void apply_q_right(int rows, int m, int k, TR* A, int lda, const TR* tau, TR* C, int ldc) {
  std::vector<TR> w(rows);
  for (int i = 0; i < k; ++i) {
    if (tau[i] == rzero<TR>()) continue;
    TR* v = A + i + (size_t)i * lda;
    const TR aii = v[0];
    v[0] = rone<TR>();
    const int len = m - i;
    for (int r = 0; r < rows; ++r) {
      TR s = rzero<TR>();
      for (int c = 0; c < len; ++c) s = s + C[r + (size_t)(i + c) * ldc] * v[c];
      w[r] = s;
    }
    for (int c = 0; c < len; ++c) {
      if (v[c] == rzero<TR>()) continue;
      const TR t = -(tau[i] * v[c]);
      for (int r = 0; r < rows; ++r) C[r + (size_t)(i + c) * ldc] = C[r + (size_t)(i + c) * ldc] + w[r] * t;
    }
    v[0] = aii;
  }
}
// c (m) <- H(k)^T ... H(1)^T c   [dorm2r L,T on one column]
template <class TR>
inline void apply_qt_left_vec(int m, int k, TR* A, int lda, const TR* tau, TR* c) {
  for (int i = 0; i < k; ++i) {
    if (tau[i] == rzero<TR>()) continue;
    TR* v = A + i + (size_t)i * lda;
    const TR aii = v[0];
    v[0] = rone<TR>();
    TR s = rzero<TR>();
    for (int r = 0; r < m - i; ++r) s = s + c[i + r] * v[r];
    if (s != rzero<TR>()) {
      const TR t = -(tau[i] * s);
      for (int r = 0; r < m - i; ++r) c[i + r] = c[i + r] + v[r] * t;
    }
    v[0] = aii;
  }
}

}  // namespace host
}  // namespace ga
