// Whole-solve driver: symeigs = dsaupd! + simple_dseupd! (src/interface.jl:215-290) with the host
// control flow of dsaup2! (src/dsaup2.jl:962-1465, `idonow` form) driving the device entry points
// ga_getv0 / ga_lanczos / ga_apply_shifts / ga_ritz_vectors.  It exists because Julia is not
// available in the build image; with Julia the reference's own dsaupd!/dsaup2! make these calls.
#include <chrono>
#include <cstring>

#include "ga_ctx.cuh"
#include "host/irl.hpp"

namespace ga {

static double hnow() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct SolveBase {
  virtual ~SolveBase() {}
  virtual int iterate(int max_restarts) = 0;
  virtual int finish(int ritzvec, void* values, void* bounds, void* vectors_host, int64_t ldz, int64_t iparam_out[11]) = 0;
};

template <class TR>
struct Solve : SolveBase {
  ga_ctx* c;
  int which, nev_in, ncv, ishift = 1, mxiter;
  TR tol;
  // dsaupd workl partition (src/dsaupd.jl:1016-1023)
  std::vector<TR> H, ritz, bounds, Q, workl;
  // dsaup2 state (src/macros.jl:316-334)
  int nev, np, nev0, np0, kplusp, nconv = 0, iter = 0, info = 0;
  TR rnorm;
  int64_t iseed[4] = {1, 3, 5, 7};  // src/macros.jl:291
  bool started = false, finished = false;
  std::vector<TR> ritz_copy, bounds_copy;  // workl(kplusp+1:3kplusp) copies (src/dsaup2.jl:1133-1134)

  Solve(ga_ctx* c_, int which_, int nev_, TR tol_, int maxiter) : c(c_), which(which_), nev_in(nev_), ncv(c_->ncv), mxiter(maxiter), tol(tol_) {
    H.assign((size_t)2 * ncv, host::rzero<TR>());
    ritz.assign(ncv, host::rzero<TR>());
    bounds.assign(ncv, host::rzero<TR>());
    Q.assign((size_t)ncv * ncv, host::rzero<TR>());
    workl.assign((size_t)3 * ncv, host::rzero<TR>());
    ritz_copy.assign(ncv, host::rzero<TR>());
    bounds_copy.assign(ncv, host::rzero<TR>());
    if (!(tol > host::rzero<TR>())) tol = Real<TR>::eps() * Real<TR>::from(0.5);  // src/dsaupd.jl:1009-1011
    nev = nev0 = nev_in;
    np = np0 = ncv - nev_in;
    kplusp = ncv;
    rnorm = host::rzero<TR>();
  }

  static void to_pair(TR v, double out[2]);
  static TR from_pair(const double in[2]);

  int start(const void* v0_host) {
    int rc;
    if (v0_host) {                                             // src/interface.jl:259-262
      if ((rc = ga_upload_resid(c, v0_host))) return rc;
    }
    double rn[2] = {0, 0};
    rc = ga_getv0(c, iseed, v0_host ? 1 : 0, 1, rn);           // src/dsaup2.jl:1064
    if (rc <= GA_ERR_CUDA) return rc;
    rnorm = from_pair(rn);
    if (rnorm == host::rzero<TR>()) { info = -9; finished = true; return 0; }  // :1071-1076
    rc = ga_lanczos(c, 0, nev0, H.data(), ncv, rn, iseed);     // :1430
    if (rc <= GA_ERR_CUDA) return rc;
    rnorm = from_pair(rn);
    if (rc > 0) { np = rc; mxiter = iter; info = -9999; finished = true; return 0; }  // :1437-1448
    started = true;
    return 0;
  }

  // one pass of the dsaup2 main loop (label 1000 .. 130); returns 99 when the solve is finished
  int one_restart() {
    using namespace host;
    iter += 1;                                                 // :1084
    double rn[2];
    to_pair(rnorm, rn);
    int rc = ga_lanczos(c, nev, np, H.data(), ncv, rn, iseed); // :1097
    if (rc <= GA_ERR_CUDA) return rc;
    rnorm = from_pair(rn);
    if (rc > 0) { np = rc; mxiter = iter; info = -9999; finished = true; return 99; }  // :1104-1113
    double t0 = hnow();
    try {
      ritz_and_bounds<TR>(rnorm, kplusp, H.data(), ncv, ritz.data(), bounds.data(), workl.data());  // :1122
    } catch (const std::exception&) {
      info = GA_ERR_TRIDIAG; finished = true; return 99;       // :1125-1129
    }
    c->stats.teigt += hnow() - t0;
    for (int i = 0; i < kplusp; ++i) { ritz_copy[i] = ritz[i]; bounds_copy[i] = bounds[i]; }  // :1133-1134
    nev = nev0; np = np0;                                      // :1145-1146
    t0 = hnow();
    select_shifts<TR>(ishift, which, nev0, np0, ritz.data(), bounds.data(), workl.data());  // :1148
    c->stats.tgets += hnow() - t0;
    t0 = hnow();
    nconv = count_converged<TR>(nev0, ritz.data() + np0, bounds.data() + np0, tol);  // :1155-1156
    c->stats.tconv += hnow() - t0;
    {
      const int nptemp = np;                                   // :1177-1183
      for (int j = 0; j < nptemp; ++j)
        if (bounds[j] == rzero<TR>()) { np -= 1; nev += 1; }
    }
    if (nconv >= nev0 || iter > mxiter || np == 0) {           // :1185
      const TR e23 = HostConst<TR>::eps23();
      if (which == GA_BE) {                                    // :1194-1211
        sort_pair<TR>(GA_SA, kplusp, ritz.data(), bounds.data());
        const int nevd2 = nev0 / 2, nevm2 = nev0 - nevd2;
        if (nev > 1) {
          np = kplusp - nev0;
          const int cnt = std::min(nevd2, np);
          const int off = std::max(kplusp - nevd2 + 1, kplusp - np + 1) - 1, ioff = nevm2;
          for (int i = 0; i < cnt; ++i) { std::swap(ritz[off + i], ritz[ioff + i]); std::swap(bounds[off + i], bounds[ioff + i]); }
        }
      } else {                                                 // :1221-1232
        const int wprime = which == GA_LM ? GA_SM : which == GA_SM ? GA_LM : which == GA_LA ? GA_SA : GA_LA;
        sort_pair<TR>(wprime, kplusp, ritz.data(), bounds.data());
      }
      for (int j = 0; j < nev0; ++j) bounds[j] = bounds[j] / r_max(e23, r_abs(ritz[j]));  // :1237-1240
      sort_pair<TR>(GA_LA, nev0, bounds.data(), ritz.data());                              // :1248-1249
      for (int j = 0; j < nev0; ++j) bounds[j] = bounds[j] * r_max(e23, r_abs(ritz[j]));  // :1253-1256
      if (which == GA_BE) sort_pair<TR>(GA_LA, nconv, ritz.data(), bounds.data());         // :1264-1276
      else sort_pair<TR>(which, nconv, ritz.data(), bounds.data());
      H[0] = rnorm;                                            // :1280
      if (iter > mxiter && nconv < nev) info = 1;              // :1288-1290
      if (np == 0 && nconv < nev0) info = 2;                   // :1293-1295
      np = nconv;                                              // :1297
      mxiter = iter; nev = nconv;                              // :1455-1458
      finished = true;
      return 99;
    } else if (nconv < nev && ishift == 1) {                   // :1300-1322
      const int nevbef = nev;
      nev = nev + std::min(nconv, np / 2);
      if (nev == 1 && kplusp >= 6) nev = kplusp / 2;
      else if (nev == 1 && kplusp > 2) nev = 2;
      np = kplusp - nev;
      if (nevbef < nev) select_shifts<TR>(ishift, which, nev, np, ritz.data(), bounds.data(), workl.data());
    }
    // dsapps: host bulge chase, device basis update + residual norm (:1374-1414)
    t0 = hnow();
    chase_shifts<TR>(nev, np, ritz.data(), H.data(), ncv, Q.data(), ncv);
    c->stats.tapps += hnow() - t0;
    double beta[2], rout[2] = {0, 0};
    to_pair(H[nev], beta);                                     // H[kev+1,1]
    rc = ga_apply_shifts(c, nev, np, Q.data(), ncv, beta, rout);
    if (rc) return rc;
    rnorm = from_pair(rout);
    return 0;
  }

  int iterate(int max_restarts) override {
    if (finished) return 99;
    if (!started) return GA_ERR_ARG;
    for (int r = 0; max_restarts < 0 || r < max_restarts; ++r) {
      int rc = one_restart();
      if (rc != 0) return rc;
    }
    return 0;
  }

  // simple_dseupd! for mode 1 (src/dseupd.jl:986-1532) + the tail of dsaupd! (src/dsaupd.jl:1079-1094)
  int finish(int ritzvec, void* values, void* bounds_out, void* vectors_host, int64_t ldz, int64_t iparam_out[11]) override {
    using namespace host;
    int aupd_info = info;
    if (aupd_info == 2) aupd_info = 3;                         // src/dsaupd.jl:1092-1094
    if (iparam_out) {
      for (int i = 0; i < 11; ++i) iparam_out[i] = 0;
      iparam_out[0] = ishift; iparam_out[3] = 1; iparam_out[6] = c->op_kind == 3 ? 2 : 1;   // mode
      iparam_out[2] = mxiter;                                  // :1081
      iparam_out[4] = finished ? np : 0;                       // :1082 (nconv)
      iparam_out[8] = c->stats.nopx; iparam_out[9] = c->stats.nbx; iparam_out[10] = c->stats.nrorth;
    }
    if (!finished) return GA_ERR_ARG;
    if (aupd_info != 0 && aupd_info != 1) return aupd_info;    // src/interface.jl:269-273 raises
    const int nc = nconv;
    if (nc == 0) return aupd_info;                             // src/dseupd.jl:1029-1032
    TR* d = (TR*)values;
    std::vector<TR> hd(ncv), hb(ncv), Qe((size_t)ncv * ncv), tau(ncv), lastrow(ncv);
    const TR rn = H[0];                                        // :1175
    const TR e23 = HostConst<TR>::eps23();
    if (ritzvec) {
      // wanted Ritz values: re-select on the dseigt copies (:1195-1254)
      std::vector<TR> irz(ritz_copy), idx(ncv);
      std::vector<int> select(ncv, 0);
      for (int j = 0; j < ncv; ++j) idx[j] = Real<TR>::from((double)(j + 1));
      select_shifts<TR>(0, which, nev_in, ncv - nev_in, irz.data(), idx.data(), workl.data());
      bool reord = false;
      int numcnv = 0;
      for (int j = 1; j <= ncv; ++j) {
        const TR temp1 = r_max(e23, r_abs(irz[ncv - j]));
        const int jj = (int)r_hi(idx[ncv - j]);
        if (numcnv < nc && bounds_copy[jj - 1] <= tol * temp1) {
          select[jj - 1] = 1;
          ++numcnv;
          if (jj > nc) reord = true;
        }
      }
      if (numcnv != nc) return -17;                            // :1251-1254
      for (int i = 0; i + 1 < ncv; ++i) hb[i] = H[i + 1];      // :1261-1262
      for (int i = 0; i < ncv; ++i) hd[i] = H[ncv + i];
      try {
        TridiagEig<TR> te(ncv, hd.data(), hb.data(), Qe.data(), ncv, Vecs::Full);  // :1264
        te.solve();
      } catch (const std::exception&) { return GA_ERR_TRIDIAG; }
      if (reord) {                                             // :1282-1320
        int left = 0, right = ncv - 1;
        while (left < right) {
          if (select[left]) ++left;
          else if (!select[right]) --right;
          else {
            std::swap(hd[left], hd[right]);
            for (int r = 0; r < ncv; ++r) std::swap(Qe[r + (size_t)ncv * left], Qe[r + (size_t)ncv * right]);
            ++left; --right;
          }
        }
      }
      for (int i = 0; i < nc; ++i) d[i] = hd[i];               // :1325
      // ascending sort of the wanted values with their tridiagonal eigenvectors (:1346)
      shell_sort<TR>(GA_LA, nc, d, [&](int a, int b) { for (int r = 0; r < ncv; ++r) std::swap(Qe[r + (size_t)ncv * a], Qe[r + (size_t)ncv * b]); });
      // QR of the wanted eigenvectors, then Qh = H(1)...H(nc) formed explicitly (:1419-1431)
      householder_qr<TR>(ncv, nc, Qe.data(), ncv, tau.data());
      std::vector<TR> Qh((size_t)ncv * ncv, rzero<TR>());
      for (int i = 0; i < ncv; ++i) Qh[i + (size_t)i * ncv] = rone<TR>();
      apply_q_right<TR>(ncv, ncv, nc, Qe.data(), ncv, tau.data(), Qh.data(), ncv);
      int rc = ga_ritz_vectors(c, nc, Qh.data(), ncv, vectors_host, ldz);
      if (rc) return rc;
      // Ritz estimates: rnorm * |last row of the eigenvector matrix| (:1441-1471)
      for (int j = 0; j < ncv; ++j) lastrow[j] = rzero<TR>();
      lastrow[ncv - 1] = rone<TR>();
      apply_qt_left_vec<TR>(ncv, nc, Qe.data(), ncv, tau.data(), lastrow.data());
      if (bounds_out) for (int j = 0; j < nc; ++j) ((TR*)bounds_out)[j] = rn * r_abs(lastrow[j]);
    } else {
      for (int i = 0; i < nc; ++i) d[i] = ritz[i];             // :1328
      if (bounds_out) for (int i = 0; i < nc; ++i) ((TR*)bounds_out)[i] = bounds[i];  // :1348
    }
    return aupd_info;
  }
};

template <> void Solve<double>::to_pair(double v, double out[2]) { out[0] = v; out[1] = 0.0; }
template <> double Solve<double>::from_pair(const double in[2]) { return in[0]; }
template <> void Solve<ddbl>::to_pair(ddbl v, double out[2]) { out[0] = v.hi; out[1] = v.lo; }
template <> ddbl Solve<ddbl>::from_pair(const double in[2]) { return ddbl(in[0], in[1]); }

void solve_free(ga_ctx* c) {
  if (c->solve) { delete (SolveBase*)c->solve; c->solve = nullptr; }
}

}  // namespace ga

using namespace ga;

extern "C" {

int ga_symeigs_begin(ga_ctx* c, int which, int nev, const double* tol, int maxiter, const void* v0_host) {
  if (!c) return GA_ERR_ARG;
  solve_free(c);
  // argument checks of dsaupd! (src/dsaupd.jl:963-1002)
  // (like the reference, a later failing check overrides an earlier one; -6/-7/-10/-11/-12 cannot be
  // reached through this ABI: bmat and mode come from the uploaded operator, workl is internal, ishift = 1)
  int ierr = 0;
  if (c->n_global <= 0) ierr = -1;
  else if (nev <= 0) ierr = -2;
  else if (c->ncv <= nev || (i64)c->ncv > c->n_global) ierr = -3;
  if (maxiter < 0) ierr = -4;                      // < 0, not <= 0: mxiter = 0 still runs one iteration (:971-976)
  if (which < GA_LM || which > GA_BE) ierr = -5;
  if (nev == 1 && which == GA_BE) ierr = -13;
  if (ierr != 0) return ierr;
  if (c->op_kind == 0) return set_err(c, GA_ERR_NOOP, "no operator uploaded (ga_set_csr)");
  const double t0 = hnow();
  int rc;
  if (c->dtype == GA_DD) {
    auto* s = new Solve<ddbl>(c, which, nev, tol ? ddbl(tol[0], tol[1]) : ddbl(0.0), maxiter);
    c->solve = s;
    rc = s->start(v0_host);
  } else {
    double t = tol ? tol[0] : 0.0;
    // Float32 vectors with a Float64 factorisation: the lowest precision sets the default tolerance
    // (src/interface.jl:33-34, docs of `tol`): eps(Float32)/2
    if (c->dtype == GA_F32 && !(t > 0.0)) t = 5.9604644775390625e-08;
    auto* s = new Solve<double>(c, which, nev, t, maxiter);
    c->solve = s;
    rc = s->start(v0_host);
  }
  c->stats.taup2 += hnow() - t0;
  return rc;
}

int ga_symeigs_iterate(ga_ctx* c, int max_restarts) {
  if (!c || !c->solve) return GA_ERR_ARG;
  const double t0 = hnow();
  int rc = ((SolveBase*)c->solve)->iterate(max_restarts);
  c->stats.taup2 += hnow() - t0;
  return rc;
}

int ga_symeigs_end(ga_ctx* c, int ritzvec, void* values, void* bounds, void* vectors_host, int64_t ldz,
                   int64_t iparam_out[11]) {
  if (!c || !c->solve) return GA_ERR_ARG;
  int rc = ((SolveBase*)c->solve)->finish(ritzvec, values, bounds, vectors_host, ldz, iparam_out);
  return rc;
}

// ---- host-only pieces of the control flow (csrc/host/*.hpp), exported so that CPU tests can hold them against the
// oracle and the reference's analytic known answers without a device.  use_dd != 0: reals are (hi, lo) pairs.
#define GA_HOST_DISPATCH(expr_double, expr_dd) \
  try {                                        \
    if (use_dd) { expr_dd; } else { expr_double; } \
  } catch (const std::exception&) {            \
    return GA_ERR_TRIDIAG;                     \
  }                                            \
  return 0

int ga_host_tridiag_eig(int use_dd, int n, void* d, void* e, void* z, int want_full, int ldz) {
  if (n < 1 || !d || !z || (n > 1 && !e) || (want_full && ldz < n)) return GA_ERR_ARG;
  const host::Vecs w = want_full ? host::Vecs::Full : host::Vecs::LastRow;
  GA_HOST_DISPATCH((host::TridiagEig<double>(n, (double*)d, (double*)e, (double*)z, want_full ? ldz : 1, w).solve()),
                   (host::TridiagEig<ddbl>(n, (ddbl*)d, (ddbl*)e, (ddbl*)z, want_full ? ldz : 1, w).solve()));
}
int ga_host_ritz_and_bounds(int use_dd, const void* rnorm, int n, const void* H, int ldh, void* eig, void* bounds, void* work) {
  if (n < 1 || !rnorm || !H || !eig || !bounds || !work || ldh < n) return GA_ERR_ARG;
  GA_HOST_DISPATCH((host::ritz_and_bounds<double>(*(const double*)rnorm, n, (const double*)H, ldh, (double*)eig, (double*)bounds, (double*)work)),
                   (host::ritz_and_bounds<ddbl>(*(const ddbl*)rnorm, n, (const ddbl*)H, ldh, (ddbl*)eig, (ddbl*)bounds, (ddbl*)work)));
}
int ga_host_sort_pair(int use_dd, int which, int n, void* x1, void* x2) {
  if (n < 0 || !x1 || !x2 || which < GA_LM || which > GA_SA) return GA_ERR_ARG;
  GA_HOST_DISPATCH((host::sort_pair<double>(which, n, (double*)x1, (double*)x2)), (host::sort_pair<ddbl>(which, n, (ddbl*)x1, (ddbl*)x2)));
}
int ga_host_select_shifts(int use_dd, int ishift, int which, int kev, int np, void* ritz, void* bounds, void* shifts) {
  if (kev < 0 || np < 0 || !ritz || !bounds || !shifts || which < GA_LM || which > GA_BE) return GA_ERR_ARG;
  GA_HOST_DISPATCH((host::select_shifts<double>(ishift, which, kev, np, (double*)ritz, (double*)bounds, (double*)shifts)),
                   (host::select_shifts<ddbl>(ishift, which, kev, np, (ddbl*)ritz, (ddbl*)bounds, (ddbl*)shifts)));
}
int ga_host_count_converged(int use_dd, int n, const void* ritz, const void* bounds, const void* tol) {
  if (n < 0 || !ritz || !bounds || !tol) return GA_ERR_ARG;
  if (use_dd) return host::count_converged<ddbl>(n, (const ddbl*)ritz, (const ddbl*)bounds, *(const ddbl*)tol);
  return host::count_converged<double>(n, (const double*)ritz, (const double*)bounds, *(const double*)tol);
}
int ga_host_chase_shifts(int use_dd, int kev, int np, const void* shift, void* H, int ldh, void* Q, int ldq) {
  if (kev < 1 || np < 0 || !shift || !H || !Q || ldh < kev + np || ldq < kev + np) return GA_ERR_ARG;
  GA_HOST_DISPATCH((host::chase_shifts<double>(kev, np, (const double*)shift, (double*)H, ldh, (double*)Q, ldq)),
                   (host::chase_shifts<ddbl>(kev, np, (const ddbl*)shift, (ddbl*)H, ldh, (ddbl*)Q, ldq)));
}

int ga_symeigs(ga_ctx* c, int which, int nev, const double* tol, int maxiter, const void* v0_host, int ritzvec,
               void* values, void* bounds, void* vectors_host, int64_t ldz, int64_t iparam_out[11]) {
  const double t0 = hnow();
  int rc = ga_symeigs_begin(c, which, nev, tol, maxiter, v0_host);
  if (rc != 0) return rc;
  rc = ga_symeigs_iterate(c, -1);
  if (rc != 99) return rc == 0 ? GA_ERR_ARG : rc;
  c->stats.taupd += hnow() - t0;
  return ga_symeigs_end(c, ritzvec, values, bounds, vectors_host, ldz, iparam_out);
}

}  // extern "C"
