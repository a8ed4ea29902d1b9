// TEST INFRASTRUCTURE ONLY -- C entry points of the CPU oracle, loaded with ctypes by
// oracle/oracle.py.  See oracle/oracle.hpp for what is restated and from where.
#include "oracle.hpp"
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace oracle;

namespace {
struct HandleBase {
  int dtype;
  virtual ~HandleBase() {}
};
template <class T>
struct Handle : HandleBase {
  Op<T>* op = nullptr;
  Problem<T>* P = nullptr;
  ~Handle() override { delete P; delete op; }
};
template <class TR> TR make_real(const double* p);
template <> double make_real<double>(const double* p) { return p[0]; }
template <> dd make_real<dd>(const double* p) { return dd(p[0], p[1]); }
}  // namespace

#define DISPATCH(h, CALL)                                              \
  switch (((HandleBase*)(h))->dtype) {                                 \
    case 0: { using T = double; auto* H_ = (Handle<T>*)(h); CALL; } break; \
    case 1: { using T = cplx; auto* H_ = (Handle<T>*)(h); CALL; } break;   \
    case 2: { using T = dd; auto* H_ = (Handle<T>*)(h); CALL; } break;     \
    case 3: { using T = float; auto* H_ = (Handle<T>*)(h); CALL; } break;  \
    default: break;                                                    \
  }

extern "C" {

// dtype: 0 = Float64, 1 = ComplexF64, 2 = Double64 (hi,lo pairs).
// m_rows == 0: ArpackSimpleOp on an n x n CSR matrix; m_rows > 0: ArpackNormalOp on an
// m_rows x n_cols matrix (problem size = min side).
void* oracle_create(int dtype, long long n, int ncv, const long long* rowptr, const int* col, const void* val,
                    long long m_rows, long long n_cols) {
  try {
    switch (dtype) {
#define MK(CODE, TT)                                                                         \
  case CODE: {                                                                               \
    auto* h = new Handle<TT>();                                                              \
    h->dtype = CODE;                                                                         \
    if (m_rows > 0) h->op = new NormalOp<TT>(m_rows, n_cols, rowptr, col, (const TT*)val);   \
    else h->op = new CsrOp<TT>(n, rowptr, col, (const TT*)val);                              \
    h->P = new Problem<TT>(h->op->n, ncv, h->op);                                            \
    return h;                                                                                \
  }
      MK(0, double)
      MK(1, cplx)
      MK(2, dd)
      MK(3, float)
#undef MK
    }
  } catch (...) {
  }
  return nullptr;
}
// ArpackSymmetricGeneralizedOp(A, factorize(B), B): mode 2, bmat = :G; A, B n x n full CSR, B positive definite.
// Returns NULL if B is not positive definite.
void* oracle_create_general(int dtype, long long n, int ncv, const long long* arp, const int* acol, const void* aval,
                            const long long* brp, const int* bcol, const void* bval) {
  try {
    switch (dtype) {
#define MK(CODE, TT)                                                                                         \
  case CODE: {                                                                                               \
    auto* op = new GeneralizedOp<TT>(n, arp, acol, (const TT*)aval, brp, bcol, (const TT*)bval);             \
    if (!op->spd) { delete op; return nullptr; }                                                             \
    auto* h = new Handle<TT>();                                                                              \
    h->dtype = CODE;                                                                                         \
    h->op = op;                                                                                              \
    h->P = new Problem<TT>(n, ncv, h->op);                                                                   \
    return h;                                                                                                \
  }
      MK(0, double)
      MK(1, cplx)
      MK(2, dd)
#undef MK
    }
  } catch (...) {
  }
  return nullptr;
}
void oracle_destroy(void* h) { delete (HandleBase*)h; }
void oracle_bx(void* h, void* y, const void* x) { DISPATCH(h, H_->op->bx((T*)y, (const T*)x)); }

long long oracle_size(void* h) { long long r = 0; DISPATCH(h, r = H_->P->n); return r; }
// what: 0 V, 1 resid, 2 workd, 3 workl
void* oracle_ptr(void* h, int what) {
  void* r = nullptr;
  DISPATCH(h, {
    switch (what) {
      case 0: r = H_->P->V.data(); break;
      case 1: r = H_->P->resid.data(); break;
      case 2: r = H_->P->workd.data(); break;
      case 3: r = H_->P->workl.data(); break;
    }
  });
  return r;
}
long long* oracle_iseed(void* h) { long long* r = nullptr; DISPATCH(h, r = H_->P->iseed); return r; }
long long* oracle_iparam(void* h) { long long* r = nullptr; DISPATCH(h, r = H_->P->iparam); return r; }
// rnorm is returned/accepted as two doubles (hi, lo); lo = 0 except for Double64.
void oracle_get_rnorm(void* h, double* out) {
  switch (((HandleBase*)h)->dtype) {
    case 0: out[0] = ((Handle<double>*)h)->P->rnorm; out[1] = 0; break;
    case 1: out[0] = ((Handle<cplx>*)h)->P->rnorm; out[1] = 0; break;
    case 3: out[0] = ((Handle<float>*)h)->P->rnorm; out[1] = 0; break;
    case 2: out[0] = ((Handle<dd>*)h)->P->rnorm.hi; out[1] = ((Handle<dd>*)h)->P->rnorm.lo; break;
  }
}
void oracle_set_rnorm(void* h, const double* in) {
  switch (((HandleBase*)h)->dtype) {
    case 0: ((Handle<double>*)h)->P->rnorm = in[0]; break;
    case 1: ((Handle<cplx>*)h)->P->rnorm = in[0]; break;
    case 3: ((Handle<float>*)h)->P->rnorm = in[0]; break;
    case 2: ((Handle<dd>*)h)->P->rnorm = dd(in[0], in[1]); break;
  }
}
// stats: nopx nbx nrorth nitref nrstrt | taupd taup2 taitr teigt tgets tapps tconv tmvopx tgetv0 titref trvec
void oracle_get_stats(void* h, long long* counts, double* times) {
  DISPATCH(h, {
    const Stats& s = H_->P->stats;
    counts[0] = s.nopx; counts[1] = s.nbx; counts[2] = s.nrorth; counts[3] = s.nitref; counts[4] = s.nrstrt;
    times[0] = s.taupd; times[1] = s.taup2; times[2] = s.taitr; times[3] = s.teigt; times[4] = s.tgets;
    times[5] = s.tapps; times[6] = s.tconv; times[7] = s.tmvopx; times[8] = s.tgetv0; times[9] = s.titref;
    times[10] = s.trvec;
  });
}
void oracle_reset_stats(void* h) { DISPATCH(h, H_->P->stats = Stats()); }

void oracle_opx(void* h, void* y, const void* x) { DISPATCH(h, H_->op->opx((T*)y, (const T*)x)); }

int oracle_dgetv0(void* h, int itry, int initv, int j) {
  int r = -999;
  DISPATCH(h, r = dgetv0<T>(*H_->P, itry, initv != 0, j));
  return r;
}
// H lives in workl[0 : 2 ncv) with ldh = ncv, exactly as dsaupd lays it out.
int oracle_dsaitr(void* h, int k, int np) {
  int r = -999;
  DISPATCH(h, r = dsaitr<T>(*H_->P, k, np, H_->P->workl.data(), H_->P->ncv));
  return r;
}
// shifts: np reals (2 doubles each for Double64).  H in workl[0:2ncv), Q in workl[4ncv : 4ncv+ncv^2).
void oracle_dsapps(void* h, int kev, int np, const void* shifts) {
  if (kev < 1 || np < 1) return;   // outside dsapps!'s domain: dsaup2! never calls it with np = 0 (H[kev+1,1] would not exist)
  DISPATCH(h, {
    using TR = typename vt<T>::real_t;
    int ncv = H_->P->ncv;
    TR* wl = H_->P->workl.data();
    dsapps<T>(*H_->P, kev, np, (const TR*)shifts, wl, ncv, wl + 4 * ncv, ncv);
  });
}
int oracle_dsaupd(void* h, int which, int nev, const double* tol, int maxiter, int initv) {
  int r = -999;
  try {
    DISPATCH(h, {
      using TR = typename vt<T>::real_t;
      Problem<T>& P = *H_->P;
      P.iparam[0] = 1; P.iparam[3] = 1; P.iparam[2] = maxiter; P.iparam[6] = P.op->mode;
      TR t = make_real<TR>(tol);
      r = dsaupd<T>(P, (Which)which, nev, t, P.op->mode, initv != 0);
    });
  } catch (const std::exception& e) {
    fprintf(stderr, "oracle_dsaupd: %s\n", e.what());
    r = -8;
  }
  return r;
}
int oracle_dseupd(void* h, int rvec, void* d, void* Z, int which, int nev, const double* tol) {
  int r = -999;
  try {
    DISPATCH(h, {
      using TR = typename vt<T>::real_t;
      TR t = make_real<TR>(tol);
      r = dseupd<T>(*H_->P, rvec != 0, (TR*)d, (T*)Z, (Which)which, nev, t);
    });
  } catch (const std::exception& e) {
    fprintf(stderr, "oracle_dseupd: %s\n", e.what());
    r = -8;
  }
  return r;
}

// ---- small stand-alone pieces (no handle) -------------------------------------------------
void oracle_dlarnv(int dtype, long long* iseed, long long n, void* x) {
  switch (dtype) {
    case 0: dlarnv_idist_2<double>(iseed, n, (double*)x); break;
    case 1: dlarnv_idist_2<cplx>(iseed, n, (cplx*)x); break;
    case 2: dlarnv_idist_2<dd>(iseed, n, (dd*)x); break;
  }
}
long long oracle_dlarnv_f32(long long* iseed, long long n, float* x) { return dlarnv_idist_2_f32(iseed, n, x); }
void oracle_lcg_table(int* out) { memcpy(out, lcg_table().mm, sizeof(int) * 512); }
void oracle_norm2(int dtype, const void* x, long long n, double* out) {
  out[1] = 0;
  switch (dtype) {
    case 0: out[0] = norm2((const double*)x, n); break;
    case 1: out[0] = norm2((const cplx*)x, n); break;
    case 3: out[0] = norm2((const float*)x, n); break;
    case 2: { dd r = norm2((const dd*)x, n); out[0] = r.hi; out[1] = r.lo; } break;
  }
}
// dstqrb!: eigenvalues into d, last eigenvector row into z.  real dtype 0 (double) or 2 (dd).
int oracle_dstqrb(int dtype, int n, void* d, void* e, void* z) {
  try {
    if (dtype == 0) { std::vector<double> w(2 * n + 2); flexible_dsteqr<double>(n, (double*)d, (double*)e, (double*)z, 1, true, w.data()); }
    else { std::vector<dd> w(2 * n + 2); flexible_dsteqr<dd>(n, (dd*)d, (dd*)e, (dd*)z, 1, true, w.data()); }
  } catch (...) { return 1; }
  return 0;
}
int oracle_dsteqr(int dtype, int n, void* d, void* e, void* Z) {
  try {
    if (dtype == 0) { std::vector<double> w(2 * n + 2); flexible_dsteqr<double>(n, (double*)d, (double*)e, (double*)Z, n, false, w.data()); }
    else { std::vector<dd> w(2 * n + 2); flexible_dsteqr<dd>(n, (dd*)d, (dd*)e, (dd*)Z, n, false, w.data()); }
  } catch (...) { return 1; }
  return 0;
}
void oracle_plane_rotation(double f, double g, double* out) { plane_rotation<double>(f, g, out[0], out[1], out[2]); }
int oracle_dsconv(int n, const double* ritz, const double* bounds, double tol) { return dsconv<double>(n, ritz, bounds, tol); }
void oracle_dsortr(int which, int apply, int n, double* x1, double* x2) { dsortr<double, double>((Which)which, apply != 0, n, x1, x2); }
void oracle_dsgets(int ishift, int which, int kev, int np, double* ritz, double* bounds, double* shifts) {
  dsgets<double>(ishift, (Which)which, kev, np, ritz, bounds, shifts);
}
void oracle_dseigt(double rnorm, int n, const double* H, int ldh, double* eig, double* bounds, double* workl) {
  dseigt<double>(rnorm, n, H, ldh, eig, bounds, workl);
}
// dd arithmetic probes for the Double64 unit tests: op 0 add 1 sub 2 mul 3 div 4 sqrt(a)
void oracle_dd_op(int op, const double* a, const double* b, double* out) {
  dd x(a[0], a[1]), y(b[0], b[1]), r;
  switch (op) { case 0: r = x + y; break; case 1: r = x - y; break; case 2: r = x * y; break; case 3: r = x / y; break; default: r = sqrt(x); }
  out[0] = r.hi; out[1] = r.lo;
}
// 0 = plain binary64 sweeps + x87 norm (default), 1 = exact dots / double-double norm / fixed update order (oracle.hpp)
void oracle_set_arith(int mode) { oracle::arith_mode() = mode == 1 ? oracle::ARITH_EXACT : oracle::ARITH_PLAIN; }
int oracle_get_arith() { return oracle::arith_mode(); }
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
int oracle_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
