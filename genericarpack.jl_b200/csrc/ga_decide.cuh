// On-device restatement of dsaitr!'s per-step scalar logic, so that np Lanczos steps run
// without a host round trip.  Executed by the last CTA of a reduction kernel (single GPU) or by
// a one-CTA kernel after the cross-rank gather (multi GPU).  Citations: /root/reference/src/dsaitr.jl.
#pragma once
#include "ga_reduce.cuh"

namespace ga {

template <class TR> __device__ __forceinline__ TR load_real(const double* p);
template <> __device__ __forceinline__ double load_real<double>(const double* p) { return p[0]; }
template <> __device__ __forceinline__ ddbl load_real<ddbl>(const double* p) { return ddbl(p[0], p[1]); }
__device__ __forceinline__ void store_real(double* p, double v) { p[0] = v; p[1] = 0.0; }
__device__ __forceinline__ void store_real(double* p, ddbl v) { p[0] = v.hi; p[1] = v.lo; }

template <class TR> __device__ __forceinline__ TR real_from_dd(ddbl v);
template <> __device__ __forceinline__ double real_from_dd<double>(ddbl v) { return v.hi + v.lo; }
template <> __device__ __forceinline__ ddbl real_from_dd<ddbl>(ddbl v) { return v; }

// reduced dot slot -> element of T
template <class T> __device__ __forceinline__ T slot_to_elem(double2 s);
template <> __device__ __forceinline__ double slot_to_elem<double>(double2 s) { return s.x + s.y; }
template <> __device__ __forceinline__ float slot_to_elem<float>(double2 s) { return (float)(s.x + s.y); }
template <> __device__ __forceinline__ cdbl slot_to_elem<cdbl>(double2 s) { cdbl z; z.re = s.x; z.im = s.y; return z; }
template <> __device__ __forceinline__ ddbl slot_to_elem<ddbl>(double2 s) { return ddbl(s.x, s.y); }

template <class T> __device__ __forceinline__ T* hbuf_ptr(DevCtl* ctl) { return reinterpret_cast<T*>(ctl->hbuf); }
template <class TR> __device__ __forceinline__ TR* H_ptr(DevCtl* ctl) { return reinterpret_cast<TR*>(ctl->H); }

// `red` holds j dot slots followed by (med, big, sml) norm slots.  All threads of the calling
// CTA enter; `j` is the 1-based Lanczos step (= number of basis columns swept), ncv = ld of H.
template <class T>
__device__ void gs_decide(DevCtl* ctl, const double2* red, int j, int ncv, int kind) {
  typedef typename Elem<T>::real_t TR;
  T* hb = hbuf_ptr<T>(ctl);
  TR* H = H_ptr<TR>(ctl);
  const int jc = j - 1;
  // Everything the scalar logic needs is fetched by different threads at once (one global-memory
  // latency instead of a chain of dependent loads in thread 0; this sits on the critical path
  // between two sweeps).
  __shared__ double sh[16];
  __shared__ int sh_rstart;
  if (kind == DK_NONE) return;
  {
    const int t = threadIdx.x;
    if (t == 0) {  // the previous coefficient vector's j-th entry, before hbuf is overwritten
      TR hj = (j >= 1 && kind != DK_NORM) ? Elem<T>::re(hb[jc < 0 ? 0 : jc]) : TR(0.0);
      store_real(&sh[0], hj);
    } else if (t == 1) {
      sh_rstart = ctl->rstart_j;
    } else if (t == 2) {
      sh[2] = ctl->rnorm[0]; sh[3] = ctl->rnorm[1];
    } else if (t == 3) {
      sh[4] = ctl->wnorm[0]; sh[5] = ctl->wnorm[1];
    } else if (t == 4) {
      TR hd = (j >= 1 && (kind == DK_ORTH1 || kind == DK_ORTH1_NODOTS || kind == DK_ORTH2)) ? H[ncv + jc] : TR(0.0);
      store_real(&sh[6], hd);
    } else if (t >= 5 && t < 8) {
      const double2 v = red[j + (t - 5)];
      sh[8 + 2 * (t - 5)] = v.x; sh[9 + 2 * (t - 5)] = v.y;
    }
  }
  __syncthreads();
  bool take_dots = false;
  if (threadIdx.x == 0) {
    const ddbl nrm_dd = norm_finalize(ddbl(sh[8], sh[9]), ddbl(sh[10], sh[11]), ddbl(sh[12], sh[13]));
    const TR nrm = real_from_dd<TR>(nrm_dd);
    const TR hj = load_real<TR>(&sh[0]);
    const TR rnorm_in = load_real<TR>(&sh[2]);
    const TR wnorm_in = load_real<TR>(&sh[4]);
    const TR hdiag_in = load_real<TR>(&sh[6]);
    const bool first = (jc == 0) || (sh_rstart == j);
    switch (kind) {
      case DK_DOT:  // wnorm = ||OP v_j|| (:1208); h = V_j^T w (:1227)
        store_real(ctl->wnorm, nrm);
        break;
      case DK_CGS: {  // r = w - V h done (:1240)
        H[ncv + jc] = hj;                                             // :1243
        H[jc] = first ? TR(0.0) : rnorm_in;                          // :1259-1263
        store_real(ctl->rnorm, nrm);                                  // :1304
        const TR wn = wnorm_in;
        if (nrm > TR(0.717) * wn) ctl->phase = PH_DONE;               // :1324
        else { ctl->phase = PH_ORTH1; ctl->nrorth += 1; }             // :1328,1335
      } break;
      case DK_ORTH1:
      case DK_ORTH1_NODOTS: {  // r -= V s done (:1363)
        if (first) H[jc] = TR(0.0);                                   // :1364-1366
        H[ncv + jc] = hdiag_in + hj;                                  // :1367
        const TR rn = rnorm_in;
        store_real(ctl->rnorm1, nrm);                                 // :1410
        store_real(ctl->rnorm, nrm);                                  // :1425 / :1431
        if (nrm > TR(0.717) * rn) ctl->phase = PH_DONE;               // :1424
        else { ctl->nitref += 1; ctl->phase = PH_ORTH2; }             // :1430-1434
      } break;
      case DK_ORTH2: {
        if (first) H[jc] = TR(0.0);
        H[ncv + jc] = hdiag_in + hj;
        const TR rn = rnorm_in;
        store_real(ctl->rnorm1, nrm);
        if (nrm > TR(0.717) * rn) store_real(ctl->rnorm, nrm);
        else {                                                        // :1440-1446
          ctl->nitref += 1;
          store_real(ctl->rnorm, TR(0.0));
          ctl->zero_resid = 1;
        }
        ctl->phase = PH_DONE;
      } break;
      case DK_GETV0_DOT:
        store_real(ctl->rnorm0, nrm);
        store_real(ctl->rnorm, nrm);
        break;
      case DK_GETV0_UPD:
        store_real(ctl->rnorm, nrm);
        break;
      case DK_NORM:
        store_real(ctl->rnorm, nrm);
        break;
    }
    if (r_isnan(nrm)) ctl->status = ST_NUMERIC;
  }
  take_dots = (kind == DK_DOT || kind == DK_CGS || kind == DK_ORTH1 || kind == DK_GETV0_DOT || kind == DK_GETV0_UPD ||
               kind == DK_TAKE);
  if (take_dots)
    for (int c = threadIdx.x; c < j; c += blockDim.x) hb[c] = slot_to_elem<T>(red[c]);
}

}  // namespace ga
