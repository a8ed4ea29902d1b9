// Basis updates V <- V*Q with a small replicated real Q:
//   - device tail of dsapps! (src/dsapps.jl:813-854): V[:,1:kev(+1)] <- V[:,1:kplusp] Q[:,1:kev(+1)],
//     resid <- Q[kplusp,kev] resid + H[kev+1,1] V[:,kev+1], fused with the ||resid|| that dsaup2
//     computes next (src/dsaup2.jl:1397-1414).  The reference issues kev+1 separate dgemv calls
//     over shrinking column ranges plus 3 kev n-vector copies; here every row of V is read once
//     and written once (row-independent, hence safe in place).
//   - device tail of simple_dseupd! (src/dseupd.jl:1413-1435): V <- V Qh where Qh is the product of
//     the nconv Householder reflectors the reference applies one at a time with dorm2r!.
//
// A CTA stages R rows x kplusp columns of V in shared memory (coalesced 16-byte loads), then each
// thread owns one row and OB = 4 consecutive output columns at a time: one shared-memory read of
// x = V[row,c] and OB reads of Q[c, o..o+3] feed OB FMAs.  Q sits in shared memory transposed to
// [c][o] so the OB values are contiguous.
#include "ga_decide.cuh"

namespace ga {

template <class T> struct VqCfg {
  static constexpr int R = (sizeof(T) == 8) ? 128 : 64;  // rows per CTA tile
  static constexpr int OG = (sizeof(T) == 8) ? 2 : 4;    // output groups (threads per row)
  static constexpr int THREADS = R * OG;
  static constexpr int OB = 4;
};

template <class T>
__global__ void __launch_bounds__(VqCfg<T>::THREADS)
k_vq(T* V, i64 ld, i64 ntiles, const typename Elem<T>::real_t* __restrict__ Qg, int ldq, int kplusp, int nout,
     T* resid, int do_resid, int kev, DevCtl* ctl, double2* partial, int decide_kind, int ncv, const DevComm* cm) {
  typedef typename Elem<T>::real_t TR;
  typedef VqCfg<T> C;
  constexpr int R = C::R, OG = C::OG, OB = C::OB;
  extern __shared__ __align__(16) unsigned char smem[];
  const int nout_pad = (nout + OB - 1) / OB * OB;
  TR* Qs = reinterpret_cast<TR*>(smem);                                   // [kplusp][nout_pad]
  T* tile = reinterpret_cast<T*>(smem + (size_t)kplusp * nout_pad * sizeof(TR));  // [kplusp][R]
  __shared__ double2 sw[C::THREADS / 32 * 3];
  __shared__ int sh_flag;
  const int tid = threadIdx.x;
  for (int idx = tid; idx < kplusp * nout_pad; idx += C::THREADS) {
    const int c = idx / nout_pad, o = idx % nout_pad;
    Qs[idx] = (o < nout) ? Qg[c + (size_t)o * ldq] : Real<TR>::from(0.0);
  }
  TR sigma = Real<TR>::from(0.0), beta = Real<TR>::from(0.0);
  if (do_resid) {
    sigma = Qg[(kplusp - 1) + (size_t)(kev - 1) * ldq];   // Q[kplusp,kev]   (src/dsapps.jl:850)
    beta = load_real<TR>(ctl->rnorm1);                    // H[kev+1,1], parked by the launcher
  }
  const int row = tid % R, og = tid / R;
  NormAcc nacc;
  nacc.init();
  for (i64 t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const i64 row0 = t * R;
    __syncthreads();
    for (int idx = tid; idx < kplusp * R; idx += C::THREADS) {
      const int c = idx / R, rr = idx % R;
      tile[idx] = V[(size_t)c * ld + row0 + rr];
    }
    __syncthreads();
    for (int ob = og * OB; ob < nout_pad; ob += OG * OB) {
      T y0 = Elem<T>::zero(), y1 = Elem<T>::zero(), y2 = Elem<T>::zero(), y3 = Elem<T>::zero();
      const TR* q = Qs + ob;
#pragma unroll 4
      for (int c = 0; c < kplusp; ++c) {
        const T x = tile[c * R + row];
        const TR* qc = q + (size_t)c * nout_pad;
        Elem<T>::add_mul_real(y0, x, qc[0]);
        Elem<T>::add_mul_real(y1, x, qc[1]);
        Elem<T>::add_mul_real(y2, x, qc[2]);
        Elem<T>::add_mul_real(y3, x, qc[3]);
      }
      T yy[4] = {y0, y1, y2, y3};
#pragma unroll
      for (int u = 0; u < OB; ++u) {
        const int o = ob + u;
        if (o < nout) {
          V[(size_t)o * ld + row0 + row] = yy[u];
          if (do_resid && o == kev) {  // beta > 0: r <- sigma r + beta v_{kev+1}  (:850-854)
            T rv = Elem<T>::mul_real(resid[row0 + row], sigma);
            Elem<T>::add_mul_real(rv, yy[u], beta);
            resid[row0 + row] = rv;
            nacc.add(rv);
          }
        }
      }
    }
    if (do_resid && nout == kev && og == 0) {  // beta == 0: r <- sigma r
      T rv = Elem<T>::mul_real(resid[row0 + row], sigma);
      resid[row0 + row] = rv;
      nacc.add(rv);
    }
  }
  if (!do_resid) return;
  ddbl m = warp_sum_dd(nacc.med), b = warp_sum_dd(nacc.big), s = warp_sum_dd(nacc.sml);
  const int warp = tid >> 5, lane = tid & 31;
  if (lane == 0) {
    sw[warp * 3 + 0] = make_double2(m.hi, m.lo);
    sw[warp * 3 + 1] = make_double2(b.hi, b.lo);
    sw[warp * 3 + 2] = make_double2(s.hi, s.lo);
  }
  __syncthreads();
  if (tid < 3) {
    double2 acc = make_double2(0.0, 0.0);
    for (int w = 0; w < C::THREADS / 32; ++w) acc = slot_add<false>(acc, sw[w * 3 + tid]);
    partial[(size_t)blockIdx.x * kMaxSlots + tid] = acc;
  }
  if (last_block_reduce<false>(ctl, partial, 0, 3, &sh_flag)) {
    if (cm) comm_exchange_slots<false>(ctl, cm, 0, 3);
    gs_decide<T>(ctl, ctl->red, 0, ncv, decide_kind);
  }
}

template <class T>
static int launch_vq_t(ga_ctx* c, int kplusp, int nout, bool do_resid, int kev) {
  typedef typename Elem<T>::real_t TR;
  typedef VqCfg<T> C;
  const i64 ntiles = c->n_pad / C::R;
  const int nout_pad = (nout + C::OB - 1) / C::OB * C::OB;
  const size_t smem = (size_t)kplusp * nout_pad * sizeof(TR) + (size_t)kplusp * C::R * sizeof(T);
  static PerDeviceOnce attr_set;
  if (attr_set.need(c->device)) {
    GA_CUDA(c, cudaFuncSetAttribute(k_vq<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set.done(c->device);
  }
  if (smem > 200 * 1024) return set_err(c, GA_ERR_ARG, "vq: shared memory budget exceeded");
  int per_sm = (int)((220 * 1024) / (smem + 2048));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;
  i64 grid = (i64)c->sm_count * per_sm;
  if (grid > ntiles) grid = ntiles;
  if (grid > kMaxCtas) grid = kMaxCtas;
  if (grid < 1) grid = 1;
  const bool single = c->world == 1 || c->devcomm != nullptr;  // decide inside the kernel
  k_vq<T><<<(int)grid, C::THREADS, smem, c->stream>>>((T*)c->V, c->n_pad, ntiles, (const TR*)c->qdev, c->ncv_pad,
                                                    kplusp, nout, (T*)c->resid, do_resid ? 1 : 0, kev, c->ctl,
                                                    c->partial, single ? DK_NORM : DK_NONE, c->ncv, c->world > 1 ? c->devcomm : nullptr);
  GA_CUDA(c, cudaGetLastError());
  c->launches += 1;
  if (do_resid && !single) return launch_decide(c, 0, -1, DK_NORM);
  return 0;
}

// Q (kplusp x nout, ld = ncv_pad, real type) must already be in c->qdev; for do_resid the
// launcher parks beta = H[kev+1,1] in ctl->rnorm1 first.
int launch_vq(ga_ctx* c, int kplusp, int nout, bool do_resid, const void* sigma, const void* beta, int kev) {
  (void)sigma;
  if (do_resid) {
    double b[2] = {0.0, 0.0};
    b[0] = ((const double*)beta)[0];
    if (c->dtype == GA_DD) b[1] = ((const double*)beta)[1];
    GA_CUDA(c, cudaMemcpyAsync(c->ctl->rnorm1, b, sizeof(b), cudaMemcpyHostToDevice, c->stream));
  }
  switch (c->dtype) {
    case GA_F64: return launch_vq_t<double>(c, kplusp, nout, do_resid, kev);
    case GA_C64: return launch_vq_t<cdbl>(c, kplusp, nout, do_resid, kev);
    case GA_F32: return launch_vq_t<float>(c, kplusp, nout, do_resid, kev);
    default: return launch_vq_t<ddbl>(c, kplusp, nout, do_resid, kev);
  }
}

}  // namespace ga
