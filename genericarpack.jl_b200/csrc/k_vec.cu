// n-vector kernels of the Lanczos step: v_j = r / rnorm (dsaitr step 2), ||resid|| and zero fill.
#include "ga_decide.cuh"

namespace ga {

// ---- dsaitr! STEP 2 (src/dsaitr.jl:1061-1105): v_j = resid * (1/rnorm), or the dlascl-safe
// rescaling of _scale_from_to (src/arpack-blas.jl:25-72) when rnorm < floatmin.  Also the
// step's entry test: rnorm <= 0 means an invariant subspace was found; the device flags it and
// every later kernel of this call returns immediately so the host can run the dgetv0 restart.
template <class TR>
__device__ __forceinline__ void lascl_factor(TR cfrom, TR cto, TR& mul, TR& cfromc, TR& ctoc, bool& done) {
  const TR smlnum = Real<TR>::floatmin();
  const TR bignum = Real<TR>::from(1.0) / smlnum;
  cfromc = cfrom; ctoc = cto;
  const TR cfrom1 = cfromc * smlnum;
  if (cfrom1 == cfromc) { mul = ctoc / cfromc; done = true; }
  else {
    const TR cto1 = ctoc / bignum;
    if (cto1 == ctoc) { mul = ctoc; done = true; cfromc = Real<TR>::from(1.0); }
    else if (r_abs(cfrom1) > r_abs(ctoc) && ctoc != Real<TR>::from(0.0)) { mul = smlnum; done = false; cfromc = cfrom1; }
    else if (r_abs(cto1) > r_abs(cfromc)) { mul = bignum; done = false; ctoc = cto1; }
    else { mul = ctoc / cfromc; done = true; }
  }
}

template <class T, class TR>
__device__ __forceinline__ T scaled_by_inv_norm(T x, TR rnorm, TR temp1, bool safe) {
  if (safe) return Elem<T>::mul_real(x, temp1);                  // src/dsaitr.jl:1099-1101
  TR mul, cf, ct; bool done;                                       // :1103 (_scale_from_to)
  lascl_factor<TR>(rnorm, Real<TR>::from(1.0), mul, cf, ct, done);
  x = Elem<T>::mul_real(x, mul);
  int guard = 0;
  while (!done && guard++ < 8) {
    TR f = cf, t = ct;
    lascl_factor<TR>(f, t, mul, cf, ct, done);
    x = Elem<T>::mul_real(x, mul);
  }
  return x;
}

// Multi-GPU fused path: besides v_j the kernel PUSHES the entries of v_j that the peers' rows
// reference straight into the peers' halo buffers (P2P stores over NVLink) and the last CTA raises
// a sequence flag in each peer's mailbox; the peers' SpMV waits on that flag.  No pack kernel, no
// NCCL send/recv on the per-step path.
template <class T>
__global__ void __launch_bounds__(256) k_normalize(T* vj, const T* resid, i64 n,  // vj may alias resid (in-place scaling)
                                                   DevCtl* ctl, int jstep, const DevComm* cm,
                                                   const int* __restrict__ send_idx) {
  typedef typename Elem<T>::real_t TR;
  if (ctl->status != ST_OK) return;
  const TR rnorm = load_real<TR>(ctl->rnorm);
  if (!(rnorm > Real<TR>::from(0.0))) {  // src/dsaitr.jl:1061
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      ctl->j_break = jstep;
      ctl->status = r_isnan(rnorm) ? ST_NUMERIC : ST_BREAKDOWN;
    }
    return;
  }
  const unsigned int hseq = cm ? ctl->hseq + 1u : 0u;
  const i64 stride = (i64)gridDim.x * blockDim.x;
  const bool safe = rnorm >= Real<TR>::floatmin();
  const TR temp1 = Real<TR>::from(1.0) / rnorm;
  // vj may alias resid, so the compiler cannot batch the loads on its own: stage four elements in
  // registers before the stores (four independent loads in flight per thread)
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n; i += 4 * stride) {
    const T a0 = resid[i], a1 = resid[i + stride], a2 = resid[i + 2 * stride], a3 = resid[i + 3 * stride];
    vj[i] = scaled_by_inv_norm<T, TR>(a0, rnorm, temp1, safe);
    vj[i + stride] = scaled_by_inv_norm<T, TR>(a1, rnorm, temp1, safe);
    vj[i + 2 * stride] = scaled_by_inv_norm<T, TR>(a2, rnorm, temp1, safe);
    vj[i + 3 * stride] = scaled_by_inv_norm<T, TR>(a3, rnorm, temp1, safe);
  }
  for (; i < n; i += stride) vj[i] = scaled_by_inv_norm<T, TR>(resid[i], rnorm, temp1, safe);
  if (!cm) return;
  const int nsend = cm->send_off[cm->world];
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < nsend; i += stride) {
    int q = 0;
    while (i >= cm->send_off[q + 1]) ++q;
    T* dst = reinterpret_cast<T*>(cm->peer_halo[q]) + cm->dest_off[q] + (i - cm->send_off[q]);
    *dst = scaled_by_inv_norm<T, TR>(resid[send_idx[i]], rnorm, temp1, safe);
  }
  __shared__ int sh_last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int t = atomicAdd(&ctl->ticket, 1u);
    sh_last = (t == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!sh_last) return;
  __threadfence_system();
  if ((int)threadIdx.x < cm->world && cm->send_off[threadIdx.x + 1] > cm->send_off[threadIdx.x])
    st_sys_u32(&cm->peer[threadIdx.x]->hflags[cm->rank], hseq);
  if (threadIdx.x == 0) { ctl->hseq = hseq; ctl->ticket = 0u; }
}

// The halo push of k_normalize on its own (fused multi-GPU path): the entries of x that the peers' rows reference go
// straight into the peers' halo buffers, the last CTA raises the sequence flags.  Used where a product is applied to a
// vector that is not being normalised (svds post-processing), so that no NCCL communicator is ever needed on that path
// (creating one costs seconds on 8 GPUs -- more than the whole C5 solve).
template <class T>
__global__ void __launch_bounds__(256) k_push_halo(const T* __restrict__ x, DevCtl* ctl, const DevComm* cm,
                                                   const int* __restrict__ send_idx, int gated) {
  if (gated && ctl->status != ST_OK) return;   // parked (every rank sees the same status: the product that would wait is parked too)
  const unsigned int hseq = ctl->hseq + 1u;
  const i64 stride = (i64)gridDim.x * blockDim.x;
  const int nsend = cm->send_off[cm->world];
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < nsend; i += stride) {
    int q = 0;
    while (i >= cm->send_off[q + 1]) ++q;
    T* dst = reinterpret_cast<T*>(cm->peer_halo[q]) + cm->dest_off[q] + (i - cm->send_off[q]);
    *dst = x[send_idx[i]];
  }
  __shared__ int sh_last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int t = atomicAdd(&ctl->ticket, 1u);
    sh_last = (t == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!sh_last) return;
  __threadfence_system();
  if ((int)threadIdx.x < cm->world && cm->send_off[threadIdx.x + 1] > cm->send_off[threadIdx.x])
    st_sys_u32(&cm->peer[threadIdx.x]->hflags[cm->rank], hseq);
  if (threadIdx.x == 0) { ctl->hseq = hseq; ctl->ticket = 0u; }
}

int launch_push_halo(ga_ctx* c, const void* x, bool gated) {
  if (c->world <= 1 || !c->devcomm) return set_err(c, GA_ERR_ARG, "halo push needs the fused multi-GPU path");
  const int threads = 256;
  i64 blocks = ((i64)c->nsend + threads - 1) / threads;
  const i64 maxb = (i64)c->sm_count * 4;
  if (blocks > maxb) blocks = maxb;
  if (blocks < 1) blocks = 1;
  switch (c->dtype) {
    case GA_F64: k_push_halo<double><<<(int)blocks, threads, 0, c->stream>>>((const double*)x, c->ctl, c->devcomm, c->send_idx, gated ? 1 : 0); break;
    case GA_C64: k_push_halo<cdbl><<<(int)blocks, threads, 0, c->stream>>>((const cdbl*)x, c->ctl, c->devcomm, c->send_idx, gated ? 1 : 0); break;
    case GA_F32: k_push_halo<float><<<(int)blocks, threads, 0, c->stream>>>((const float*)x, c->ctl, c->devcomm, c->send_idx, gated ? 1 : 0); break;
    default: k_push_halo<ddbl><<<(int)blocks, threads, 0, c->stream>>>((const ddbl*)x, c->ctl, c->devcomm, c->send_idx, gated ? 1 : 0); break;
  }
  GA_CUDA(c, cudaGetLastError());
  c->launches += 1;
  return 0;
}

int launch_normalize(ga_ctx* c, int jcol, bool push_halo) {
  const int threads = 256;
  const i64 n = c->vs.rows, ld = c->vs.ld;
  i64 blocks = (n + threads - 1) / threads;
  const i64 maxb = (i64)c->sm_count * 8;
  if (blocks > maxb) blocks = maxb;
  if (blocks < 1) blocks = 1;
  const DevComm* cm = (c->world > 1 && push_halo) ? c->devcomm : nullptr;
  switch (c->dtype) {
    case GA_F64:
      k_normalize<double><<<(int)blocks, threads, 0, c->stream>>>((double*)c->vs.V + (size_t)jcol * ld,
                                                                 (const double*)c->vs.w, n, c->ctl, jcol + 1, cm, c->send_idx);
      break;
    case GA_C64:
      k_normalize<cdbl><<<(int)blocks, threads, 0, c->stream>>>((cdbl*)c->vs.V + (size_t)jcol * ld,
                                                               (const cdbl*)c->vs.w, n, c->ctl, jcol + 1, cm, c->send_idx);
      break;
    case GA_F32:
      k_normalize<float><<<(int)blocks, threads, 0, c->stream>>>((float*)c->vs.V + (size_t)jcol * ld,
                                                               (const float*)c->vs.w, n, c->ctl, jcol + 1, cm, c->send_idx);
      break;
    default:
      k_normalize<ddbl><<<(int)blocks, threads, 0, c->stream>>>((ddbl*)c->vs.V + (size_t)jcol * ld,
                                                               (const ddbl*)c->vs.w, n, c->ctl, jcol + 1, cm, c->send_idx);
      break;
  }
  GA_CUDA(c, cudaGetLastError());
  c->launches += 1;
  return 0;
}

// ---- ||resid||_2 (norm2(TR, resid): src/arpack-blas-nrm2.jl:355-372; call sites
// src/dsaup2.jl:1414, src/dgetv0.jl:630) ------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(256) k_norm(const T* __restrict__ x, i64 n, DevCtl* ctl, double2* partial,
                                              int decide_kind, int ncv, const DevComm* cm) {
  if (ctl->status != ST_OK && ctl->status != ST_BREAKDOWN) return;
  __shared__ double2 sw[8 * 3];
  __shared__ int sh_flag;
  NormAcc a;
  a.init();
  const i64 stride = (i64)gridDim.x * blockDim.x;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) a.add(x[i]);
  ddbl m = warp_sum_dd(a.med), b = warp_sum_dd(a.big), s = warp_sum_dd(a.sml);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    sw[warp * 3 + 0] = make_double2(m.hi, m.lo);
    sw[warp * 3 + 1] = make_double2(b.hi, b.lo);
    sw[warp * 3 + 2] = make_double2(s.hi, s.lo);
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double2 acc = make_double2(0.0, 0.0);
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) acc = slot_add<false>(acc, sw[w * 3 + threadIdx.x]);
    partial[(size_t)blockIdx.x * kMaxSlots + threadIdx.x] = acc;
  }
  if (last_block_reduce<false>(ctl, partial, 0, 3, &sh_flag)) {
    if (cm) comm_exchange_slots<false>(ctl, cm, 0, 3);
    gs_decide<T>(ctl, ctl->red, 0, ncv, decide_kind);
  }
}

int launch_norm_resid(ga_ctx* c, int decide_kind) {
  const int threads = 256;
  const i64 n = c->vs.rows;
  i64 blocks = (n + threads * 4 - 1) / (threads * 4);
  const i64 maxb = (i64)c->sm_count * 4;
  if (blocks > maxb) blocks = maxb;
  if (blocks < 1) blocks = 1;
  const bool single = c->world == 1 || c->devcomm != nullptr;  // decide inside the kernel
  const int dk = single ? decide_kind : DK_NONE;
  const DevComm* cm = c->world > 1 ? c->devcomm : nullptr;
  switch (c->dtype) {
    case GA_F64: k_norm<double><<<(int)blocks, threads, 0, c->stream>>>((const double*)c->vs.w, n, c->ctl, c->partial, dk, c->ncv, cm); break;
    case GA_C64: k_norm<cdbl><<<(int)blocks, threads, 0, c->stream>>>((const cdbl*)c->vs.w, n, c->ctl, c->partial, dk, c->ncv, cm); break;
    case GA_F32: k_norm<float><<<(int)blocks, threads, 0, c->stream>>>((const float*)c->vs.w, n, c->ctl, c->partial, dk, c->ncv, cm); break;
    default: k_norm<ddbl><<<(int)blocks, threads, 0, c->stream>>>((const ddbl*)c->vs.w, n, c->ctl, c->partial, dk, c->ncv, cm); break;
  }
  GA_CUDA(c, cudaGetLastError());
  c->launches += 1;
  if (!single) return launch_decide(c, 0, -1, decide_kind);
  return 0;
}

int launch_zero_resid(ga_ctx* c) {
  GA_CUDA(c, cudaMemsetAsync(c->resid, 0, (size_t)c->n_pad * c->esize, c->stream));
  return 0;
}

}  // namespace ga
