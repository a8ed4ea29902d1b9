// Generalised symmetric eigenproblem A x = lambda B x, ARPACK mode 2 with bmat = :G
// (ArpackSymmetricGeneralizedOp, src/idonow_ops.jl:462-492; reached by eigs(Symmetric(A), Symmetric(B), k),
// src/interface.jl:143-163).
//
// What changes against the standard problem (mode 1, bmat = :I):
//   * OP v = inv(B) (A v); genopx! also hands A v back (ARPACK remark 5, src/idonow_ops.jl:471-487), because
//     for mode 2 the B-inner products with OP v are taken with A v = B OP v (src/dsaitr.jl:1195-1200,1232);
//   * every norm is a B-norm sqrt(|x' B x|): one extra B*x per norm (src/dsaitr.jl:1272-1300,1373-1407,
//     src/dgetv0.jl:596-627,669-698, src/dsaup2.jl:1383-1411), and the Gram-Schmidt coefficients are
//     V' (B r), so the basis is B-orthonormal.
// The tall-skinny work is the same two kernels as the standard path (k_gs MODE 0 with the dot vector
// B r / A v_j, MODE 2 for r -= V h), the SpMVs are the streamed CSR kernel.  B r is not row-local, so the
// standard path's fusion of "update + next dots" does not apply and the step's scalar logic (the 0.717
// tests) runs on the host from three small reductions per Gram-Schmidt round.
//
// The solve with B (the reference takes any factorisation object S from the caller) is a
// Jacobi-preconditioned conjugate-gradient iteration: x = 0, r = b, p = z = D^-1 r and per iteration
// q = B p (streamed SpMV), alpha = (r,z)/(p,q), x += alpha p, r -= alpha q, z = D^-1 r, beta = (r,z)'/(r,z),
// p = z + beta p.  The scalars stay on the device; the host enqueues batches of iterations and looks at the
// control block once per batch; when the residual test passes the device parks the rest of the batch
// (status ST_CG_DONE).  All reductions are fixed-order double-double sums (bit-reproducible).
//
// Several GPUs (ga_set_generalized_csr_dist): rows of A, B and of every vector are block-partitioned like the standard
// problem, A and B share ONE halo plan (the union of their remote columns).  Every product pushes the operand's halo
// entries into the peers' buffers (k_push_halo) and waits on their flags inside the SpMV; every reduction ends with the
// mailbox exchange of the Gram-Schmidt sweeps (comm_exchange_slots), so all ranks hold bit-identical scalars, take the
// same branches and park the same CG iterations.  Two products are always separated by at least one such exchange
// (CG: (p,q) and (r,z); Lanczos: the B-norms and projections), which is what makes the single halo buffer safe: a
// peer cannot push the next operand before this rank has finished reading the current one.  Peer-memory path only.
#include "ga_decide.cuh"

#include <chrono>
#include <cmath>
#include <cstring>
#include <vector>

namespace ga {

// ---- exact products accumulated in double-double: (re, im) += conj(a) * b --------------------------------
template <class T> struct DotAcc;
template <> struct DotAcc<double> {
  static __device__ __forceinline__ void add(ddbl& re, ddbl& im, double a, double b) {
    double p, e;
    two_prod(a, b, p, e);
    re = dd_add(re, ddbl(p, e));
  }
};
template <> struct DotAcc<cdbl> {
  static __device__ __forceinline__ void add(ddbl& re, ddbl& im, cdbl a, cdbl b) {
    double p, e;
    two_prod(a.re, b.re, p, e); re = dd_add(re, ddbl(p, e));
    two_prod(a.im, b.im, p, e); re = dd_add(re, ddbl(p, e));
    two_prod(a.re, b.im, p, e); im = dd_add(im, ddbl(p, e));
    two_prod(-a.im, b.re, p, e); im = dd_add(im, ddbl(p, e));
  }
};
template <> struct DotAcc<ddbl> {
  static __device__ __forceinline__ void add(ddbl& re, ddbl& im, ddbl a, ddbl b) { re = dd_add(re, dd_mul(a, b)); }
};

// CTA-level fixed-order reduction of NS double-double accumulators into partial[cta][0..NS), then the
// grid-level last-block reduction into ctl->red[0..NS).  Returns true in the last CTA.
template <int NS>
__device__ __forceinline__ bool reduce_slots(DevCtl* ctl, double2* partial, ddbl (&v)[NS], const DevComm* cm) {
  __shared__ double2 sw[8 * NS];
  __shared__ int sh_flag;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < NS; ++k) {
    const ddbl s = warp_sum_dd(v[k]);
    if (lane == 0) sw[warp * NS + k] = make_double2(s.hi, s.lo);
  }
  __syncthreads();
  if (threadIdx.x < NS) {
    double2 acc = make_double2(0.0, 0.0);
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) acc = slot_add<false>(acc, sw[w * NS + threadIdx.x]);
    partial[(size_t)blockIdx.x * kMaxSlots + threadIdx.x] = acc;
  }
  const bool last = last_block_reduce<false>(ctl, partial, 0, NS, &sh_flag);
  if (last && cm) comm_exchange_slots<false>(ctl, cm, 0, NS);   // several GPUs: ctl->red <- the world's sums, rank order
  return last;
}

// ---- dot(a, b) = sum conj(a_i) b_i  ->  ctl->red[0] = Re (hi, lo), ctl->red[1] = Im (hi, lo) ------------------
template <class T>
__global__ void __launch_bounds__(256) k_dotc(const T* __restrict__ a, const T* __restrict__ b, i64 n, DevCtl* ctl,
                                              double2* partial, const DevComm* cm) {
  if (ctl->status != ST_OK) return;
  ddbl v[2] = {ddbl(0.0, 0.0), ddbl(0.0, 0.0)};
  const i64 stride = (i64)gridDim.x * blockDim.x;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) DotAcc<T>::add(v[0], v[1], a[i], b[i]);
  reduce_slots<2>(ctl, partial, v, cm);
}

// ---- conjugate gradients with B ----------------------------------------------------------------------------
template <class T, class TR>
__global__ void __launch_bounds__(256) k_cg_init(const T* __restrict__ b, T* __restrict__ x, T* __restrict__ r,
                                                 T* __restrict__ p, const TR* __restrict__ dinv, i64 n, DevCtl* ctl,
                                                 double2* partial, const DevComm* cm) {
  if (ctl->status != ST_OK) return;
  ddbl v[2] = {ddbl(0.0, 0.0), ddbl(0.0, 0.0)};   // (r, z), (b, b)
  ddbl unused(0.0, 0.0);
  const i64 stride = (i64)gridDim.x * blockDim.x;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const T bi = b[i];
    const T zi = Elem<T>::mul_real(bi, dinv[i]);
    x[i] = Elem<T>::zero();
    r[i] = bi;
    p[i] = zi;
    DotAcc<T>::add(v[0], unused, bi, zi);
    DotAcc<T>::add(v[1], unused, bi, bi);
  }
  if (reduce_slots<2>(ctl, partial, v, cm) && threadIdx.x == 0) {
    ctl->cg_rz[0] = ctl->red[0].x; ctl->cg_rz[1] = ctl->red[0].y;
    ctl->cg_bb[0] = ctl->red[1].x; ctl->cg_bb[1] = ctl->red[1].y;
    ctl->cg_rr[0] = ctl->red[1].x; ctl->cg_rr[1] = ctl->red[1].y;
    ctl->cg_iters = 0;
    if (!(ctl->red[1].x > 0.0)) ctl->status = (ctl->red[1].x == 0.0) ? ST_CG_DONE : ST_NUMERIC;   // b = 0 -> x = 0
  }
}

template <class T>
__global__ void __launch_bounds__(256) k_cg_pq(const T* __restrict__ p, const T* __restrict__ q, i64 n, DevCtl* ctl,
                                               double2* partial, const DevComm* cm) {
  if (ctl->status != ST_OK) return;
  ddbl v[1] = {ddbl(0.0, 0.0)};
  ddbl unused(0.0, 0.0);
  const i64 stride = (i64)gridDim.x * blockDim.x;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) DotAcc<T>::add(v[0], unused, p[i], q[i]);
  if (reduce_slots<1>(ctl, partial, v, cm) && threadIdx.x == 0) {
    ctl->cg_pq[0] = ctl->red[0].x; ctl->cg_pq[1] = ctl->red[0].y;
    if (!(ctl->red[0].x > 0.0)) ctl->status = ST_NUMERIC;    // B is not positive definite (or p = 0)
  }
}

template <class TR> __device__ __forceinline__ TR cg_real(const double* p);
template <> __device__ __forceinline__ double cg_real<double>(const double* p) { return p[0] + p[1]; }
template <> __device__ __forceinline__ ddbl cg_real<ddbl>(const double* p) { return ddbl(p[0], p[1]); }
__device__ __forceinline__ void cg_store(double* p, double v) { p[0] = v; p[1] = 0.0; }
__device__ __forceinline__ void cg_store(double* p, ddbl v) { p[0] = v.hi; p[1] = v.lo; }

template <class T, class TR>
__global__ void __launch_bounds__(256) k_cg_update(T* __restrict__ x, T* __restrict__ r, const T* __restrict__ p,
                                                   const T* __restrict__ q, const TR* __restrict__ dinv, i64 n,
                                                   DevCtl* ctl, double2* partial, const DevComm* cm) {
  if (ctl->status != ST_OK) return;
  const TR rz = cg_real<TR>(ctl->cg_rz), pq = cg_real<TR>(ctl->cg_pq);
  const TR alpha = rz / pq;
  const TR malpha = -alpha;
  ddbl v[2] = {ddbl(0.0, 0.0), ddbl(0.0, 0.0)};   // (r, z)_new, (r, r)_new
  ddbl unused(0.0, 0.0);
  const i64 stride = (i64)gridDim.x * blockDim.x;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    T xi = x[i], ri = r[i];
    Elem<T>::add_mul_real(xi, p[i], alpha);
    Elem<T>::add_mul_real(ri, q[i], malpha);
    x[i] = xi;
    r[i] = ri;
    const T zi = Elem<T>::mul_real(ri, dinv[i]);
    DotAcc<T>::add(v[0], unused, ri, zi);
    DotAcc<T>::add(v[1], unused, ri, ri);
  }
  if (reduce_slots<2>(ctl, partial, v, cm) && threadIdx.x == 0) {
    const TR rz_new = cg_real<TR>(reinterpret_cast<const double*>(&ctl->red[0]));
    cg_store(ctl->cg_beta, rz_new / rz);
    ctl->cg_rz[0] = ctl->red[0].x; ctl->cg_rz[1] = ctl->red[0].y;
    ctl->cg_rr[0] = ctl->red[1].x; ctl->cg_rr[1] = ctl->red[1].y;
    ctl->cg_iters += 1;
    // ||r||^2 <= tol^2 ||b||^2 (compared in double: both sides are far from the rounding level of the test)
    const double rr = ctl->red[1].x, bb = ctl->cg_bb[0], tol2 = ctl->cg_tol2[0];
    if (!(rr == rr)) ctl->status = ST_NUMERIC;
    else if (rr <= tol2 * bb) ctl->status = ST_CG_DONE;
  }
}

template <class T, class TR>
__global__ void __launch_bounds__(256) k_cg_p(const T* __restrict__ r, T* __restrict__ p, const TR* __restrict__ dinv,
                                              i64 n, const DevCtl* ctl) {
  if (ctl->status != ST_OK) return;
  const TR beta = cg_real<TR>(ctl->cg_beta);
  const i64 stride = (i64)gridDim.x * blockDim.x;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    T zi = Elem<T>::mul_real(r[i], dinv[i]);
    Elem<T>::add_mul_real(zi, p[i], beta);
    p[i] = zi;
  }
}

// ------------------------------------------------------------------------------------------------ host side
static double gnow() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static int vec_grid(const ga_ctx* c) {
  i64 blocks = (c->n + 256 * 4 - 1) / (256 * 4);
  const i64 maxb = (i64)c->sm_count * 4;
  if (blocks > maxb) blocks = maxb;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}
static const DevComm* gcm(const ga_ctx* c) { return c->world > 1 ? c->devcomm : nullptr; }
static int need_peer_path(ga_ctx* c) {
  if (c->world > 1 && !c->devcomm)
    return set_err(c, GA_ERR_COMM, "generalised problem on several GPUs: the peer-memory path is required (ga_comm_ipc_export / ga_comm_ipc_import)");
  return 0;
}
// y = M x for the (possibly row-sharded) A or B: on several GPUs the operand's halo entries are pushed first
static int general_spmv(ga_ctx* c, const CsrDev& M, const void* x, void* y, bool gated) {
  if (c->world <= 1) return launch_spmv_mat(c, M, x, y, gated);
  int rc = need_peer_path(c);
  if (rc) return rc;
  if ((rc = launch_push_halo(c, x, gated))) return rc;
  return launch_spmv_mat(c, M, x, y, gated, true);
}
static int set_status_ok(ga_ctx* c) {
  static const int zeros[4] = {0, 0, 0, 0};  // status, phase, j_break, rstart_j
  GA_CUDA(c, cudaMemcpyAsync(c->ctl, zeros, sizeof(zeros), cudaMemcpyHostToDevice, c->stream));
  return 0;
}

// real scalar of the host-side step logic: (hi, lo); lo = 0 unless Double64
struct HReal {
  double hi = 0.0, lo = 0.0;
};
static HReal bnorm_from_red(const ga_ctx* c, const DevCtl* hc) {   // sqrt(abs(dot)) (src/dsaitr.jl:1195-1204)
  HReal out;
  if (c->dtype == GA_DD) {
    const ddbl s = r_sqrt(r_abs(ddbl(hc->red[0].x, hc->red[0].y)));
    out.hi = s.hi; out.lo = s.lo;
  } else if (c->dtype == GA_C64) {
    out.hi = std::sqrt(std::hypot(hc->red[0].x + hc->red[0].y, hc->red[1].x + hc->red[1].y));
  } else {
    out.hi = std::sqrt(std::fabs(hc->red[0].x + hc->red[0].y));
  }
  return out;
}
static bool h_gt_717(const ga_ctx* c, HReal a, HReal b) {            // a > 0.717 b
  if (c->dtype == GA_DD) return ddbl(a.hi, a.lo) > dd_mul_d(ddbl(b.hi, b.lo), 0.717);
  return a.hi > 0.717 * b.hi;
}
static bool h_pos(HReal a) { return a.hi > 0.0 || (a.hi == 0.0 && a.lo > 0.0); }

int launch_dotc(ga_ctx* c, const void* a, const void* b) {
  const int grid = vec_grid(c);
  switch (c->dtype) {
    case GA_F64: k_dotc<double><<<grid, 256, 0, c->stream>>>((const double*)a, (const double*)b, c->n, c->ctl, c->partial, gcm(c)); break;
    case GA_C64: k_dotc<cdbl><<<grid, 256, 0, c->stream>>>((const cdbl*)a, (const cdbl*)b, c->n, c->ctl, c->partial, gcm(c)); break;
    default: k_dotc<ddbl><<<grid, 256, 0, c->stream>>>((const ddbl*)a, (const ddbl*)b, c->n, c->ctl, c->partial, gcm(c)); break;
  }
  GA_CUDA(c, cudaGetLastError());
  c->launches += 1;
  return 0;
}

// sqrt|x' y| on the host after one reduction kernel and one look at the control block
static int bnorm(ga_ctx* c, const void* x, const void* y, HReal* out) {
  int rc = launch_dotc(c, x, y);
  if (rc) return rc;
  if ((rc = sync_ctl(c))) return rc;
  if (c->ctl_host->status != ST_OK) return set_err(c, GA_ERR_NUMERIC, "generalised problem: device status set during a B-norm");
  *out = bnorm_from_red(c, c->ctl_host);
  if (out->hi != out->hi) return set_err(c, GA_ERR_NUMERIC, "generalised problem: non-finite B-norm");
  return 0;
}

// One CG iteration = the B product (streamed SpMV) + three fused vector kernels; every launch early-exits once the
// convergence flag is set on the device.  At the sizes where this solver matters (n ~ 1e5) an iteration is four launches
// of a few microseconds each and the HOST is the bottleneck (round 1: 0.78 ms per Lanczos step at n = 90 000), so a batch
// of 8 iterations is captured once per context into a CUDA graph and replayed: one launch call per 32 kernels.  The graph
// always solves into the context's own vector (cg_x) so that its kernel arguments never change.
template <class T, class TR>
static int cg_enqueue_iterations(ga_ctx* c, int count) {
  const int grid = vec_grid(c);
  const i64 n = c->n;
  T* r = (T*)c->cg_r; T* p = (T*)c->cg_p; T* q = (T*)c->cg_q; T* x = (T*)c->cg_x;
  const TR* dinv = (const TR*)c->b_dinv;
  for (int k = 0; k < count; ++k) {
    int rc;
    if ((rc = general_spmv(c, c->Bm, p, q, true))) return rc;
    k_cg_pq<T><<<grid, 256, 0, c->stream>>>(p, q, n, c->ctl, c->partial, gcm(c));
    k_cg_update<T, TR><<<grid, 256, 0, c->stream>>>(x, r, p, q, dinv, n, c->ctl, c->partial, gcm(c));
    k_cg_p<T, TR><<<grid, 256, 0, c->stream>>>(r, p, dinv, n, c->ctl);
    GA_CUDA(c, cudaGetLastError());
    c->launches += 3;
  }
  return 0;
}

constexpr int kCgGraphIters = 8;

template <class T, class TR>
static void cg_build_graph(ga_ctx* c) {
  c->cg_graph_tried = true;
  if (!c->cg_use_graph) return;
  const size_t launches0 = c->launches;
  if (cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return; }
  const int rc = cg_enqueue_iterations<T, TR>(c, kCgGraphIters);
  cudaGraph_t g = nullptr;
  const cudaError_t e = cudaStreamEndCapture(c->stream, &g);
  c->cg_graph_launches = c->launches - launches0;
  c->launches = launches0;
  if (rc != 0 || e != cudaSuccess || !g) { cudaGetLastError(); if (g) cudaGraphDestroy(g); return; }
  cudaGraphExec_t ex = nullptr;
  if (cudaGraphInstantiate(&ex, g, 0) == cudaSuccess) c->cg_graph = ex;
  else cudaGetLastError();
  cudaGraphDestroy(g);
}

template <class T, class TR>
static int cg_solve_t(ga_ctx* c, const void* b, void* x) {
  const int grid = vec_grid(c);
  const i64 n = c->n;
  T* r = (T*)c->cg_r; T* p = (T*)c->cg_p;
  const TR* dinv = (const TR*)c->b_dinv;
  double tol = c->cg_tol[0] > 0.0 ? c->cg_tol[0] : (c->dtype == GA_DD ? 8.0 * 4.930380657631324e-32 : 8.0 * 2.220446049250313e-16);
  const double tol2[2] = {tol * tol, 0.0};
  GA_CUDA(c, cudaMemcpyAsync(c->ctl->cg_tol2, tol2, sizeof(tol2), cudaMemcpyHostToDevice, c->stream));
  k_cg_init<T, TR><<<grid, 256, 0, c->stream>>>((const T*)b, (T*)c->cg_x, r, p, dinv, n, c->ctl, c->partial, gcm(c));
  GA_CUDA(c, cudaGetLastError());
  c->launches += 1;
  if (!c->cg_graph_tried) cg_build_graph<T, TR>(c);
  const int maxit = c->cg_maxit > 0 ? c->cg_maxit : 1000;
  int it = 0, rc;
  bool done = false;
  int batch = 8;
  while (!done) {
    if (c->cg_graph) {                                   // whole graphs of kCgGraphIters iterations (extra ones are no-ops once converged)
      for (int k = 0; k < batch && it < maxit; k += kCgGraphIters, it += kCgGraphIters) {
        GA_CUDA(c, cudaGraphLaunch((cudaGraphExec_t)c->cg_graph, c->stream));
        c->launches += c->cg_graph_launches;
      }
    } else {
      const int cnt = std::min(batch, maxit - it);
      if ((rc = cg_enqueue_iterations<T, TR>(c, cnt))) return rc;
      it += cnt;
    }
    if ((rc = sync_ctl(c))) return rc;
    const DevCtl* hc = c->ctl_host;
    if (hc->status == ST_CG_DONE) done = true;
    else if (hc->status != ST_OK) {
      set_status_ok(c);
      return set_err(c, GA_ERR_SOLVE, "generalised problem: CG with B broke down (B not positive definite, or non-finite data)");
    } else if (it >= maxit) {
      return set_err(c, GA_ERR_SOLVE, "generalised problem: CG with B did not reach the residual target in cg_maxit iterations");
    }
    if (batch < 32) batch *= 2;
  }
  const DevCtl* hc = c->ctl_host;
  c->solve_count += 1;
  c->solve_iters += hc->cg_iters;
  if (hc->cg_bb[0] > 0.0) {
    const double rel = std::sqrt(hc->cg_rr[0] / hc->cg_bb[0]);
    if (rel > c->solve_max_relres) c->solve_max_relres = rel;
  }
  GA_CUDA(c, cudaMemcpyAsync(x, c->cg_x, (size_t)c->n * c->esize, cudaMemcpyDeviceToDevice, c->stream));
  return set_status_ok(c);
}

// x <- inv(B) b (device vectors, x != b)
static int general_solve(ga_ctx* c, const void* b, void* x) {
  if (c->solve_cb) {
    const int rc = c->solve_cb(c->solve_user, b, x, (int64_t)c->n, (void*)c->stream);
    if (rc) return set_err(c, GA_ERR_SOLVE, "generalised problem: the user solve callback failed");
    c->solve_count += 1;
    return 0;
  }
  switch (c->dtype) {
    case GA_F64: return cg_solve_t<double, double>(c, b, x);
    case GA_C64: return cg_solve_t<cdbl, double>(c, b, x);
    default: return cg_solve_t<ddbl, ddbl>(c, b, x);
  }
}

void general_free(ga_ctx* c) {
  c->Bm.free_all();
  if (c->cg_graph) { cudaGraphExecDestroy((cudaGraphExec_t)c->cg_graph); c->cg_graph = nullptr; }
  c->cg_graph_tried = false;
  cudaFree(c->gz); cudaFree(c->cg_r); cudaFree(c->cg_p); cudaFree(c->cg_q); cudaFree(c->cg_x); cudaFree(c->b_dinv);
  c->gz = c->cg_r = c->cg_p = c->cg_q = c->cg_x = c->b_dinv = nullptr;
}

// ncols: n on one GPU; n_pad + nhalo for this rank's rows with the columns remapped to [own | halo] (ga_set_csr_dist)
int general_upload(ga_ctx* c, i64 ncols, const int64_t* arp, const int32_t* acol, const void* aval, const int64_t* brp,
                   const int32_t* bcol, const void* bval) {
  general_free(c);
  const i64 n = c->n;
  int rc = csr_upload(c, c->A, n, ncols, arp, acol, aval, 0);
  if (rc) return rc;
  if ((rc = csr_upload(c, c->Bm, n, ncols, brp, bcol, bval, 0))) return rc;
  // Jacobi preconditioner: 1 / real(diag(B)); a non-positive diagonal entry means B is not positive definite
  const size_t rs = c->dtype == GA_DD ? 16 : 8;
  std::vector<double> dinv((size_t)c->n_pad * (rs / 8), 0.0);
  for (i64 i = 0; i < n; ++i) {
    double dh = 0.0, dl = 0.0;
    bool found = false;
    for (int64_t k = brp[i]; k < brp[i + 1]; ++k) {
      if (bcol[k] != (int32_t)i) continue;
      const double* e = (const double*)((const char*)bval + (size_t)k * c->esize);
      dh = e[0];
      dl = c->dtype == GA_DD ? e[1] : 0.0;
      found = true;
    }
    if (!found || !(dh > 0.0)) return set_err(c, GA_ERR_ARG, "generalised problem: B must have a positive diagonal (B is not positive definite)");
    if (c->dtype == GA_DD) {
      const ddbl inv = ddbl(1.0) / ddbl(dh, dl);
      dinv[2 * (size_t)i] = inv.hi; dinv[2 * (size_t)i + 1] = inv.lo;
    } else {
      dinv[(size_t)i] = 1.0 / dh;
    }
  }
  const size_t vbytes = (size_t)c->n_pad * c->esize;
  GA_CUDA(c, cudaMalloc(&c->gz, vbytes));
  GA_CUDA(c, cudaMalloc(&c->cg_r, vbytes));
  GA_CUDA(c, cudaMalloc(&c->cg_p, vbytes));
  GA_CUDA(c, cudaMalloc(&c->cg_q, vbytes));
  GA_CUDA(c, cudaMalloc(&c->cg_x, vbytes));
  GA_CUDA(c, cudaMemsetAsync(c->cg_x, 0, vbytes, c->stream));
  GA_CUDA(c, cudaMalloc(&c->b_dinv, (size_t)c->n_pad * rs));
  GA_CUDA(c, cudaMemsetAsync(c->gz, 0, vbytes, c->stream));
  GA_CUDA(c, cudaMemsetAsync(c->cg_r, 0, vbytes, c->stream));
  GA_CUDA(c, cudaMemsetAsync(c->cg_p, 0, vbytes, c->stream));
  GA_CUDA(c, cudaMemsetAsync(c->cg_q, 0, vbytes, c->stream));
  GA_CUDA(c, cudaMemcpyAsync(c->b_dinv, dinv.data(), (size_t)c->n_pad * rs, cudaMemcpyHostToDevice, c->stream));
  GA_CUDA(c, cudaStreamSynchronize(c->stream));
  if ((rc = spmv_autotune(c, c->A, c->V, c->gz))) return rc;
  return spmv_autotune(c, c->Bm, c->V, c->gz);
}

// y = inv(B) (A x); ax_out (optional) receives A x, which genopx! leaves in place of x
int general_opx(ga_ctx* c, const void* x, void* y, void* ax_out) {
  const double t0 = gnow();
  int rc = general_spmv(c, c->A, x, c->gz, false);
  if (rc) return rc;
  if ((rc = general_solve(c, c->gz, y))) return rc;
  if (ax_out && ax_out != c->gz)
    GA_CUDA(c, cudaMemcpyAsync(ax_out, c->gz, (size_t)c->n * c->esize, cudaMemcpyDeviceToDevice, c->stream));
  c->stats.tmvopx += gnow() - t0;
  return 0;
}

int general_bx(ga_ctx* c, const void* x, void* y) {
  const double t0 = gnow();
  const int rc = general_spmv(c, c->Bm, x, y, false);
  c->stats.nbx += 1;
  c->stats.tmvbx += gnow() - t0;
  return rc;
}

// s = V[:,1:ncols]' (dot vector gz) ; resid -= V[:,1:ncols] s.  The coefficients stay in ctl->hbuf.
static int project_out(ga_ctx* c, int ncols) {
  c->gs_dotvec = c->gz;
  int rc = launch_gs(c, 0, ncols, -1, DK_TAKE);
  c->gs_dotvec = nullptr;
  if (rc) return rc;
  return launch_gs(c, 2, ncols, -1, DK_NONE);
}

// sqrt|resid' B resid| with gz <- B resid (src/dsaup2.jl:1383-1411; also the tail of every Gram-Schmidt round)
static int resid_bnorm(ga_ctx* c, HReal* out) {
  int rc = general_bx(c, c->resid, c->gz);
  if (rc) return rc;
  return bnorm(c, c->resid, c->gz, out);
}
int general_resid_bnorm(ga_ctx* c, double* rnorm_out) {
  HReal r;
  int rc = set_status_ok(c);
  if (rc) return rc;
  if ((rc = resid_bnorm(c, &r))) return rc;
  rnorm_out[0] = r.hi;
  if (c->dtype == GA_DD) rnorm_out[1] = r.lo;
  return 0;
}

// dgetv0! with bmat = :G (src/dgetv0.jl:504-753)
int general_getv0(ga_ctx* c, int64_t iseed[4], int initv, int j, double* rnorm) {
  const double t0 = gnow();
  int rc = set_status_ok(c);
  if (rc) return rc;
  int ierr = 0;
  if (!initv) {                                                  // :542-553
    i64 seed[4] = {iseed[0], iseed[1], iseed[2], iseed[3]};
    if ((rc = launch_lcg_fill(c, seed))) return rc;
    lcg_advance_seed(seed, (c->dtype == GA_C64 ? 2 : 1) * c->n_global);
    for (int i = 0; i < 4; ++i) iseed[i] = seed[i];
  }
  // force the starting vector into the range of OP (:566-576): resid <- inv(B) A resid
  c->stats.nopx += 1;
  GA_CUDA(c, cudaMemcpyAsync(c->cg_q, c->resid, (size_t)c->n * c->esize, cudaMemcpyDeviceToDevice, c->stream));
  if ((rc = general_opx(c, c->cg_q, c->resid, nullptr))) return rc;
  HReal rnorm0, rn;
  if ((rc = resid_bnorm(c, &rnorm0))) return rc;                 // :596-627
  rn = rnorm0;
  if (j > 1) {
    int iter = 0;
    while (true) {
      if ((rc = project_out(c, j - 1))) return rc;               // :662-665 (coefficients V' B r)
      if ((rc = resid_bnorm(c, &rn))) return rc;                 // :669-698
      if (h_gt_717(c, rn, rnorm0)) break;                        // :711
      iter += 1;
      if (iter <= 5) rnorm0 = rn;                                // :716-720
      else {                                                     // :725-729
        if ((rc = launch_zero_resid(c))) return rc;
        rn = HReal();
        ierr = -1;
        break;
      }
    }
  }
  rnorm[0] = rn.hi;
  if (c->dtype == GA_DD) rnorm[1] = rn.lo;
  c->stats.tgetv0 += gnow() - t0;
  return ierr;
}

// one coefficient of the most recent projection, real part (H[j,2] contributions, src/dsaitr.jl:1243,1367)
static HReal hbuf_real(const ga_ctx* c, int idx) {
  HReal out;
  const DevCtl* hc = c->ctl_host;
  if (c->dtype == GA_F64) out.hi = reinterpret_cast<const double*>(hc->hbuf)[idx];
  else if (c->dtype == GA_C64) out.hi = hc->hbuf[idx].x;
  else { out.hi = hc->hbuf[idx].x; out.lo = hc->hbuf[idx].y; }
  return out;
}
static void h_store(const ga_ctx* c, void* H, int ldh, int row, int col, HReal v) {
  if (c->dtype == GA_DD) {
    double* p = (double*)H + 2 * ((size_t)col * ldh + row);
    p[0] = v.hi; p[1] = v.lo;
  } else {
    ((double*)H)[(size_t)col * ldh + row] = v.hi;
  }
}
static HReal h_load(const ga_ctx* c, const void* H, int ldh, int row, int col) {
  HReal v;
  if (c->dtype == GA_DD) {
    const double* p = (const double*)H + 2 * ((size_t)col * ldh + row);
    v.hi = p[0]; v.lo = p[1];
  } else {
    v.hi = ((const double*)H)[(size_t)col * ldh + row];
  }
  return v;
}
static HReal h_add(const ga_ctx* c, HReal a, HReal b) {
  HReal o;
  if (c->dtype == GA_DD) { const ddbl s = ddbl(a.hi, a.lo) + ddbl(b.hi, b.lo); o.hi = s.hi; o.lo = s.lo; }
  else o.hi = a.hi + b.hi;
  return o;
}

// dsaitr! for mode 2, bmat = :G (src/dsaitr.jl:962-1535).  H is ldh x 2 reals on the host.
int general_lanczos(ga_ctx* c, int k, int np, void* H, int ldh, double* rnorm_io, int64_t iseed[4]) {
  const double t0 = gnow();
  int rc = set_status_ok(c);
  if (rc) return rc;
  HReal rnorm;
  rnorm.hi = rnorm_io[0];
  rnorm.lo = c->dtype == GA_DD ? rnorm_io[1] : 0.0;
  int info = 0;
  for (int j = k + 1; j <= k + np; ++j) {                         // 1-based step, column j-1
    bool rstart = false;
    if (!h_pos(rnorm)) {                                          // :1061-1087, :1497-1525
      c->stats.nrstrt += 1;
      rstart = true;
      int itry = 1;
      double rn[2] = {0.0, 0.0};
      while (true) {
        const int ierr = general_getv0(c, iseed, 0, j, rn);
        if (ierr < 0 && ierr != -1) return ierr;
        if (ierr == -1) {
          itry += 1;
          if (itry <= 3) continue;
          info = j - 1;                                           // :1520
        }
        break;
      }
      if (info > 0) break;
      rnorm.hi = rn[0]; rnorm.lo = c->dtype == GA_DD ? rn[1] : 0.0;
    }
    // STEP 2: v_j = resid / rnorm (:1097-1105)
    const double rn2[2] = {rnorm.hi, rnorm.lo};
    GA_CUDA(c, cudaMemcpyAsync(c->ctl->rnorm, rn2, sizeof(rn2), cudaMemcpyHostToDevice, c->stream));
    if ((rc = launch_normalize(c, j - 1, false))) return rc;
    // STEP 3: resid = inv(B) A v_j, gz = A v_j (:1113-1142, genopx!)
    c->stats.nopx += 1;
    const void* vj = (const char*)c->V + (size_t)(j - 1) * c->n_pad * c->esize;
    if ((rc = general_opx(c, vj, c->resid, nullptr))) return rc;
    // STEP 4: wnorm = sqrt|w' (A v_j)| (:1195-1200); h = V_j' (A v_j) (:1232); r = w - V_j h (:1240)
    HReal wnorm;
    if ((rc = bnorm(c, c->resid, c->gz, &wnorm))) return rc;
    if ((rc = project_out(c, j))) return rc;
    if ((rc = sync_ctl(c))) return rc;
    h_store(c, H, ldh, j - 1, 1, hbuf_real(c, j - 1));            // :1243
    h_store(c, H, ldh, j - 1, 0, (j == 1 || rstart) ? HReal() : rnorm);   // :1259-1263
    const double t4 = gnow();
    if ((rc = resid_bnorm(c, &rnorm))) return rc;                 // :1272-1300
    if (!h_gt_717(c, rnorm, wnorm)) {                             // :1324
      c->stats.nrorth += 1;                                       // :1335 (net of :1439)
      int iter = 0;
      while (true) {
        if ((rc = project_out(c, j))) return rc;                  // :1352-1363: s = V_j' B r ; r -= V_j s
        if ((rc = sync_ctl(c))) return rc;
        if (j == 1 || rstart) h_store(c, H, ldh, j - 1, 0, HReal());                          // :1364-1366
        h_store(c, H, ldh, j - 1, 1, h_add(c, h_load(c, H, ldh, j - 1, 1), hbuf_real(c, j - 1)));  // :1367
        HReal rnorm1;
        if ((rc = resid_bnorm(c, &rnorm1))) return rc;            // :1373-1407
        if (h_gt_717(c, rnorm1, rnorm)) { rnorm = rnorm1; break; }  // :1424-1426
        c->stats.nitref += 1;                                     // :1430
        rnorm = rnorm1;
        iter += 1;
        if (iter <= 1) continue;                                  // :1433-1439
        if ((rc = launch_zero_resid(c))) return rc;               // :1442-1443
        rnorm = HReal();
        break;
      }
    }
    c->stats.titref += gnow() - t4;
    // :1466-1473: H[j,1] is a B-norm (>= 0), so the sign flip never fires
  }
  GA_CUDA(c, cudaStreamSynchronize(c->stream));
  rnorm_io[0] = rnorm.hi;
  if (c->dtype == GA_DD) rnorm_io[1] = rnorm.lo;
  c->stats.taitr += gnow() - t0;
  return info;
}

}  // namespace ga
