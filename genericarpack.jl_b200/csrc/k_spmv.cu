// OP*x of dsaitr step 3 (src/dsaitr.jl:1120-1133 -> opx!(y, ArpackSimpleOp, x) = mul!(y, A, x),
// src/idonow_ops.jl:122; ArpackNormalOp: y = A'(A x), :267-276) as a CSR SpMV.
//
// Row-segment ("CSR-stream") kernel: a CTA owns a block of consecutive rows whose non-zeros form
// ONE contiguous segment of val/col.  The segment is read with fully coalesced loads, multiplied
// with the gathered x entries and parked in shared memory; then one thread per row sums its
// (short) row from shared memory.  Rows longer than the segment capacity get a CTA of their own
// and a block-wide reduction.  Row blocks are computed once on the host when the matrix is
// uploaded.  No atomics: results are bit-reproducible.
#include "ga_decide.cuh"
#include "ga_reduce.cuh"

#include <algorithm>
#include <vector>

namespace ga {

template <class T> struct SpmvCfg {
  static constexpr int THREADS = 256;
  static constexpr int CAP = (sizeof(T) == 8) ? 4096 : 2048;  // products parked per CTA
  static constexpr int MAXROWS = 1024;
};

template <class T>
__global__ void __launch_bounds__(SpmvCfg<T>::THREADS) k_spmv(const int* __restrict__ rowblk,
                                                              const int* __restrict__ rowptr,
                                                              const int* __restrict__ col,
                                                              const T* __restrict__ val, const T* __restrict__ x,
                                                              const T* __restrict__ xh, int nsplit,
                                                              T* __restrict__ y, const DevCtl* ctl, int gated,
                                                              const DevComm* cm) {
  typedef SpmvCfg<T> C;
  if (gated && ctl->status != ST_OK) return;
  if (cm) {
    // fused multi-GPU path: the peers' normalisation kernels push our halo entries; wait for their
    // sequence flags (set after a system-scope fence) before touching xh
    if (threadIdx.x < (unsigned)cm->world && cm->recv_counts[threadIdx.x] > 0)
      spin_until(&cm->peer[cm->rank]->hflags[threadIdx.x], ctl->hseq);
    __syncthreads();
  }
  __shared__ T prod[C::CAP];
  const int r0 = rowblk[blockIdx.x], r1 = rowblk[blockIdx.x + 1];
  const int k0 = rowptr[r0], k1 = rowptr[r1];
  const int tid = threadIdx.x;
  // x entry for (remapped) column c: own rows live in x[0, nsplit), halo entries received from
  // the peers in xh (single GPU: nsplit = INT_MAX, xh unused)
  auto xat = [&](int c) -> T { return c < nsplit ? x[c] : xh[c - nsplit]; };
  if (k1 - k0 <= C::CAP) {
    // coalesced segment pass, 4 independent gathers in flight per thread
    int k = k0 + tid;
    for (; k + 3 * C::THREADS < k1; k += 4 * C::THREADS) {
      const int c0 = col[k], c1 = col[k + C::THREADS], c2 = col[k + 2 * C::THREADS], c3 = col[k + 3 * C::THREADS];
      const T v0 = val[k], v1 = val[k + C::THREADS], v2 = val[k + 2 * C::THREADS], v3 = val[k + 3 * C::THREADS];
      const T x0 = xat(c0), x1 = xat(c1), x2 = xat(c2), x3 = xat(c3);
      prod[k - k0] = Elem<T>::mul(v0, x0);
      prod[k - k0 + C::THREADS] = Elem<T>::mul(v1, x1);
      prod[k - k0 + 2 * C::THREADS] = Elem<T>::mul(v2, x2);
      prod[k - k0 + 3 * C::THREADS] = Elem<T>::mul(v3, x3);
    }
    for (; k < k1; k += C::THREADS) prod[k - k0] = Elem<T>::mul(val[k], xat(col[k]));
    __syncthreads();
    for (int r = r0 + tid; r < r1; r += C::THREADS) {
      const int a = rowptr[r] - k0, b = rowptr[r + 1] - k0;
      T s = Elem<T>::zero();
      for (int q = a; q < b; ++q) Elem<T>::add(s, prod[q]);
      y[r] = s;
    }
  } else {
    // one long row for the whole CTA (r1 == r0 + 1 by construction)
    T s = Elem<T>::zero();
    for (int k = k0 + tid; k < k1; k += C::THREADS) Elem<T>::add(s, Elem<T>::mul(val[k], xat(col[k])));
    prod[tid] = s;
    __syncthreads();
    for (int off = C::THREADS / 2; off > 0; off >>= 1) {
      if (tid < off) { T a = prod[tid]; Elem<T>::add(a, prod[tid + off]); prod[tid] = a; }
      __syncthreads();
    }
    if (tid == 0) y[r0] = prod[0];
  }
}

template <class T>
static void spmv_launch_t(ga_ctx* c, const CsrDev& M, const void* x, void* y, bool gated, bool wait_halo) {
  if (M.nblk <= 0) return;
  const bool split = c->world > 1 && &M == &c->A && c->op_kind == 1;
  k_spmv<T><<<M.nblk, SpmvCfg<T>::THREADS, 0, c->stream>>>(M.rowblk, M.rowptr, M.col, (const T*)M.val, (const T*)x,
                                                          (const T*)c->halo, split ? (int)c->n_pad : 2147483647,
                                                          (T*)y, c->ctl, gated ? 1 : 0,
                                                          (split && wait_halo) ? c->devcomm : nullptr);
}
static int spmv_launch(ga_ctx* c, const CsrDev& M, const void* x, void* y, bool gated, bool wait_halo = false) {
  switch (c->dtype) {
    case GA_F64: spmv_launch_t<double>(c, M, x, y, gated, wait_halo); break;
    case GA_C64: spmv_launch_t<cdbl>(c, M, x, y, gated, wait_halo); break;
    default: spmv_launch_t<ddbl>(c, M, x, y, gated, wait_halo); break;
  }
  GA_CUDA(c, cudaGetLastError());
  c->launches += 1;
  return 0;
}

int comm_halo_exchange(ga_ctx* c, const void* x_local);  // ga_comm.cu

// y = OP x for device vectors of the (local) problem size.
int launch_opx(ga_ctx* c, const void* x, void* y, bool gated) {
  if (c->op_kind == 0) return set_err(c, GA_ERR_NOOP, "no operator uploaded (ga_set_csr)");
  const void* xin = x;
  // Lanczos steps (gated) on the fused path: the normalisation kernel already pushed the halo and
  // the SpMV waits on the peers' flags.  Otherwise the halo travels by NCCL send/recv first.
  const bool fused = gated && c->world > 1 && c->devcomm != nullptr;
  if (c->world > 1 && !fused) {
    int rc = comm_halo_exchange(c, x);
    if (rc) return rc;
  }
  if (c->op_kind == 1) return spmv_launch(c, c->A, xin, y, gated, fused);
  // normal operator (single GPU): left: y = A^H (A x); right: y = A (A^H x)
  if (c->normal_left) {
    int rc = spmv_launch(c, c->A, xin, c->xfull, gated);
    if (rc) return rc;
    return spmv_launch(c, c->At, c->xfull, y, gated);
  } else {
    int rc = spmv_launch(c, c->At, xin, c->xfull, gated);
    if (rc) return rc;
    return spmv_launch(c, c->A, c->xfull, y, gated);
  }
}

// Upload a host CSR (int64 rowptr, int32 col, T val) and build the row blocks.
int csr_upload(ga_ctx* c, CsrDev& M, i64 nrows, i64 ncols, const int64_t* rowptr, const int32_t* col,
               const void* val, i64 col_shift) {
  M.free_all();
  const i64 nnz = rowptr[nrows] - rowptr[0];
  if (nnz >= (i64)2147483647LL || nrows >= (i64)2147483647LL)
    return set_err(c, GA_ERR_ARG, "csr: more than 2^31-1 non-zeros or rows per rank is not supported");
  const int cap = c->esize == 8 ? SpmvCfg<double>::CAP : SpmvCfg<cdbl>::CAP;
  const int maxrows = SpmvCfg<double>::MAXROWS;
  std::vector<int> rp((size_t)nrows + 1);
  const int64_t base = rowptr[0];
  for (i64 i = 0; i <= nrows; ++i) rp[(size_t)i] = (int)(rowptr[i] - base);
  std::vector<int> blk;
  blk.reserve((size_t)(nnz / (cap / 2) + nrows / maxrows + 16));
  i64 r = 0;
  while (r < nrows) {
    blk.push_back((int)r);
    i64 e = r + 1;  // a row always fits "alone" (long-row path if needed)
    while (e < nrows && (e - r) < maxrows && (rp[(size_t)e + 1] - rp[(size_t)r]) <= cap) ++e;
    r = e;
  }
  blk.push_back((int)nrows);
  M.nrows = nrows; M.ncols = ncols; M.nnz = nnz; M.nblk = (int)blk.size() - 1;
  GA_CUDA(c, cudaMalloc(&M.rowptr, sizeof(int) * ((size_t)nrows + 1)));
  GA_CUDA(c, cudaMalloc(&M.col, sizeof(int) * (size_t)std::max<i64>(nnz, 1)));
  GA_CUDA(c, cudaMalloc(&M.val, c->esize * (size_t)std::max<i64>(nnz, 1)));
  GA_CUDA(c, cudaMalloc(&M.rowblk, sizeof(int) * blk.size()));
  GA_CUDA(c, cudaMemcpyAsync(M.rowptr, rp.data(), sizeof(int) * rp.size(), cudaMemcpyHostToDevice, c->stream));
  GA_CUDA(c, cudaMemcpyAsync(M.rowblk, blk.data(), sizeof(int) * blk.size(), cudaMemcpyHostToDevice, c->stream));
  if (col_shift == 0) {
    GA_CUDA(c, cudaMemcpyAsync(M.col, col + base, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice, c->stream));
  } else {
    std::vector<int> cc((size_t)nnz);
    for (i64 k = 0; k < nnz; ++k) cc[(size_t)k] = (int)(col[base + k] - col_shift);
    GA_CUDA(c, cudaMemcpyAsync(M.col, cc.data(), sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice, c->stream));
    GA_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  GA_CUDA(c, cudaMemcpyAsync(M.val, (const char*)val + (size_t)base * c->esize, c->esize * (size_t)nnz,
                             cudaMemcpyHostToDevice, c->stream));
  GA_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}

}  // namespace ga
