// OP*x of dsaitr step 3 (src/dsaitr.jl:1120-1133 -> opx!(y, ArpackSimpleOp, x) = mul!(y, A, x),
// src/idonow_ops.jl:122; ArpackNormalOp: y = A'(A x), :267-276) as a CSR SpMV.
//
// Row-segment ("CSR-stream") kernel: a CTA owns a block of consecutive rows whose non-zeros form
// ONE contiguous segment of val/col.  The segment is read with fully coalesced loads, multiplied
// with the gathered x entries and parked in shared memory; then one thread per row sums its
// (short) row from shared memory.  Rows longer than the segment capacity get a CTA of their own
// and a block-wide reduction.  Row blocks are computed once on the host when the matrix is
// uploaded.  No atomics: results are bit-reproducible.
//
// k_spmv_stream (default): the same row-segment idea as a persistent, warp-specialised kernel.  The
// segment of a row block (col, val and the rowptr slice: three contiguous byte ranges) is pulled
// into a multi-stage shared-memory ring by TMA bulk copies (cp.async.bulk, UBLKCP) issued by a
// producer warp, so HBM streaming never waits for the dependent x gathers; four consumer groups
// of 128 threads take the stages round-robin: gather x, multiply in place, group barrier, one
// thread per row sums its row.  Identical arithmetic (products rounded, then summed in CSR
// order), bit-identical to the first kernel.  Measured on B200 (profiles/): k_spmv 3.66 TB/s ->
// see profiles/RESULTS.md for the streamed kernel.
#include "ga_decide.cuh"
#include "ga_reduce.cuh"
#include "ga_tma.cuh"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
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
                                                              const T* xh, int nsplit,
                                                              T* __restrict__ y, DevCtl* ctl, int gated,
                                                              const DevComm* cm, int accumulate) {
  typedef SpmvCfg<T> C;
  if (gated && ctl->status != ST_OK) return;
  if (cm) {
    // fused multi-GPU path: the peers' normalisation kernels push our halo entries; wait for their
    // sequence flags (set after a system-scope fence) before touching xh
    if (threadIdx.x < (unsigned)cm->world && cm->recv_counts[threadIdx.x] > 0)
      if (!spin_until(&cm->peer[cm->rank]->hflags[threadIdx.x], ctl->hseq)) ctl->status = ST_COMM;
    __syncthreads();
  }
  __shared__ T prod[C::CAP];
  const int r0 = rowblk[blockIdx.x], r1 = rowblk[blockIdx.x + 1];
  const int k0 = rowptr[r0], k1 = rowptr[r1];
  const int tid = threadIdx.x;
  // x entry for (remapped) column c: own rows live in x[0, nsplit), halo entries received from
  // the peers in xh (single GPU: nsplit = INT_MAX, xh unused)
  auto xat = [&](int c) -> T { return c < nsplit ? x[c] : ld_halo(xh + (c - nsplit)); };
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
      T s = accumulate ? y[r] : Elem<T>::zero();      // column-blocked operators: continue the row's running sum
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
    if (tid == 0) {
      T a = prod[0];
      if (accumulate) { T b = y[r0]; Elem<T>::add(b, a); a = b; }
      y[r0] = a;
    }
  }
}


// ---------------------------------------------------------------------------------------------
// Streamed SpMV: persistent CTAs, TMA-staged row blocks.
template <class T, int GROUPS_ = 4> struct SpmvStreamCfg {
  static constexpr int CAP = (sizeof(T) <= 8) ? 2048 : 1024;  // non-zeros per row block (one pipeline stage)
  static constexpr int MAXROWS = 512;    // rows per row block
  static constexpr int GROUPS = GROUPS_, GT = 128;
  static constexpr int THREADS = GROUPS * GT + 32;      // consumer groups + producer warp
  static constexpr int COL_BYTES = (CAP + 8) * 4;       // +8: 16-byte alignment slack of the segment
  static constexpr int VAL_BYTES = (CAP + 8) * (int)sizeof(T);
  static constexpr int RP_BYTES = (MAXROWS + 8) * 4;
  static constexpr int STAGE_BYTES = COL_BYTES + VAL_BYTES + RP_BYTES;
  // 6 stages (160 KB / 136 KB): measured on B200, 4 stages lose 20 % on the Laplacian (pipeline too
  // shallow) while 8 stages (214 KB) leave ~14 KB of L1 and halve the rate of gather-heavy matrices
  // (the L1 lines double as miss buffers for the x gathers); see profiles/RESULTS.md
  //
  // Ring protocol.  The stages form ONE ring shared by all consumer groups, so unless GROUPS divides
  // the stage count a group observes only some of the phases of a stage's `full` barrier, and a
  // parity wait alone is ambiguous: if the previous fill of the stage is still in flight when the
  // group arrives, the awaited parity equals that of the phase before it and the wait falls
  // through.  That happens on B200 (bulk copies do NOT complete in issue order: with (groups,
  // stages) = (4, 5) / (6, 8) the complex stencil gave wrong products once per ~10^6 row blocks, or
  // read a never-filled stage), so the producer also publishes the ring iteration held by a stage
  // (`seq[s]`, written before it arms the barrier) and a consumer re-waits until the stage holds ITS
  // iteration (wait_stage below).  Any stage count >= GROUPS is then correct.
  static constexpr int MAX_STAGES = 8;
  static constexpr int STAGES = 6;
  static constexpr int BAR_BYTES = 2 * MAX_STAGES * 8 + MAX_STAGES * 4 + 16;   // full[], empty[], seq[]
  static constexpr int smem_bytes(int nst) { return nst * STAGE_BYTES + BAR_BYTES; }
};

// Block until stage `s` holds ring iteration `it` (phase parity `ph` of its `full` barrier).  While an
// OLDER fill of the stage is still in flight the parity wait falls through and `seq` names that older
// iteration: poll (the window is the tail of one bulk copy) until the barrier has moved on; from then
// on the parity wait blocks properly.  No second blocking wait on the opposite parity: the barrier may
// advance two phases between the two instructions, which would park this group forever.
__device__ __forceinline__ void wait_stage(uint64_t* full_s, const volatile int* seq_s, uint32_t ph, int it) {
  for (;;) {
    mbar_wait(full_s, ph);
    if (*seq_s == it) return;       // armed for `it` by the producer, and complete
    __nanosleep(32);
  }
}

template <class T, int NG>
__global__ void __launch_bounds__(SpmvStreamCfg<T, NG>::THREADS, 1)
k_spmv_stream(const int4* __restrict__ desc, int ndesc, const int* __restrict__ rowptr, const int* __restrict__ col,
              const T* __restrict__ val, const T* __restrict__ x, const T* xh, int nsplit,
              T* __restrict__ y, DevCtl* ctl, int gated, const DevComm* cm, int NST, unsigned int* chk, int accumulate) {
  typedef SpmvStreamCfg<T, NG> C;
  constexpr int GT = C::GT, GROUPS = C::GROUPS;
  if (gated && ctl->status != ST_OK) return;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)NST * C::STAGE_BYTES);
  uint64_t* empty = full + C::MAX_STAGES;
  volatile int* seq = reinterpret_cast<volatile int*>(empty + C::MAX_STAGES);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // contiguous chunk of row blocks per CTA: consecutive blocks share most of their x window, so the
  // four consumer groups of a CTA (working on adjacent blocks) hit the same L1 lines
  const int per = ndesc / (int)gridDim.x, rem = ndesc % (int)gridDim.x;
  const int b_begin = (int)blockIdx.x * per + ((int)blockIdx.x < rem ? (int)blockIdx.x : rem);
  const int b_end = b_begin + per + ((int)blockIdx.x < rem ? 1 : 0);
  if (tid == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], GT / 32); seq[s] = -1; }
    fence_mbar_init();
  }
  __syncthreads();

  if (warp == GROUPS * GT / 32) {
    // ---------------- producer warp (one elected lane): three bulk copies per row block
    if (lane == 0) {
      int it = 0;
      int b = b_begin;
      int4 d = b < b_end ? __ldg(&desc[b]) : make_int4(0, 0, 0, 0);
      for (; b < b_end; ++b, ++it) {
        const int bn = b + 1;
        const int4 dn = bn < b_end ? __ldg(&desc[bn]) : make_int4(0, 0, 0, 0);  // next descriptor in flight
        const int s = it % NST;
        const uint32_t ph = (uint32_t)(it / NST) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);
        unsigned char* st = smem + (size_t)s * C::STAGE_BYTES;
        const int k0a = d.z & ~3, k1a = (d.w + 3) & ~3;
        const int r0a = d.x & ~3, nrp = (d.y - r0a + 1 + 3) & ~3;
        const uint32_t nk = (uint32_t)(k1a - k0a);
        seq[s] = it;   // published by the release of the arrive below
        mbar_arrive_expect_tx(&full[s], nk * (4u + (uint32_t)sizeof(T)) + (uint32_t)nrp * 4u);
        if (nk > 0) {
          bulk_g2s(st, col + k0a, nk * 4u, &full[s]);
          bulk_g2s(st + C::COL_BYTES, val + k0a, nk * (uint32_t)sizeof(T), &full[s]);
        }
        bulk_g2s(st + C::COL_BYTES + C::VAL_BYTES, rowptr + r0a, (uint32_t)nrp * 4u, &full[s]);
        d = dn;
      }
    }
    return;
  }

  // ---------------- consumer group g takes row blocks g, g + GROUPS, ... of this CTA's sequence
  const int g = warp / (GT / 32), gt = tid - g * GT;
  if (cm) {
    // fused multi-GPU path: the peers' normalisation kernels push our halo entries; wait for their
    // sequence flags (set after a system-scope fence) before touching xh (the producer warp is
    // already streaming the matrix meanwhile)
    if (gt < cm->world && cm->recv_counts[gt] > 0 && !spin_until(&cm->peer[cm->rank]->hflags[gt], ctl->hseq)) ctl->status = ST_COMM;
    named_bar_sync(1 + g, GT);
  }
  auto xat = [&](int c) -> T { return c < nsplit ? x[c] : ld_halo(xh + (c - nsplit)); };
  int b = b_begin + g;
  int4 d = b < b_end ? __ldg(&desc[b]) : make_int4(0, 0, 0, 0);
  for (int it = g; b < b_end; b += GROUPS, it += GROUPS) {
    const int bn = b + GROUPS;
    const int4 dn = bn < b_end ? __ldg(&desc[bn]) : make_int4(0, 0, 0, 0);
    const int s = it % NST;
    const uint32_t ph = (uint32_t)(it / NST) & 1u;
    unsigned char* st = smem + (size_t)s * C::STAGE_BYTES;
    const int k0 = d.z, cnt = d.w - d.z, sh = k0 - (k0 & ~3);
    const int* cs = reinterpret_cast<const int*>(st) + sh;
    T* vs = reinterpret_cast<T*>(st + C::COL_BYTES) + sh;
    const int* rp = reinterpret_cast<const int*>(st + C::COL_BYTES + C::VAL_BYTES) + (d.x - (d.x & ~3));
    wait_stage(&full[s], &seq[s], ph, it);
    if (chk != nullptr && (gt & 31) == 0) {
      // debug self-check (ga_debug_spmv_check): the staged rowptr slice must be the one of THIS row block
      const int f0 = rp[0], f1 = rp[d.y - d.x];
      if (f0 != d.z || f1 != d.w) {
        const unsigned int i = atomicAdd(&chk[0], 1u);
        if (i < 15u) {
          unsigned int* o = chk + 8 + 8 * i;
          o[0] = blockIdx.x; o[1] = (unsigned)it; o[2] = (unsigned)(b - b_begin); o[3] = (unsigned)d.z; o[4] = (unsigned)f0;
          o[5] = (unsigned)d.w; o[6] = (unsigned)f1; o[7] = (unsigned)(NST * 16 + GROUPS);
        }
      }
    }
    // products in place: vs[k] <- val[k] * x[col[k]], 8 independent gathers in flight per thread
    for (int k = gt; k < cnt; k += 8 * GT) {
      int cc[8];
      T xx[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) cc[u] = (k + u * GT < cnt) ? cs[k + u * GT] : -1;
#pragma unroll
      for (int u = 0; u < 8; ++u) xx[u] = cc[u] >= 0 ? xat(cc[u]) : Elem<T>::zero();
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (cc[u] >= 0) vs[k + u * GT] = Elem<T>::mul(vs[k + u * GT], xx[u]);
    }
    named_bar_sync(1 + g, GT);
    const int nr = d.y - d.x;
    for (int r = gt; r < nr; r += GT) {
      const int a = rp[r] - k0, e = rp[r + 1] - k0;
      T sum = accumulate ? y[d.x + r] : Elem<T>::zero();   // column-blocked operators: continue the row's running sum
      for (int q = a; q < e; ++q) Elem<T>::add(sum, vs[q]);
      y[d.x + r] = sum;
    }
    // the stage was written through the generic proxy; the next TMA copy into it is async-proxy
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
    d = dn;
  }
}

// rows longer than one stage: one CTA per row, block-wide reduction (same order as the v1 path)
template <class T>
__global__ void __launch_bounds__(256) k_spmv_long(const int* __restrict__ longrow, const int* __restrict__ rowptr,
                                                   const int* __restrict__ col, const T* __restrict__ val,
                                                   const T* __restrict__ x, const T* xh, int nsplit,
                                                   T* __restrict__ y, DevCtl* ctl, int gated, const DevComm* cm,
                                                   int accumulate) {
  if (gated && ctl->status != ST_OK) return;
  if (cm) {
    if (threadIdx.x < (unsigned)cm->world && cm->recv_counts[threadIdx.x] > 0)
      if (!spin_until(&cm->peer[cm->rank]->hflags[threadIdx.x], ctl->hseq)) ctl->status = ST_COMM;
    __syncthreads();
  }
  __shared__ T red[256];
  const int r = longrow[blockIdx.x], k0 = rowptr[r], k1 = rowptr[r + 1], tid = threadIdx.x;
  auto xat = [&](int c) -> T { return c < nsplit ? x[c] : ld_halo(xh + (c - nsplit)); };
  T s = Elem<T>::zero();
  for (int k = k0 + tid; k < k1; k += 256) Elem<T>::add(s, Elem<T>::mul(val[k], xat(col[k])));
  red[tid] = s;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if (tid < off) { T a = red[tid]; Elem<T>::add(a, red[tid + off]); red[tid] = a; }
    __syncthreads();
  }
  if (tid == 0) {
    T a = red[0];
    if (accumulate) { T b = y[r]; Elem<T>::add(b, a); a = b; }
    y[r] = a;
  }
}

// tuning overrides (GA_SPMV_STAGES, GA_SPMV_GROUPS); 0 = use the per-matrix autotuned values
static int g_spmv_force_stages = -1, g_spmv_force_groups = -1;
static void spmv_read_env() {
  if (g_spmv_force_stages >= 0) return;
  g_spmv_force_stages = 0; g_spmv_force_groups = 0;
  if (const char* e = getenv("GA_SPMV_STAGES")) { const int v = atoi(e); if (v >= 4 && v <= 8) g_spmv_force_stages = v; }
  if (const char* e = getenv("GA_SPMV_GROUPS")) { const int v = atoi(e); if (v == 4 || v == 6) g_spmv_force_groups = v; }
}

// does M's column space have the sharded layout [own rows: n_pad | halo]?  (the sharded simple / normal operator's A, and
// both matrices of the sharded generalised problem)
static bool is_split(const ga_ctx* c, const CsrDev& M) {
  if (c->world <= 1) return false;
  if (c->general_dist) return &M == &c->A || &M == &c->Bm;
  return &M == &c->A && (c->op_kind == 1 || c->normal_dist);
}

template <class T, int NG>
static int spmv_stream_launch_g(ga_ctx* c, const CsrDev& M, const void* x, void* y, bool gated, int nsplit,
                                const DevComm* cm, int nst, bool accumulate) {
  typedef SpmvStreamCfg<T, NG> C;
  static PerDeviceOnce attr_set;
  if (attr_set.need(c->device)) {
    GA_CUDA(c, cudaFuncSetAttribute(k_spmv_stream<T, NG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    C::smem_bytes(C::MAX_STAGES)));
    attr_set.done(c->device);
  }
  if (nst < NG) nst = NG;
  const int grid = M.ndesc < c->sm_count ? M.ndesc : c->sm_count;
  k_spmv_stream<T, NG><<<grid, C::THREADS, C::smem_bytes(nst), c->stream>>>(M.desc, M.ndesc, M.rowptr, M.col,
                                                                           (const T*)M.val, (const T*)x,
                                                                           (const T*)c->halo, nsplit, (T*)y, c->ctl,
                                                                           gated ? 1 : 0, cm, nst, c->spmv_chk, accumulate ? 1 : 0);
  GA_CUDA(c, cudaGetLastError());
  c->launches += 1;
  return 0;
}

template <class T>
static int spmv_stream_launch_t(ga_ctx* c, const CsrDev& M, const void* x, void* y, bool gated, bool wait_halo, bool accumulate) {
  typedef SpmvStreamCfg<T> C;
  const bool split = is_split(c, M);
  const int nsplit = split ? (int)c->n_pad : 2147483647;
  const DevComm* cm = (split && wait_halo) ? c->devcomm : nullptr;
  spmv_read_env();
  int nst = M.tune_stages, ngroups = M.tune_groups;
  if (g_spmv_force_stages > 0) nst = g_spmv_force_stages;
  if (g_spmv_force_groups > 0) ngroups = g_spmv_force_groups;
  if (nst > C::MAX_STAGES) nst = C::MAX_STAGES;
  if (M.ndesc > 0) {
    int rc = ngroups == 6 ? spmv_stream_launch_g<T, 6>(c, M, x, y, gated, nsplit, cm, nst, accumulate)
                          : spmv_stream_launch_g<T, 4>(c, M, x, y, gated, nsplit, cm, nst, accumulate);
    if (rc) return rc;
  }
  if (M.nlong > 0) {
    k_spmv_long<T><<<M.nlong, 256, 0, c->stream>>>(M.longrow, M.rowptr, M.col, (const T*)M.val, (const T*)x,
                                                  (const T*)c->halo, nsplit, (T*)y, c->ctl, gated ? 1 : 0, cm, accumulate ? 1 : 0);
    GA_CUDA(c, cudaGetLastError());
    c->launches += 1;
  }
  return 0;
}

template <class T>
static void spmv_launch_t(ga_ctx* c, const CsrDev& M, const void* x, void* y, bool gated, bool wait_halo, bool accumulate) {
  if (M.nblk <= 0) return;
  const bool split = is_split(c, M);
  k_spmv<T><<<M.nblk, SpmvCfg<T>::THREADS, 0, c->stream>>>(M.rowblk, M.rowptr, M.col, (const T*)M.val, (const T*)x,
                                                          (const T*)c->halo, split ? (int)c->n_pad : 2147483647,
                                                          (T*)y, c->ctl, gated ? 1 : 0,
                                                          (split && wait_halo) ? c->devcomm : nullptr, accumulate ? 1 : 0);
}
static int spmv_launch(ga_ctx* c, const CsrDev& M, const void* x, void* y, bool gated, bool wait_halo = false, bool accumulate = false) {
  if (c->use_spmv_stream) {
    switch (c->dtype) {
      case GA_F64: return spmv_stream_launch_t<double>(c, M, x, y, gated, wait_halo, accumulate);
      case GA_C64: return spmv_stream_launch_t<cdbl>(c, M, x, y, gated, wait_halo, accumulate);
      case GA_F32: return spmv_stream_launch_t<float>(c, M, x, y, gated, wait_halo, accumulate);
      default: return spmv_stream_launch_t<ddbl>(c, M, x, y, gated, wait_halo, accumulate);
    }
  }
  switch (c->dtype) {
    case GA_F64: spmv_launch_t<double>(c, M, x, y, gated, wait_halo, accumulate); break;
    case GA_C64: spmv_launch_t<cdbl>(c, M, x, y, gated, wait_halo, accumulate); break;
    case GA_F32: spmv_launch_t<float>(c, M, x, y, gated, wait_halo, accumulate); break;
    default: spmv_launch_t<ddbl>(c, M, x, y, gated, wait_halo, accumulate); break;
  }
  GA_CUDA(c, cudaGetLastError());
  c->launches += 1;
  return 0;
}

int launch_spmv_mat(ga_ctx* c, const CsrDev& M, const void* x, void* y, bool gated, bool wait_halo) { return spmv_launch(c, M, x, y, gated, wait_halo); }

// ---------------------------------------------------------------------------------------------
// Sharded normal operator OP = T^H T: after z = T_loc^H (T_loc x) every rank owns a full-length
// partial result in its [own | halo] layout; the halo segment "owned by rank q" is rank q's share.
// k_zsignal publishes "my partial result is complete" to the peers (fused path); k_reduce_scatter
// adds, in rank order, the peers' segments to the own part.  src[q] points at the segment rank q
// holds for this rank: in q's memory over NVLink (fused) or in the receive buffer (NCCL path).
__global__ void k_zsignal(DevCtl* ctl, const DevComm* cm, int gated) {
  if (gated && ctl->status != ST_OK) return;
  __threadfence_system();
  const unsigned int zs = ctl->zseq + 1u;
  __syncthreads();
  if ((int)threadIdx.x < cm->world && (int)threadIdx.x != cm->rank) st_sys_u32(&cm->peer[threadIdx.x]->zflags[cm->rank], zs);
  if (threadIdx.x == 0) ctl->zseq = zs;
}

struct ZSrc { const void* p[kMaxWorld]; };

template <class T>
__global__ void __launch_bounds__(256) k_reduce_scatter(T* __restrict__ out, const T* __restrict__ own, ZSrc src,
                                                        const int* __restrict__ inv, i64 n, i64 n_pad, int world,
                                                        int rank, DevCtl* ctl, int gated, const DevComm* cm) {
  if (gated && ctl->status != ST_OK) return;
  if (cm) {  // wait until every peer's partial result is complete (flags written after a system-scope fence)
    if ((int)threadIdx.x < world && (int)threadIdx.x != rank) {
      if (!spin_until(&cm->peer[rank]->zflags[threadIdx.x], ctl->zseq)) ctl->status = ST_COMM;
    }
    __syncthreads();
  }
  const i64 stride = (i64)gridDim.x * blockDim.x;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    T acc = own[i];
    for (int q = 0; q < world; ++q) {
      if (q == rank) continue;
      const int p = inv[(size_t)q * n_pad + i];
      if (p >= 0) {
        const T* sq = reinterpret_cast<const T*>(src.p[q]);
        T v;
        if (sizeof(T) == 4) {
          float f;
          asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(f) : "l"(sq + p));
          *reinterpret_cast<float*>(&v) = f;
        } else if (sizeof(T) == 8) {
          double d;
          asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(d) : "l"(sq + p));
          *reinterpret_cast<double*>(&v) = d;
        } else {
          double2 d = ld_sys_d2(reinterpret_cast<const double2*>(sq + p));
          *reinterpret_cast<double2*>(&v) = d;
        }
        Elem<T>::add(acc, v);
      }
    }
    out[i] = acc;
  }
}

int comm_zp_exchange(ga_ctx* c, const void* zp);  // ga_comm.cu (NCCL path)

static int launch_normal_dist(ga_ctx* c, const void* x, void* y, bool gated, bool fused) {
  // (1) the halo of x is in place (pushed by the normalisation kernel, or exchanged by the caller)
  int rc = spmv_launch(c, c->A, x, c->xfull, gated, fused);               // (2) y_loc = T_loc [x_own | halo]
  if (rc) return rc;
  void* zp = c->zp[c->zpar & 1u];
  rc = spmv_launch(c, c->At, c->xfull, zp, gated);                        // (3) z = T_loc^H y_loc
  if (rc) return rc;
  const bool peer = c->devcomm != nullptr;                                // (4) reduce-scatter
  ZSrc src;
  const size_t zoff = (size_t)((const char*)zp - (const char*)c->halo);
  size_t soff = 0;
  for (int q = 0; q < c->world; ++q) {
    if (peer) src.p[q] = (const char*)c->hostcomm.peer_halo[q] + zoff + ((size_t)c->n_pad + (size_t)c->dest_off[q]) * c->esize;
    else src.p[q] = (const char*)c->zrecv + soff * c->esize;
    soff += (size_t)c->send_counts[q];
  }
  if (peer) {
    k_zsignal<<<1, 32, 0, c->stream>>>(c->ctl, c->devcomm, gated ? 1 : 0);
    GA_CUDA(c, cudaGetLastError());
    c->launches += 1;
  } else if ((rc = comm_zp_exchange(c, zp))) return rc;
  const int threads = 256;
  i64 blocks = (c->n + threads - 1) / threads;
  if (blocks > (i64)c->sm_count * 8) blocks = (i64)c->sm_count * 8;
  if (blocks < 1) blocks = 1;
  const DevComm* cm = peer ? c->devcomm : nullptr;
  switch (c->dtype) {
    case GA_F32: k_reduce_scatter<float><<<(int)blocks, threads, 0, c->stream>>>((float*)y, (const float*)zp, src, c->zinv, c->n, c->n_pad, c->world, c->rank, c->ctl, gated ? 1 : 0, cm); break;
    case GA_F64: k_reduce_scatter<double><<<(int)blocks, threads, 0, c->stream>>>((double*)y, (const double*)zp, src, c->zinv, c->n, c->n_pad, c->world, c->rank, c->ctl, gated ? 1 : 0, cm); break;
    case GA_C64: k_reduce_scatter<cdbl><<<(int)blocks, threads, 0, c->stream>>>((cdbl*)y, (const cdbl*)zp, src, c->zinv, c->n, c->n_pad, c->world, c->rank, c->ctl, gated ? 1 : 0, cm); break;
    default: k_reduce_scatter<ddbl><<<(int)blocks, threads, 0, c->stream>>>((ddbl*)y, (const ddbl*)zp, src, c->zinv, c->n, c->n_pad, c->world, c->rank, c->ctl, gated ? 1 : 0, cm); break;
  }
  GA_CUDA(c, cudaGetLastError());
  c->launches += 1;
  c->zpar ^= 1u;
  return 0;
}

int comm_halo_exchange(ga_ctx* c, const void* x_local);  // ga_comm.cu

// the caller's operator (function ops): enqueue on the context's stream
static int call_op_fn(ga_ctx* c, ga_op_fn fn, const void* x, void* y, i64 nx, i64 ny) {
  const int rc = fn(c->op_user, x, y, (int64_t)nx, (int64_t)ny, (void*)c->stream);
  if (rc) return set_err(c, GA_ERR_ARG, "the operator callback returned " + std::to_string(rc));
  c->launches += 1;
  return 0;
}
// one half of the normal operator: y = A x (adjoint = false, x: n entries, y: m) or y = A^H x (x: m, y: n)
static int normal_half(ga_ctx* c, bool adjoint, const void* x, void* y, bool gated) {
  if (c->av_cb)
    return adjoint ? call_op_fn(c, c->atv_cb, x, y, c->normal_m, c->normal_n) : call_op_fn(c, c->av_cb, x, y, c->normal_n, c->normal_m);
  return spmv_launch(c, adjoint ? c->At : c->A, x, y, gated);
}

// Second product of the normal operator, y = W z with the WIDE matrix W (A^H for a tall A, A for a wide one) and the
// long intermediate z.  Its x gathers are random over z: at C5's size (z = 50M doubles) almost every gather is its own
// 32-byte DRAM sector and the product runs at ~0.4 of the HBM peak.  W is therefore stored in COLUMN BLOCKS whose slice
// of z stays L2-resident (ga_set_normal_csr); the blocks are applied in ascending column order, each continuing the
// row's running sum, so every y_i sees the same additions in the same order as the unblocked row sum: same bits.
static int normal_wide(ga_ctx* c, const void* z, void* y, bool gated) {
  if (c->Wblk.empty()) return spmv_launch(c, c->normal_left ? c->At : c->A, z, y, gated);
  for (size_t b = 0; b < c->Wblk.size(); ++b) {
    int rc = spmv_launch(c, c->Wblk[b], (const char*)z + (size_t)b * c->wblk_cols * c->esize, y, gated, false, b > 0);
    if (rc) return rc;
  }
  return 0;
}

// y = OP x for device vectors of the (local) problem size.
int launch_opx(ga_ctx* c, const void* x, void* y, bool gated) {
  if (c->op_kind == 0) return set_err(c, GA_ERR_NOOP, "no operator uploaded (ga_set_csr)");
  const void* xin = x;
  if (c->op_kind == 3) return general_opx(c, xin, y, nullptr);   // moves its own halos (k_general.cu)
  // Lanczos steps (gated) on the fused path: the normalisation kernel already pushed the halo and
  // the SpMV waits on the peers' flags.  Otherwise the halo travels by NCCL send/recv first.
  const bool fused = gated && c->world > 1 && c->devcomm != nullptr;
  if (c->world > 1 && !fused) {
    int rc = comm_halo_exchange(c, x);
    if (rc) return rc;
  }
  if (c->op_kind == 1 && c->op_cb) return call_op_fn(c, c->op_cb, xin, y, c->n, c->n);
  if (c->op_kind == 1) return spmv_launch(c, c->A, xin, y, gated, fused);
  if (c->normal_dist) return launch_normal_dist(c, xin, y, gated, fused);
  // normal operator (single GPU): left: y = A^H (A x); right: y = A (A^H x)
  if (c->normal_left) {
    int rc = normal_half(c, false, xin, c->xfull, gated);
    if (rc) return rc;
    return c->av_cb ? normal_half(c, true, c->xfull, y, gated) : normal_wide(c, c->xfull, y, gated);
  } else {
    int rc = normal_half(c, true, xin, c->xfull, gated);
    if (rc) return rc;
    return c->av_cb ? normal_half(c, false, c->xfull, y, gated) : normal_wide(c, c->xfull, y, gated);
  }
}

// Upload-time autotune of the streamed SpMV: the best (consumer groups, pipeline stages) pair depends
// on the matrix -- stencils want the deepest pipeline (6 groups, 8 stages: 5.99 TB/s on the n = 16M
// Laplacian vs 5.16 with 4 x 6), matrices with scattered columns want fewer stages so that more of
// the SM's unified memory stays L1 for the x gathers (profiles/r01/spmv_stage_sweep.jsonl).  Every
// configuration computes the same bits, so the choice affects speed only.
int spmv_autotune(ga_ctx* c, CsrDev& M, const void* x, void* y) {
  spmv_read_env();
  if (!c->use_spmv_stream || M.ndesc < 4 * c->sm_count) return 0;   // tiny matrices: launch-bound either way
  static const int cand[][2] = {{4, 6}, {4, 5}, {6, 8}, {4, 8}, {6, 6}};
  cudaEvent_t e0, e1;
  GA_CUDA(c, cudaEventCreate(&e0));
  GA_CUDA(c, cudaEventCreate(&e1));
  float best = 1e30f;
  int bg = M.tune_groups, bs = M.tune_stages;
  for (const auto& cd : cand) {
    M.tune_groups = cd[0]; M.tune_stages = cd[1];
    int rc = spmv_launch(c, M, x, y, false);           // warm-up (also sets the kernel attributes)
    if (rc) return rc;
    GA_CUDA(c, cudaEventRecord(e0, c->stream));
    for (int r = 0; r < 3 && rc == 0; ++r) rc = spmv_launch(c, M, x, y, false);
    if (rc) return rc;
    GA_CUDA(c, cudaEventRecord(e1, c->stream));
    GA_CUDA(c, cudaEventSynchronize(e1));
    float ms = 0.f;
    GA_CUDA(c, cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) { best = ms; bg = cd[0]; bs = cd[1]; }
  }
  M.tune_groups = bg; M.tune_stages = bs;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  c->launches = 0;
  return 0;
}

// First half of the normal operator on device vectors: y = A x (At_side = :left, y has m rows) or
// y = A^H x (:right, y has n rows): what eigenvecs_to_singvecs! applies to each eigenvector
// (src/idonow_ops.jl:311,322).
int launch_normal_half(ga_ctx* c, const void* x, void* y) {
  if (c->op_kind != 2) return set_err(c, GA_ERR_NOOP, "not a normal-operator context (ga_set_normal_csr)");
  if (c->normal_dist) {  // y_loc = T_loc [x_own | halo]: the halo of x travels first
    if (c->devcomm) {    // fused path: push into the peers' halo buffers, the product waits on their flags
      int rc = launch_push_halo(c, x);
      if (rc) return rc;
      return spmv_launch(c, c->A, x, y, false, true);
    }
    int rc = comm_halo_exchange(c, x);   // grouped ncclSend/ncclRecv
    if (rc) return rc;
  }
  return normal_half(c, !c->normal_left, x, y, false);
}

// Upload a host CSR (int64 rowptr, int32 col, T val) and build the row blocks.
// like GA_CUDA, but drains the stream first: asynchronous copies out of the CALLER's arrays may be in flight
#define GA_CUDA_DRAIN(ctx, call)                                                                     \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess) {                                                                         \
      cudaStreamSynchronize((ctx)->stream);                                                          \
      return ga::set_err((ctx), GA_ERR_CUDA, std::string(__FILE__) + ":" + std::to_string(__LINE__) + " " + #call + ": " + cudaGetErrorString(e_)); \
    }                                                                                                \
  } while (0)

int csr_upload(ga_ctx* c, CsrDev& M, i64 nrows, i64 ncols, const int64_t* rowptr, const int32_t* col,
               const void* val, i64 col_shift) {
  M.free_all();
  const bool timing = getenv("GA_TIMING") != nullptr;   // debug: stderr breakdown of the upload
  auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double tt0 = now();
  const i64 nnz = rowptr[nrows] - rowptr[0];
  if (nnz >= (i64)2147483647LL || nrows >= (i64)2147483647LL)
    return set_err(c, GA_ERR_ARG, "csr: more than 2^31-1 non-zeros or rows per rank is not supported");
  if (nnz < 0) return set_err(c, GA_ERR_ARG, "csr: rowptr is not non-decreasing");
  const int cap = c->esize == 8 ? SpmvCfg<double>::CAP : SpmvCfg<cdbl>::CAP;
  const int maxrows = SpmvCfg<double>::MAXROWS;
  const int64_t base = rowptr[0];
  M.nrows = nrows; M.ncols = ncols; M.nnz = nnz;
  // The two big arrays go first: their copies (from pinned host memory: true DMA) run while the host validates the
  // indices and builds the row blocks below.  +8 elements: the streamed kernel copies 16-byte aligned supersets of each segment.
  GA_CUDA(c, dev_alloc((void**)&M.col, sizeof(int) * (size_t)(std::max<i64>(nnz, 1) + 8)));
  GA_CUDA(c, dev_alloc(&M.val, c->esize * (size_t)(std::max<i64>(nnz, 1) + 8)));
  GA_CUDA(c, cudaMemsetAsync((char*)M.col + sizeof(int) * (size_t)std::max<i64>(nnz, 1), 0, sizeof(int) * 8, c->stream));
  GA_CUDA(c, cudaMemsetAsync((char*)M.val + c->esize * (size_t)std::max<i64>(nnz, 1), 0, c->esize * 8, c->stream));
  if (nnz > 0) {
    GA_CUDA(c, cudaMemcpyAsync(M.val, (const char*)val + (size_t)base * c->esize, c->esize * (size_t)nnz, cudaMemcpyHostToDevice, c->stream));
    if (col_shift == 0) GA_CUDA(c, cudaMemcpyAsync(M.col, col + base, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice, c->stream));
  }
  const double tt1 = now();
  auto fail = [&](const char* msg) {   // nothing may still be reading the caller's arrays when we return
    cudaStreamSynchronize(c->stream);
    M.free_all();
    return set_err(c, GA_ERR_ARG, msg);
  };
  // the kernels gather x[col] unchecked: a bad index must be an argument error here, not a device fault there.
  // Host passes over the index arrays (range check of col, monotonicity + 32-bit copy of rowptr, the two row-block
  // lists) run on a few host threads: at C2's size (16M rows, 80M entries) they are 125 ms on one thread -- more than
  // the DMA of the matrix itself, and all of it inside the end-to-end time of a solve.
  std::vector<int> rp((size_t)nrows + 1);
  {
    const unsigned hw = std::thread::hardware_concurrency();
    const int nthr = (nnz + nrows < (1 << 20)) ? 1 : (int)std::max(1u, std::min((unsigned)c->upload_threads, hw ? hw : 1u));
    std::vector<int64_t> dmin((size_t)nthr, 0);
    std::vector<int32_t> cmin((size_t)nthr, 2147483647), cmax((size_t)nthr, -2147483647 - 1);
    const int32_t* cb = col + base;
    auto scan = [&](int t) {
      const i64 r0 = nrows * t / nthr, r1 = nrows * (t + 1) / nthr;
      int64_t d = 0;
      for (i64 i = r0; i < r1; ++i) {
        d = std::min<int64_t>(d, rowptr[i + 1] - rowptr[i]);
        rp[(size_t)i] = (int)(rowptr[i] - base);
      }
      dmin[(size_t)t] = d;
      const i64 k0 = nnz * t / nthr, k1 = nnz * (t + 1) / nthr;
      int32_t lo = 2147483647, hi = -2147483647 - 1;
      for (i64 k = k0; k < k1; ++k) { lo = std::min(lo, cb[k]); hi = std::max(hi, cb[k]); }
      cmin[(size_t)t] = lo; cmax[(size_t)t] = hi;
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthr; ++t) th.emplace_back(scan, t);
    scan(0);
    for (auto& x : th) x.join();
    rp[(size_t)nrows] = (int)(rowptr[nrows] - base);
    for (int t = 0; t < nthr; ++t)
      if (dmin[(size_t)t] < 0) return fail("csr: rowptr is not non-decreasing");
    if (nnz > 0) {
      const int32_t lo = *std::min_element(cmin.begin(), cmin.end()), hi = *std::max_element(cmax.begin(), cmax.end());
      if ((i64)lo - col_shift < 0 || (i64)hi - col_shift >= ncols) return fail("csr: column index outside [0, ncols)");
    }
  }
  const double tt2 = now();
  // row blocks of the first kernel (<= CAP non-zeros, <= MAXROWS rows) and of the streamed kernel (one pipeline stage each;
  // longer rows go to a list): two independent sequential scans, one host thread each
  std::vector<int> blk;
  std::vector<int4> desc;
  std::vector<int> longrow;
  auto build_blk = [&]() {
    blk.reserve((size_t)(nnz / (cap / 2) + nrows / maxrows + 16));
    i64 r = 0;
    while (r < nrows) {
      blk.push_back((int)r);
      i64 e = r + 1;  // a row always fits "alone" (long-row path if needed)
      while (e < nrows && (e - r) < maxrows && (rp[(size_t)e + 1] - rp[(size_t)r]) <= cap) ++e;
      r = e;
    }
    blk.push_back((int)nrows);
  };
  auto build_desc = [&]() {
    const int scap = c->esize <= 8 ? SpmvStreamCfg<double>::CAP : SpmvStreamCfg<cdbl>::CAP;   // same rule as the kernels' sizeof(T)
    const int smax = SpmvStreamCfg<double>::MAXROWS;
    desc.reserve((size_t)(nnz / (scap / 2) + nrows / smax + 16));
    i64 r0 = 0;
    while (r0 < nrows) {
      if (rp[(size_t)r0 + 1] - rp[(size_t)r0] > scap) { longrow.push_back((int)r0); ++r0; continue; }
      i64 e = r0 + 1;
      while (e < nrows && (e - r0) < smax && (rp[(size_t)e + 1] - rp[(size_t)r0]) <= scap) ++e;
      desc.push_back(make_int4((int)r0, (int)e, rp[(size_t)r0], rp[(size_t)e]));
      r0 = e;
    }
  };
  if (nrows > (1 << 18) && c->upload_threads > 1) {
    std::thread t1(build_blk);
    build_desc();
    t1.join();
  } else {
    build_blk();
    build_desc();
  }
  M.nblk = (int)blk.size() - 1;
  const double tt3 = now();
  M.ndesc = (int)desc.size(); M.nlong = (int)longrow.size();
  GA_CUDA_DRAIN(c, dev_alloc((void**)&M.rowptr, sizeof(int) * ((size_t)nrows + 1 + 8)));
  GA_CUDA_DRAIN(c, cudaMemsetAsync((char*)M.rowptr + sizeof(int) * ((size_t)nrows + 1), 0, sizeof(int) * 8, c->stream));
  GA_CUDA_DRAIN(c, cudaMalloc(&M.rowblk, sizeof(int) * blk.size()));
  GA_CUDA_DRAIN(c, cudaMalloc(&M.desc, sizeof(int4) * std::max<size_t>(desc.size(), 1)));
  GA_CUDA_DRAIN(c, cudaMalloc(&M.longrow, sizeof(int) * std::max<size_t>(longrow.size(), 1)));
  if (!desc.empty())
    GA_CUDA_DRAIN(c, cudaMemcpyAsync(M.desc, desc.data(), sizeof(int4) * desc.size(), cudaMemcpyHostToDevice, c->stream));
  if (!longrow.empty())
    GA_CUDA_DRAIN(c, cudaMemcpyAsync(M.longrow, longrow.data(), sizeof(int) * longrow.size(), cudaMemcpyHostToDevice, c->stream));
  GA_CUDA_DRAIN(c, cudaMemcpyAsync(M.rowptr, rp.data(), sizeof(int) * rp.size(), cudaMemcpyHostToDevice, c->stream));
  GA_CUDA_DRAIN(c, cudaMemcpyAsync(M.rowblk, blk.data(), sizeof(int) * blk.size(), cudaMemcpyHostToDevice, c->stream));
  if (col_shift != 0 && nnz > 0) {
    std::vector<int> cc((size_t)nnz);
    for (i64 k = 0; k < nnz; ++k) cc[(size_t)k] = (int)(col[base + k] - col_shift);
    GA_CUDA_DRAIN(c, cudaMemcpyAsync(M.col, cc.data(), sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice, c->stream));
    GA_CUDA_DRAIN(c, cudaStreamSynchronize(c->stream));
  }
  const double tt4 = now();
  GA_CUDA_DRAIN(c, cudaStreamSynchronize(c->stream));
  if (timing)
    fprintf(stderr, "csr_upload nnz=%lld: alloc+enqueue %.1f ms, index scans %.1f ms, row blocks %.1f ms, small uploads %.1f ms, drain %.1f ms\n",
            (long long)nnz, (tt1 - tt0) * 1e3, (tt2 - tt1) * 1e3, (tt3 - tt2) * 1e3, (tt4 - tt3) * 1e3, (now() - tt4) * 1e3);
  return 0;
}

}  // namespace ga
