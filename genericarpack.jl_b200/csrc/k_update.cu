// Basis updates V <- V*Q with a small replicated real Q:
//   - device tail of dsapps! (src/dsapps.jl:813-854): V[:,1:kev(+1)] <- V[:,1:kplusp] Q[:,1:kev(+1)],
//     resid <- Q[kplusp,kev] resid + H[kev+1,1] V[:,kev+1], fused with the ||resid|| that dsaup2
//     computes next (src/dsaup2.jl:1397-1414).  The reference issues kev+1 separate dgemv calls
//     over shrinking column ranges plus 3 kev n-vector copies; here every row of V is read once
//     and written once (row-independent, hence safe in place).
//   - device tail of simple_dseupd! (src/dseupd.jl:1413-1435): V <- V Qh where Qh is the product of
//     the nconv Householder reflectors the reference applies one at a time with dorm2r!.
//
// Persistent, warp-specialised kernel (one CTA per SM).  A tile is R rows (one 2 KiB column segment: the TMA bulk-copy
// issue rate makes shorter copies slow, ga_gs_cfg.cuh) of all kplusp columns and one stage of a shared-memory ring;
// a producer warp keeps the ring full with TMA bulk copies (completion on an mbarrier per stage), the consumer warps
// take the stages in order.  A consumer thread owns ONE ROW of the tile and keeps 4*NOB outputs of that row in
// registers: per basis column one shared-memory read of x = V[row,c] and 2*NOB broadcast reads of Q[c, o..] (Q sits in
// shared memory transposed to [c][o]) feed 4*NOB FMAs.  Results go straight to global memory (V is updated in place:
// every row is in the ring before it is written, rows are independent); outputs beyond 4*NOB take further passes over
// the resident tile.
// Why this shape (profiles/r02/README.md): round 1's kernel loaded the tile synchronously (load -> __syncthreads -> compute
// -> store) and re-read x once per block of four outputs; its ncu capture shows 3.45 TB/s, long-scoreboard and barrier
// stalls, the FP64 pipe at 23 % and the LSU pipe at 42 % -- counting shared-memory wavefronts, that mapping needs 3840 per
// 256-row tile against 4180 cycles of HBM time, i.e. it is shared-memory bound right at the roofline.  More outputs per
// thread cut the x re-reads (2560 wavefronts for 12 outputs); the FP64 pipe needs 46 % at the roofline, so the answer
// is register blocking + the asynchronous ring, not tensor cores: DMMA would not add FP64 throughput on this part (same
// rate as the FMA pipe) and would change the rounding order of V*Q away from the reference's dgemv.
#include "ga_decide.cuh"
#include "ga_tma.cuh"

namespace ga {

constexpr size_t kVqSmemBytes = 216 * 1024;

template <class T> struct VqCfg {
  static constexpr int SEG = 2048;                       // bytes of one column segment of a tile
  static constexpr int R = SEG / (int)sizeof(T);         // rows per tile
  static constexpr int RPT = (sizeof(T) == 16) ? 1 : 2;  // adjacent rows per consumer thread (one 16-byte shared-memory read)
  static constexpr int CTHREADS = R / RPT;
  static constexpr int THREADS = CTHREADS + 32;          // + producer warp
  static constexpr int OB = 4;
  static constexpr int MAX_STAGES = 8;
  static constexpr int MAX_NOB = 4;                      // register blocks of 4 outputs per thread (x RPT rows)
};

template <class T, int NOB>
__global__ void __launch_bounds__(VqCfg<T>::THREADS, 1)
k_vq(T* V, i64 ld, i64 ntiles, const typename Elem<T>::real_t* __restrict__ Qg, int ldq, int kplusp, int nout,
     T* resid, int do_resid, int kev, DevCtl* ctl, double2* partial, int decide_kind, int ncv, const DevComm* cm,
     int nstages) {
  typedef typename Elem<T>::real_t TR;
  typedef VqCfg<T> C;
  constexpr int R = C::R, OB = C::OB, CT = C::CTHREADS, SEG = C::SEG, NO = OB * NOB, RPT = C::RPT;
  extern __shared__ __align__(1024) unsigned char smem[];
  const int nout_pad = (nout + NO - 1) / NO * NO;                          // whole passes of NO outputs
  const uint32_t stage_bytes = (uint32_t)kplusp * SEG;
  unsigned char* p = smem + (size_t)nstages * stage_bytes;
  TR* Qs = reinterpret_cast<TR*>(p);                                       // [kplusp][nout_pad]
  p += ((size_t)kplusp * nout_pad * sizeof(TR) + 15) / 16 * 16;
  uint64_t* full = reinterpret_cast<uint64_t*>(p);  p += C::MAX_STAGES * 8;
  uint64_t* empty = reinterpret_cast<uint64_t*>(p);
  __shared__ double2 sw[CT / 32 * 3];
  __shared__ int sh_flag;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < nstages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CT / 32); }
    fence_mbar_init();
  }
  for (int idx = tid; idx < kplusp * nout_pad; idx += C::THREADS) {
    const int c = idx / nout_pad, o = idx % nout_pad;
    Qs[idx] = (o < nout) ? Qg[c + (size_t)o * ldq] : Real<TR>::from(0.0);
  }
  __syncthreads();
  NormAcc nacc;
  nacc.init();
  if (warp == CT / 32) {
    // ---------------- producer warp
    int s = 0;
    uint32_t par = 0;
    for (i64 t = blockIdx.x; t < ntiles; t += gridDim.x) {
      mbar_wait(&empty[s], par ^ 1u);
      if (lane == 0) mbar_arrive_expect_tx(&full[s], stage_bytes);
      __syncwarp();
      unsigned char* st = smem + (size_t)s * stage_bytes;
      const i64 row0 = t * R;
      for (int c = lane; c < kplusp; c += 32) bulk_g2s(st + (size_t)c * SEG, V + (size_t)c * ld + row0, SEG, &full[s]);
      if (++s == nstages) { s = 0; par ^= 1u; }
    }
  } else {
    // ---------------- consumers: thread = RPT adjacent rows x NO outputs in registers (RPT * NO independent FMA chains:
    // the chains are kplusp long, so it is this count, not the warp count, that hides the FP64 latency)
    TR sigma = Real<TR>::from(0.0), beta = Real<TR>::from(0.0);
    if (do_resid) {
      sigma = Qg[(kplusp - 1) + (size_t)(kev - 1) * ldq];   // Q[kplusp,kev]   (src/dsapps.jl:850)
      beta = load_real<TR>(ctl->rnorm1);                    // H[kev+1,1], parked by the launcher
    }
    const int row = tid * RPT;
    int s = 0;
    uint32_t par = 0;
    for (i64 t = blockIdx.x; t < ntiles; t += gridDim.x) {
      const i64 row0 = t * R;
      // the residual entries of these rows: issued before the wait so that their latency hides behind it
      T rres[RPT];
#pragma unroll
      for (int q = 0; q < RPT; ++q) rres[q] = do_resid ? resid[row0 + row + q] : Elem<T>::zero();
      mbar_wait(&full[s], par);
      const T* xs = reinterpret_cast<const T*>(smem + (size_t)s * stage_bytes) + row;
      for (int o0 = 0; o0 < nout_pad; o0 += NO) {
        T y[RPT][NO];
#pragma unroll
        for (int q = 0; q < RPT; ++q)
#pragma unroll
          for (int u = 0; u < NO; ++u) y[q][u] = Elem<T>::zero();
        const TR* qc = Qs + o0;
#pragma unroll 4
        for (int c = 0; c < kplusp; ++c) {
          T x[RPT];
#pragma unroll
          for (int q = 0; q < RPT; ++q) x[q] = xs[(size_t)c * R + q];
#pragma unroll
          for (int u = 0; u < NO; ++u) {
            const TR qv = qc[u];
#pragma unroll
            for (int q = 0; q < RPT; ++q) Elem<T>::add_mul_real(y[q][u], x[q], qv);
          }
          qc += nout_pad;
        }
#pragma unroll
        for (int u = 0; u < NO; ++u) {
          const int o = o0 + u;
          if (o < nout) {
#pragma unroll
            for (int q = 0; q < RPT; ++q) V[(size_t)o * ld + row0 + row + q] = y[q][u];
            if (do_resid && o == kev) {  // beta > 0: r <- sigma r + beta v_{kev+1}  (:850-854)
#pragma unroll
              for (int q = 0; q < RPT; ++q) {
                T rv = Elem<T>::mul_real(rres[q], sigma);
                Elem<T>::add_mul_real(rv, y[q][u], beta);
                resid[row0 + row + q] = rv;
                nacc.add(rv);
              }
            }
          }
        }
      }
      if (do_resid && nout == kev) {  // beta == 0: r <- sigma r
#pragma unroll
        for (int q = 0; q < RPT; ++q) {
          T rv = Elem<T>::mul_real(rres[q], sigma);
          resid[row0 + row + q] = rv;
          nacc.add(rv);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
      if (++s == nstages) { s = 0; par ^= 1u; }
    }
  }
  if (!do_resid) return;
  if (warp < CT / 32) {
    ddbl m = warp_sum_dd(nacc.med), b = warp_sum_dd(nacc.big), s = warp_sum_dd(nacc.sml);
    if (lane == 0) {
      sw[warp * 3 + 0] = make_double2(m.hi, m.lo);
      sw[warp * 3 + 1] = make_double2(b.hi, b.lo);
      sw[warp * 3 + 2] = make_double2(s.hi, s.lo);
    }
  }
  __syncthreads();
  if (tid < 3) {
    double2 acc = make_double2(0.0, 0.0);
    for (int w = 0; w < CT / 32; ++w) acc = slot_add<false>(acc, sw[w * 3 + tid]);
    partial[(size_t)blockIdx.x * kMaxSlots + tid] = acc;
  }
  if (last_block_reduce<false>(ctl, partial, 0, 3, &sh_flag)) {
    if (cm) comm_exchange_slots<false>(ctl, cm, 0, 3);
    gs_decide<T>(ctl, ctl->red, 0, ncv, decide_kind);
  }
}

template <class T, int NOB>
static int launch_vq_nob(ga_ctx* c, int kplusp, int nout, bool do_resid, int kev) {
  typedef typename Elem<T>::real_t TR;
  typedef VqCfg<T> C;
  constexpr int NO = C::OB * NOB;
  const int nout_pad = (nout + NO - 1) / NO * NO;
  const size_t fixed = ((size_t)kplusp * nout_pad * sizeof(TR) + 15) / 16 * 16 + 2 * C::MAX_STAGES * 8 + 64;
  int nst = (int)((kVqSmemBytes - fixed) / ((size_t)kplusp * C::SEG));
  if (nst > C::MAX_STAGES) nst = C::MAX_STAGES;
  if (nst < 1) return set_err(c, GA_ERR_ARG, "vq: shared memory budget exceeded");
  const i64 ntiles = c->n_pad / C::R;
  const size_t smem = fixed + (size_t)nst * kplusp * C::SEG;
  static PerDeviceOnce attr_set;
  if (attr_set.need(c->device)) {
    GA_CUDA(c, cudaFuncSetAttribute(k_vq<T, NOB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kVqSmemBytes));
    attr_set.done(c->device);
  }
  i64 grid = (i64)c->sm_count;
  if (grid > ntiles) grid = ntiles;
  if (grid > kMaxCtas) grid = kMaxCtas;
  if (grid < 1) grid = 1;
  const bool single = c->world == 1 || c->devcomm != nullptr;  // decide inside the kernel
  k_vq<T, NOB><<<(int)grid, C::THREADS, smem, c->stream>>>((T*)c->V, c->n_pad, ntiles, (const TR*)c->qdev, c->ncv_pad,
                                                         kplusp, nout, (T*)c->resid, do_resid ? 1 : 0, kev, c->ctl,
                                                         c->partial, single ? DK_NORM : DK_NONE, c->ncv,
                                                         c->world > 1 ? c->devcomm : nullptr, nst);
  GA_CUDA(c, cudaGetLastError());
  c->launches += 1;
  if (do_resid && !single) return launch_decide(c, 0, -1, DK_NORM);
  return 0;
}

// ---- streaming variant for up to 4 * MAX_NOB outputs (the usual implicit restart: nout = kev + 1) ------------------------
// The whole-tile ring above holds only two 80 KiB tiles of a 40-column basis and its consumer is one warp per scheduler: the
// ncu capture shows that warp stalled on its own fixed latencies (`wait` 1.8 per issue) and on data (`long_scoreboard` 1.2).
// Here a tile travels in CHUNKS of KC columns (one ring stage each) and the outputs of a row live in registers ACROSS the
// chunks, so a chunk is released as soon as it has been multiplied: the ring stays full (12 x 16 KiB in flight) and TWO
// independent consumer groups take alternate tiles, hiding each other's latencies.
template <class T, int NOB>
__global__ void __launch_bounds__(2 * VqCfg<T>::CTHREADS + 32, 1)
k_vq_stream(T* V, i64 ld, i64 ntiles, const typename Elem<T>::real_t* __restrict__ Qg, int ldq, int kplusp, int nout,
            T* resid, int do_resid, int kev, DevCtl* ctl, double2* partial, int decide_kind, int ncv, const DevComm* cm,
            int nstages, int KC) {
  typedef typename Elem<T>::real_t TR;
  typedef VqCfg<T> C;
  constexpr int R = C::R, OB = C::OB, GT = C::CTHREADS, SEG = C::SEG, NO = OB * NOB, RPT = C::RPT, GROUPS = 2;
  constexpr int CT = GROUPS * GT;
  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t chunk_bytes = (uint32_t)KC * SEG;
  const int nch = (kplusp + KC - 1) / KC;
  unsigned char* p = smem + (size_t)nstages * chunk_bytes;
  TR* Qs = reinterpret_cast<TR*>(p);                                       // [kplusp][NO]
  p += ((size_t)kplusp * NO * sizeof(TR) + 15) / 16 * 16;
  uint64_t* full = reinterpret_cast<uint64_t*>(p);  p += 32 * 8;
  uint64_t* empty = reinterpret_cast<uint64_t*>(p);
  __shared__ double2 sw[CT / 32 * 3];
  __shared__ int sh_flag;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < nstages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], GT / 32); }
    fence_mbar_init();
  }
  for (int idx = tid; idx < kplusp * NO; idx += blockDim.x) {
    const int c = idx / NO, o = idx % NO;
    Qs[idx] = (o < nout) ? Qg[c + (size_t)o * ldq] : Real<TR>::from(0.0);
  }
  __syncthreads();
  NormAcc nacc;
  nacc.init();
  if (warp == CT / 32) {
    // ---------------- producer warp: chunk after chunk, tile after tile
    int s = 0;
    uint32_t par = 0;
    for (i64 t = blockIdx.x; t < ntiles; t += gridDim.x) {
      const i64 row0 = t * R;
      for (int ch = 0; ch < nch; ++ch) {
        const int c0 = ch * KC, kc = (kplusp - c0 < KC) ? (kplusp - c0) : KC;
        mbar_wait(&empty[s], par ^ 1u);
        if (lane == 0) mbar_arrive_expect_tx(&full[s], (uint32_t)kc * SEG);
        __syncwarp();
        unsigned char* st = smem + (size_t)s * chunk_bytes;
        for (int c = lane; c < kc; c += 32) bulk_g2s(st + (size_t)c * SEG, V + (size_t)(c0 + c) * ld + row0, SEG, &full[s]);
        if (++s == nstages) { s = 0; par ^= 1u; }
      }
    }
  } else {
    // ---------------- consumer group g: tiles g, g + 2, ... of this CTA's sequence
    TR sigma = Real<TR>::from(0.0), beta = Real<TR>::from(0.0);
    if (do_resid) {
      sigma = Qg[(kplusp - 1) + (size_t)(kev - 1) * ldq];   // Q[kplusp,kev]   (src/dsapps.jl:850)
      beta = load_real<TR>(ctl->rnorm1);                    // H[kev+1,1], parked by the launcher
    }
    const int g = tid / GT, row = (tid - g * GT) * RPT;
    // ring position of this group's first tile: g * nch stages in
    int s = 0;
    uint32_t par = 0;
    auto advance = [&](int n) { s += n; while (s >= nstages) { s -= nstages; par ^= 1u; } };
    advance(g * nch);
    for (i64 t = (i64)blockIdx.x + (i64)g * gridDim.x; t < ntiles; t += (i64)GROUPS * gridDim.x) {
      const i64 row0 = t * R;
      T rres[RPT];
#pragma unroll
      for (int q = 0; q < RPT; ++q) rres[q] = do_resid ? resid[row0 + row + q] : Elem<T>::zero();
      T y[RPT][NO];
#pragma unroll
      for (int q = 0; q < RPT; ++q)
#pragma unroll
        for (int u = 0; u < NO; ++u) y[q][u] = Elem<T>::zero();
      const TR* qc = Qs;
      for (int c0 = 0; c0 < kplusp; c0 += KC) {
        const int kc = (kplusp - c0 < KC) ? (kplusp - c0) : KC;
        mbar_wait(&full[s], par);
        const T* xs = reinterpret_cast<const T*>(smem + (size_t)s * chunk_bytes) + row;
#pragma unroll 4
        for (int c = 0; c < kc; ++c) {
          T x[RPT];
#pragma unroll
          for (int q = 0; q < RPT; ++q) x[q] = xs[(size_t)c * R + q];
#pragma unroll
          for (int u = 0; u < NO; ++u) {
            const TR qv = qc[u];
#pragma unroll
            for (int q = 0; q < RPT; ++q) Elem<T>::add_mul_real(y[q][u], x[q], qv);
          }
          qc += NO;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        advance(1);
      }
#pragma unroll
      for (int u = 0; u < NO; ++u) {
        if (u < nout) {
#pragma unroll
          for (int q = 0; q < RPT; ++q) V[(size_t)u * ld + row0 + row + q] = y[q][u];
          if (do_resid && u == kev) {  // beta > 0: r <- sigma r + beta v_{kev+1}  (:850-854)
#pragma unroll
            for (int q = 0; q < RPT; ++q) {
              T rv = Elem<T>::mul_real(rres[q], sigma);
              Elem<T>::add_mul_real(rv, y[q][u], beta);
              resid[row0 + row + q] = rv;
              nacc.add(rv);
            }
          }
        }
      }
      if (do_resid && nout == kev) {  // beta == 0: r <- sigma r
#pragma unroll
        for (int q = 0; q < RPT; ++q) {
          T rv = Elem<T>::mul_real(rres[q], sigma);
          resid[row0 + row + q] = rv;
          nacc.add(rv);
        }
      }
      advance((GROUPS - 1) * nch);        // skip the other group's tile
    }
  }
  if (!do_resid) return;
  if (warp < CT / 32) {
    ddbl m = warp_sum_dd(nacc.med), b = warp_sum_dd(nacc.big), sm2 = warp_sum_dd(nacc.sml);
    if (lane == 0) {
      sw[warp * 3 + 0] = make_double2(m.hi, m.lo);
      sw[warp * 3 + 1] = make_double2(b.hi, b.lo);
      sw[warp * 3 + 2] = make_double2(sm2.hi, sm2.lo);
    }
  }
  __syncthreads();
  if (tid < 3) {
    double2 acc = make_double2(0.0, 0.0);
    for (int w = 0; w < CT / 32; ++w) acc = slot_add<false>(acc, sw[w * 3 + tid]);
    partial[(size_t)blockIdx.x * kMaxSlots + tid] = acc;
  }
  if (last_block_reduce<false>(ctl, partial, 0, 3, &sh_flag)) {
    if (cm) comm_exchange_slots<false>(ctl, cm, 0, 3);
    gs_decide<T>(ctl, ctl->red, 0, ncv, decide_kind);
  }
}

// returns 1 if the streaming kernel does not apply (the caller falls back to the whole-tile kernel)
template <class T, int NOB>
static int launch_vq_stream_nob(ga_ctx* c, int kplusp, int nout, bool do_resid, int kev) {
  typedef typename Elem<T>::real_t TR;
  typedef VqCfg<T> C;
  constexpr int NO = C::OB * NOB;
  const int KC = 8;
  const int nch = (kplusp + KC - 1) / KC;
  const size_t fixed = ((size_t)kplusp * NO * sizeof(TR) + 15) / 16 * 16 + 2 * 32 * 8 + 64;
  int nst = (int)((kVqSmemBytes - fixed) / ((size_t)KC * C::SEG));
  if (nst > 32) nst = 32;
  if (nst < 2 * nch + 1) return 1;                 // both groups' tiles + something in flight
  const i64 ntiles = c->n_pad / C::R;
  const size_t smem = fixed + (size_t)nst * KC * C::SEG;
  static PerDeviceOnce attr_set;
  if (attr_set.need(c->device)) {
    GA_CUDA(c, cudaFuncSetAttribute(k_vq_stream<T, NOB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kVqSmemBytes));
    attr_set.done(c->device);
  }
  i64 grid = (i64)c->sm_count;
  if (grid > ntiles) grid = ntiles;
  if (grid > kMaxCtas) grid = kMaxCtas;
  if (grid < 1) grid = 1;
  const bool single = c->world == 1 || c->devcomm != nullptr;  // decide inside the kernel
  k_vq_stream<T, NOB><<<(int)grid, 2 * C::CTHREADS + 32, smem, c->stream>>>((T*)c->V, c->n_pad, ntiles, (const TR*)c->qdev, c->ncv_pad,
                                                                          kplusp, nout, (T*)c->resid, do_resid ? 1 : 0, kev, c->ctl,
                                                                          c->partial, single ? DK_NORM : DK_NONE, c->ncv,
                                                                          c->world > 1 ? c->devcomm : nullptr, nst, KC);
  GA_CUDA(c, cudaGetLastError());
  c->launches += 1;
  if (do_resid && !single) return launch_decide(c, 0, -1, DK_NORM);
  return 0;
}

template <class T>
static int launch_vq_t(ga_ctx* c, int kplusp, int nout, bool do_resid, int kev) {
  // the smallest register blocking that covers the outputs in the fewest passes
  constexpr int MAXB = VqCfg<T>::MAX_NOB;
  const int blocks = (nout + 3) / 4;
  if constexpr (sizeof(T) == 8) {
   if (c->vq_stream && blocks <= MAXB) {
    int rc = 1;
    switch (blocks) {
      case 1: rc = launch_vq_stream_nob<T, 1>(c, kplusp, nout, do_resid, kev); break;
      case 2: rc = launch_vq_stream_nob<T, 2>(c, kplusp, nout, do_resid, kev); break;
      case 3: rc = launch_vq_stream_nob<T, 3>(c, kplusp, nout, do_resid, kev); break;
      default: rc = launch_vq_stream_nob<T, 4>(c, kplusp, nout, do_resid, kev); break;
    }
    if (rc != 1) return rc;
   }
  }
  const int passes = (blocks + MAXB - 1) / MAXB;
  const int nob = (blocks + passes - 1) / passes;
  if (nob <= 1) return launch_vq_nob<T, 1>(c, kplusp, nout, do_resid, kev);
  if (nob == 2 || MAXB == 2) return launch_vq_nob<T, 2>(c, kplusp, nout, do_resid, kev);
  if constexpr (MAXB >= 4) {
    if (nob == 3) return launch_vq_nob<T, 3>(c, kplusp, nout, do_resid, kev);
    return launch_vq_nob<T, 4>(c, kplusp, nout, do_resid, kev);
  }
  return launch_vq_nob<T, 2>(c, kplusp, nout, do_resid, kev);
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
