// Fused tall-skinny Gram-Schmidt sweeps of the Lanczos step (dsaitr! steps 4-5).
//
// Reference (src/dsaitr.jl):  h = V_j' w (:1227, dgemv 'T');  r = w - V_j h (:1240, dgemv 'N');
// DGKS: s = V_j' r (:1352); r -= V_j s (:1363); each followed by a norm2 pass (:1208,1304,1410).
// That is 2 sweeps of V[:,1:j] per Gram-Schmidt round plus ~6 n-vector passes.
//
// Here one persistent kernel streams V[:,1:j] through shared memory ONCE per sweep with TMA bulk
// copies (one 1 KiB column segment per copy, multi-stage mbarrier pipeline, dedicated producer
// warp) and does, per row tile resident in shared memory:
//     MODE 0:  s += V' w,  ||w||^2                                   (first dgemv 'T' + norm)
//     MODE 1:  r = w - V h  (written once),  s += V' r,  ||r||^2    (dgemv 'N' + the NEXT dgemv 'T' + norm)
//     MODE 2:  r = w - V h,  ||r||^2                                (last correction)
// MODE 1 is what removes a whole sweep: the DGKS coefficients s = V' r are row-local once r is
// known for the tile, so they are accumulated while the tile is still in shared memory.  The
// arithmetic (classical Gram-Schmidt with DGKS correction) is unchanged.
//
// Thread layout: R rows per tile (R * sizeof(T) = 1 KiB), TG thread groups; thread (row, g) owns
// row `row` of the tile and columns c = g, g+TG, ...  -> NC = 64/TG register accumulators.
// Groups exchange their partial row sums through shared memory behind a named barrier that
// only spans the TG warps sharing 32 rows.  Reductions are fixed-order (bit-reproducible).
#include "ga_gs_cfg.cuh"

namespace ga {

// NCT = register accumulators per consumer thread; the launcher picks the smallest instantiation covering j
// columns (the binary64 kernel's Dot2 accumulators are two doubles each, see ga_types.cuh).
template <class T, int MODE, int SEG, int NCT>
__global__ void __launch_bounds__(GsCfg<T, SEG>::THREADS, 1)
k_gs(const T* __restrict__ V, i64 ldv, int j, const T* w, T* r, i64 ntiles, DevCtl* ctl, double2* partial,
     int nstages, int gate_phase, int decide_kind, int ncv, const DevComm* cm, unsigned long long* trace, int reverse) {
  typedef GsCfg<T, SEG> C;
#define GA_TRACE(slot, cond)                                                        \
  if (trace != nullptr && (cond)) {                                                 \
    unsigned long long t_;                                                          \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                          \
    trace[slot] = t_;                                                               \
  }
  GA_TRACE(0, blockIdx.x == 0 && threadIdx.x == 0)
  constexpr int R = C::R, TG = C::TG, NC = C::NC, SLICES = C::SLICES, CWARPS = C::CWARPS, RPT = C::RPT, ROWL = C::ROWL;
  constexpr uint32_t SEGB = C::SEG_BYTES;
  constexpr bool CPLX = Elem<T>::is_complex;

  if (ctl->status != ST_OK) return;
  if (gate_phase >= 0 && ctl->phase != gate_phase) return;

  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t stage_bytes = (uint32_t)(j + 1) * SEGB;
  unsigned char* p = smem + (size_t)nstages * stage_bytes;
  T* hs = reinterpret_cast<T*>(p);                  p += kMaxNcv * 16;
  T* pr = reinterpret_cast<T*>(p);                  p += 2 * TG * SEGB;
  double2* redw = reinterpret_cast<double2*>(p);    p += CWARPS * (NC + 3) * 16;
  uint64_t* full = reinterpret_cast<uint64_t*>(p);  p += C::MAX_STAGES * 8;
  uint64_t* empty = reinterpret_cast<uint64_t*>(p); p += C::MAX_STAGES * 8;
  int* sh_flag = reinterpret_cast<int*>(p);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < nstages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CWARPS); }
    fence_mbar_init();
  }
  if (MODE != 0) {
    const T* hb = hbuf_ptr<T>(ctl);
    for (int c = tid; c < j; c += blockDim.x) hs[c] = hb[c];
  }
  __syncthreads();
  GA_TRACE(1, blockIdx.x == 0 && threadIdx.x == 0)

  typename Elem<T>::acc_t acc[NCT];
#pragma unroll
  for (int i = 0; i < NCT; ++i) acc[i] = typename Elem<T>::acc_t(Elem<T>::zero());
  NormAcc nacc;
  nacc.init();

  if (warp == CWARPS) {
    // ---------------- producer warp: TMA bulk copies of the (j columns + w) x R tile
    int it = 0;
    for (i64 t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      const int s = it % nstages;
      const uint32_t ph = (uint32_t)(it / nstages) & 1u;
      mbar_wait(&empty[s], ph ^ 1u);
      if (lane == 0) mbar_arrive_expect_tx(&full[s], stage_bytes);
      __syncwarp();
      unsigned char* st = smem + (size_t)s * stage_bytes;
      const i64 row0 = (reverse ? (ntiles - 1 - t) : t) * R;
      for (int c = lane; c <= j; c += 32) {
        const T* src = (c < j) ? (V + (size_t)c * ldv + row0) : (w + row0);
        bulk_g2s(st + (size_t)c * SEGB, src, SEGB, &full[s]);
      }
    }
  } else {
    // ---------------- consumers: thread (rl, g) owns rows 2rl, 2rl+1 and columns g, g+TG, ...
    const int slice = warp % SLICES, g = warp / SLICES, rl = slice * 32 + lane;
    // rows of the tile owned by this thread: adjacent for 8-byte elements (one LDS.128 fetches
    // both), ROWL apart for 16-byte elements (adjacent rows would be a 32-byte lane stride =
    // 2-way bank conflicts on every shared-memory access; measured 50 % conflict wavefronts)
#define ROWIDX(q) ((sizeof(T) <= 8) ? (rl * RPT + (q)) : (rl + (q) * ROWL))
    int it = 0;
    for (i64 t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      const int s = it % nstages;
      const uint32_t ph = (uint32_t)(it / nstages) & 1u;
      mbar_wait(&full[s], ph);
      GA_TRACE(2, blockIdx.x == 0 && tid == 0 && it == 0)
      const T* tile = reinterpret_cast<const T*>(smem + (size_t)s * stage_bytes);
      T rr[RPT];
#pragma unroll
      for (int q = 0; q < RPT; ++q) rr[q] = tile[(size_t)j * R + ROWIDX(q)];  // w
      if (MODE != 0) {
        T p0[RPT];
#pragma unroll
        for (int q = 0; q < RPT; ++q) p0[q] = Elem<T>::zero();
#pragma unroll
        for (int i = 0; i < NCT; ++i) {
          const int c = g + TG * i;
          if (c >= j) break;
          const T hc = hs[c];
#pragma unroll
          for (int q = 0; q < RPT; ++q) Elem<T>::sub_mul(p0[q], tile[(size_t)c * R + ROWIDX(q)], hc);
        }
        T* prb = pr + (size_t)(it & 1) * TG * R;
#pragma unroll
        for (int q = 0; q < RPT; ++q) prb[g * R + ROWIDX(q)] = p0[q];
        named_bar_sync(1 + slice, TG * 32);
#pragma unroll
        for (int g2 = 0; g2 < TG; ++g2) {
#pragma unroll
          for (int q = 0; q < RPT; ++q) Elem<T>::add(rr[q], prb[g2 * R + ROWIDX(q)]);
        }
        if (g == 0) {
#pragma unroll
          for (int q = 0; q < RPT; ++q) r[(reverse ? (ntiles - 1 - t) : t) * R + ROWIDX(q)] = rr[q];
        }
      }
      if (MODE != 2) {
#pragma unroll
        for (int i = 0; i < NCT; ++i) {
          const int c = g + TG * i;
          if (c >= j) break;
#pragma unroll
          for (int q = 0; q < RPT; ++q) Elem<T>::acc_conj_mul(acc[i], tile[(size_t)c * R + ROWIDX(q)], rr[q]);
        }
      }
      if (g == 0) {
#pragma unroll
        for (int q = 0; q < RPT; ++q) nacc.add(rr[q]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
#undef ROWIDX
    GA_TRACE(3, blockIdx.x == 0 && tid == 0)
    // warp-level reduction of this warp's accumulators (only the columns this group owns)
    if (MODE != 2) {
#pragma unroll
      for (int i = 0; i < NCT; ++i) {
        if (g + TG * i < j) {
          double2 sl = warp_sum_slot_fast(acc[i]);
          if (lane == 0) redw[warp * (NC + 3) + i] = sl;
        }
      }
    }
    if (g == 0) {
      ddbl m = warp_sum_dd(nacc.med), b = warp_sum_dd(nacc.big), sm = warp_sum_dd(nacc.sml);
      if (lane == 0) {
        redw[warp * (NC + 3) + NC + 0] = make_double2(m.hi, m.lo);
        redw[warp * (NC + 3) + NC + 1] = make_double2(b.hi, b.lo);
        redw[warp * (NC + 3) + NC + 2] = make_double2(sm.hi, sm.lo);
      }
    }
  }
  GA_TRACE(4, blockIdx.x == 0 && tid == 0)
  __syncthreads();
  GA_TRACE(5, blockIdx.x == 0 && tid == 0)
  // CTA partials -> global
  const int nslots = j + 3;
  for (int s = tid; s < nslots; s += blockDim.x) {
    double2 a = make_double2(0.0, 0.0);
    if (s < j) {
      if (MODE != 2) {
        const int g = s % TG, i = s / TG;
        for (int sl = 0; sl < SLICES; ++sl) {
          double2 v = redw[(g * SLICES + sl) * (NC + 3) + i];
          a = CPLX ? slot_add<true>(a, v) : slot_add<false>(a, v);
        }
      }
    } else {
      const int k = s - j;
      for (int sl = 0; sl < SLICES; ++sl) a = slot_add<false>(a, redw[sl * (NC + 3) + NC + k]);
    }
    partial[(size_t)blockIdx.x * kMaxSlots + s] = a;
  }
  if (last_block_reduce<CPLX>(ctl, partial, j, nslots, sh_flag)) {
    GA_TRACE(6, tid == 0)
    if (tid == 0) { ctl->prof_runs += 1u; ctl->prof_units += (unsigned long long)(j + 1 + (MODE != 0 ? 1 : 0)); }
    if (cm) comm_exchange_slots<CPLX>(ctl, cm, j, nslots);
    gs_decide<T>(ctl, ctl->red, j, ncv, decide_kind);
    GA_TRACE(7, tid == 0)
  }
#undef GA_TRACE
}

// One-CTA kernel that sums the gathered per-rank slots and runs the decision logic (multi GPU).
template <class T>
__global__ void k_decide(DevCtl* ctl, const double2* gathered, int world, int j, int ncv, int gate_phase,
                         int decide_kind) {
  if (ctl->status != ST_OK) return;
  if (gate_phase >= 0 && ctl->phase != gate_phase) return;
  constexpr bool CPLX = Elem<T>::is_complex;
  __shared__ double2 red[kMaxSlots];
  const int nslots = j + 3;
  for (int s = threadIdx.x; s < nslots; s += blockDim.x) {
    double2 a = make_double2(0.0, 0.0);
    for (int rk = 0; rk < world; ++rk) {
      double2 v = gathered[(size_t)rk * kMaxSlots + s];
      a = (CPLX && s < j) ? slot_add<true>(a, v) : slot_add<false>(a, v);
    }
    red[s] = a;
  }
  __syncthreads();
  gs_decide<T>(ctl, red, j, ncv, decide_kind);
}

template <class T, int MODE, int SEG, int NCT>
static int launch_gs_nct(ga_ctx* c, int j, int gate_phase, int decide_kind, int nstages) {
  typedef GsCfg<T, SEG> C;
  const i64 ntiles = c->vs.ld / C::R;
  const size_t stage_bytes = (size_t)(j + 1) * C::SEG_BYTES;
  const size_t smem = (size_t)nstages * stage_bytes + C::FIXED_SMEM;
  static PerDeviceOnce attr_set;
  if (attr_set.need(c->device)) {
    GA_CUDA(c, cudaFuncSetAttribute(k_gs<T, MODE, SEG, NCT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGsSmemBytes));
    attr_set.done(c->device);
  }
  int grid = (int)(ntiles < (i64)c->sm_count ? ntiles : (i64)c->sm_count);
  if (grid > kMaxCtas) grid = kMaxCtas;
  const T* V = (const T*)c->vs.V;
  T* r = (T*)c->vs.w;
  // bmat = :G: the dots of a MODE 0 sweep are taken with B r / A v_j (c->gs_dotvec) instead of r itself
  const T* wv = (MODE == 0 && c->gs_dotvec != nullptr) ? (const T*)c->gs_dotvec : r;
  const bool fused = c->world == 1 || c->devcomm != nullptr;  // decide inside the sweep kernel
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (c->prof) {
    if (c->prof_used + 2 > c->prof_ev.size()) {
      for (int i = 0; i < 2; ++i) { cudaEvent_t e; GA_CUDA(c, cudaEventCreate(&e)); c->prof_ev.push_back(e); }
    }
    e0 = c->prof_ev[c->prof_used]; e1 = c->prof_ev[c->prof_used + 1];
    c->prof_used += 2;
    GA_CUDA(c, cudaEventRecord(e0, c->stream));
  }
  k_gs<T, MODE, SEG, NCT><<<grid, C::THREADS, smem, c->stream>>>(V, c->vs.ld, j, wv, r, ntiles, c->ctl, c->partial, nstages,
                                                           gate_phase, fused ? decide_kind : DK_NONE, c->ncv,
                                                           c->world > 1 ? c->devcomm : nullptr, c->trace,
                                                           (int)(c->gs_flip++ & 1));
  GA_CUDA(c, cudaGetLastError());
  c->launches += 1;
  if (c->prof) GA_CUDA(c, cudaEventRecord(e1, c->stream));
  return 0;
}

template <class T, int MODE, int SEG>
static int launch_gs_seg(ga_ctx* c, int j, int gate_phase, int decide_kind, int nstages) {
  typedef GsCfg<T, SEG> C;
  if (sizeof(T) == 8 && MODE != 2) {   // binary64 sweeps with dots: Dot2 accumulators, register-bound -> three widths
    if (j <= C::TG * (C::NC / 2))
      return launch_gs_nct<T, MODE, SEG, (sizeof(T) == 8 && MODE != 2 ? C::NC / 2 : C::NC)>(c, j, gate_phase, decide_kind, nstages);
    if (j <= C::TG * (3 * C::NC / 4))
      return launch_gs_nct<T, MODE, SEG, (sizeof(T) == 8 && MODE != 2 ? 3 * C::NC / 4 : C::NC)>(c, j, gate_phase, decide_kind, nstages);
  }
  return launch_gs_nct<T, MODE, SEG, C::NC>(c, j, gate_phase, decide_kind, nstages);
}

template <class T, int MODE>
static int launch_gs_t(ga_ctx* c, int j, int gate_phase, int decide_kind) {
  // 2 KiB segments whenever at least two pipeline stages fit, else 1 KiB segments
  const size_t budget2 = kGsSmemBytes - GsCfg<T, 2048>::FIXED_SMEM;
  int nst = (int)(budget2 / ((size_t)(j + 1) * 2048));
  if (nst >= 2) {
    if (nst > GsCfg<T, 2048>::MAX_STAGES) nst = GsCfg<T, 2048>::MAX_STAGES;
    return launch_gs_seg<T, MODE, 2048>(c, j, gate_phase, decide_kind, nst);
  }
  const size_t budget1 = kGsSmemBytes - GsCfg<T, 1024>::FIXED_SMEM;
  nst = (int)(budget1 / ((size_t)(j + 1) * 1024));
  if (nst > GsCfg<T, 1024>::MAX_STAGES) nst = GsCfg<T, 1024>::MAX_STAGES;
  if (nst < 2) return set_err(c, GA_ERR_ARG, "gs: too many columns for the shared-memory pipeline");
  return launch_gs_seg<T, MODE, 1024>(c, j, gate_phase, decide_kind, nst);
}

template <class T>
static int launch_gs_mode(ga_ctx* c, int mode, int j, int gate_phase, int decide_kind) {
  switch (mode) {
    case 0: return launch_gs_t<T, 0>(c, j, gate_phase, decide_kind);
    case 1: return launch_gs_t<T, 1>(c, j, gate_phase, decide_kind);
    default: return launch_gs_t<T, 2>(c, j, gate_phase, decide_kind);
  }
}

// mode: 0 dot, 1 update+dot, 2 update.  j = number of basis columns.  The vector being
// orthogonalised is c->resid (in place).  gate_phase >= 0: run only if ctl->phase == gate_phase.
int launch_gs(ga_ctx* c, int mode, int j, int gate_phase, int decide_kind) {
  if (j < 0 || j > kMaxNcv) return set_err(c, GA_ERR_ARG, "gs: j out of range");
  int rc;
  switch (c->dtype) {
    case GA_F64: rc = launch_gs_mode<double>(c, mode, j, gate_phase, decide_kind); break;
    case GA_C64: rc = launch_gs_mode<cdbl>(c, mode, j, gate_phase, decide_kind); break;
    case GA_F32: rc = launch_gs_mode<float>(c, mode, j, gate_phase, decide_kind); break;
    default: rc = launch_gs_mode<ddbl>(c, mode, j, gate_phase, decide_kind); break;
  }
  if (rc) return rc;
  if (c->world > 1 && !c->devcomm) return launch_decide(c, j, gate_phase, decide_kind);  // NCCL fallback
  return 0;
}

int comm_allgather_slots(ga_ctx* c, int nslots);  // ga_comm.cu

int launch_decide(ga_ctx* c, int j, int gate_phase, int decide_kind) {
  int rc = comm_allgather_slots(c, j + 3);
  if (rc) return rc;
  switch (c->dtype) {
    case GA_F64: k_decide<double><<<1, 128, 0, c->stream>>>(c->ctl, c->gather, c->world, j, c->ncv, gate_phase, decide_kind); break;
    case GA_C64: k_decide<cdbl><<<1, 128, 0, c->stream>>>(c->ctl, c->gather, c->world, j, c->ncv, gate_phase, decide_kind); break;
    case GA_F32: k_decide<float><<<1, 128, 0, c->stream>>>(c->ctl, c->gather, c->world, j, c->ncv, gate_phase, decide_kind); break;
    default: k_decide<ddbl><<<1, 128, 0, c->stream>>>(c->ctl, c->gather, c->world, j, c->ncv, gate_phase, decide_kind); break;
  }
  GA_CUDA(c, cudaGetLastError());
  c->launches += 1;
  return 0;
}

}  // namespace ga
