// All Gram-Schmidt sweeps of ONE Lanczos step in a single persistent, cooperative launch.
//
// k_gs.cu runs one sweep per launch and lets the last CTA reduce and decide; the measured fixed
// cost per sweep (profiles/RESULTS.md, in-kernel timeline) is ~23 us: last-CTA reduction + decision
// logic + launch gap + pipeline refill.  At n = 16M that is 4 % of a sweep, at 2M rows per GPU
// (8-GPU strong scaling) 25 %, and for L2-resident problems it dominates.  Here the 148 CTAs stay
// resident for the whole step and the phases
//     0  s = V'w, ||w||                      (dsaitr.jl:1208,1227)
//     1  r = w - V h, s = V'r, ||r||         (:1240-1324)
//     2  r -= V s, ||r||             only if the 0.717 test asked for DGKS          (:1352-1434)
//     3  s = V'r                     only if a second refinement is needed         (:1352 of the second round)
//     4  r -= V s, ||r||             "                                             (:1424-1446)
// (round 1 formed the second round's dots inside sweep 2 to save a sweep in the rare case; with error-free dot
// products that speculation costs 10 FP64 instructions per element in EVERY step, so the dots now get their own sweep
// when -- and only when -- the 0.717 test asks for them; same numbers either way)
// are separated by grid-wide barriers: CTA partials -> barrier -> CTA 0 sums them in fixed order,
// (multi-GPU: exchanges the slots with the peers over NVLink peer memory), runs the step's
// decision logic -> barrier -> everybody reads the verdict and the new coefficients and goes on.
// Same arithmetic, same fixed summation order as k_gs.cu, bit-identical results.
//
// Consumer inner loops (round 2, profiles/r02/README.md): the ncu source page of the round-1 kernel showed one
// branch per basis column (`if (c >= j) break` in the unrolled loops) -- every column's LDS -> FP64 chain sat in
// its own basic block, nothing overlapped, and `wait` / `short_scoreboard` stalls dominated -- plus an integer
// division (I2F/MUFU.RCP sequence) per tile for the stage index.  Now the kernel is instantiated per NCT =
// ceil(j / TG) (columns per thread), the column loops are branch-free straight-line code (columns beyond the
// basis are clamped onto the vector's own column: their coefficient is zero and their dot-product accumulator
// is never read), and the stage ring is walked with an incrementing index and parity.
#pragma once
#include <cooperative_groups.h>

#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "ga_gs_cfg.cuh"

namespace ga {

__device__ __forceinline__ unsigned int ld_gpu_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_gpu_u32(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// sense-reversing grid barrier on two words of the control block; all CTAs are co-resident
// (cooperative launch).  Bounded spin: a broken barrier flags ST_COMM instead of hanging the GPU.
__device__ __forceinline__ void grid_barrier(DevCtl* ctl) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int gen = ld_gpu_u32(&ctl->gbar_gen);
    if (atomicAdd(&ctl->gbar_count, 1u) == gridDim.x - 1) {
      ctl->gbar_count = 0u;
      __threadfence();
      st_gpu_u32(&ctl->gbar_gen, gen + 1u);
    } else {
      long long it = 0;
      while (ld_gpu_u32(&ctl->gbar_gen) == gen) {
        if (++it > (1LL << 27)) { ctl->status = ST_COMM; break; }
      }
    }
    __threadfence();
  }
  __syncthreads();
}

// dot-product accumulator of the sweep: Elem<T>::acc_t, or a plain binary64 sum for GA_DOT_PLAIN (Float64 only)
template <class T, bool PLAIN> struct StepAcc { typedef typename Elem<T>::acc_t type; };
template <> struct StepAcc<double, true> { typedef double type; };

// reduced (hi, lo) / (re, im) slot of one register accumulator
__device__ __forceinline__ double2 slot_from_acc(double v) { return make_double2(v, 0.0); }
__device__ __forceinline__ double2 slot_from_acc(dot2 v) { double h, l; two_sum(v.hi, v.lo, h, l); return make_double2(h, l); }
__device__ __forceinline__ double2 slot_from_acc(cdbl v) { return make_double2(v.re, v.im); }
__device__ __forceinline__ double2 slot_from_acc(ddbl v) { return make_double2(v.hi, v.lo); }
// a + b for two reduced slots, in the arithmetic of the accumulator kind (first argument: tag only)
__device__ __forceinline__ double2 slot_add_recv(double, double2 a, double2 b) { return make_double2(a.x + b.x, 0.0); }
__device__ __forceinline__ double2 slot_add_recv(cdbl, double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 slot_add_recv_dd(double2 a, double2 b) {
  const ddbl r = dd_add(ddbl(a.x, a.y), ddbl(b.x, b.y));
  return make_double2(r.hi, r.lo);
}
__device__ __forceinline__ double2 slot_add_recv(dot2, double2 a, double2 b) { return slot_add_recv_dd(a, b); }
__device__ __forceinline__ double2 slot_add_recv(ddbl, double2 a, double2 b) { return slot_add_recv_dd(a, b); }

// NCT = register accumulators per consumer thread = ceil(j / TG): the thread's columns are g, g + TG, ...
// SETS = independent consumer sets.  SETS = 1: all 16 consumer warps work on every tile (2 rows per thread).  SETS = 2 (short
// bases): the warps form two sets of 8 that take alternate tiles, a thread owning 4 rows of its tile.  A short basis makes a
// small tile, and what a warp spends per tile is then mostly a latency chain (wait -> LDS -> FP64 chain -> named barrier ->
// ...) that does not shrink with the tile: twice the rows per thread halve the chains per byte, and the two sets hide
// each other's waits (profiles/r02/README.md: j = 11 ran at 4.7 TB/s against 6.7 TB/s at j = 28).
template <class T, int SEG, int NCT, bool PLAIN, int SETS>
__global__ void __launch_bounds__(GsCfg<T, SEG>::THREADS, 1)
k_gs_step(const T* __restrict__ V, i64 ldv, int j, T* wr, i64 ntiles, DevCtl* ctl, double2* partial, int nstages,
          int ncv, const DevComm* cm, unsigned long long* trace, int xprefetch, const __grid_constant__ CUtensorMap vmap) {
  typedef GsCfg<T, SEG> C;
  // 1 KiB column segments (tiles of half the rows: used when fewer than three 2 KiB stages fit) are moved by ONE tensor-map
  // copy per tile (box = R rows x j columns): 1 KiB 1-D bulk copies are capped by the TMA issue rate (4.3 TB/s chip-wide,
  // profiles/microbench/tma_bulk_rate.cu), a box copy is one instruction whatever its row length
  constexpr bool TMAP = (SEG == 1024);
  constexpr int XS = (int)(sizeof(T) / (sizeof(T) == 4 ? 4 : 8));   // map elements per row (Float32 maps count floats, the others doubles)
  typedef typename StepAcc<T, PLAIN>::type acc_t;
#define GA_TRACE(slot, cond)                                                        \
  if (trace != nullptr && (cond)) {                                                 \
    unsigned long long t_;                                                          \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                          \
    trace[slot] = t_;                                                               \
  }
  GA_TRACE(0, blockIdx.x == 0 && threadIdx.x == 0)
  constexpr int R = C::R, TG = C::TG, NC = C::NC, CWARPS = C::CWARPS;
  constexpr int ROWL = C::ROWL / SETS;          // row lanes of one consumer set
  constexpr int RPT = R / ROWL;                  // rows per thread
  constexpr int SLICES = ROWL / 32;              // warps per column group inside a set
  constexpr int SWARPS = CWARPS / SETS;          // consumer warps per set
  static_assert(SETS == 1 || SETS == 2, "consumer sets");
  static_assert(SLICES >= 1 && RPT * ROWL == R, "thread layout");
  constexpr uint32_t SEGB = C::SEG_BYTES;
  constexpr bool CPLX = Elem<T>::is_complex;
  static_assert(NCT >= 1 && NCT <= NC, "columns per thread");

  if (__ldcg(&ctl->status) != ST_OK) return;  // uniform across the grid: nobody reaches a barrier

  extern __shared__ __align__(1024) unsigned char smem[];
  const uint32_t stage_bytes = (uint32_t)(j + 1) * SEGB;
  unsigned char* p = smem + (size_t)nstages * stage_bytes;
  T* hs = reinterpret_cast<T*>(p);                  p += kMaxNcv * 16;
  T* pr = reinterpret_cast<T*>(p);                  p += SETS * 2 * TG * SEGB;
  double2* redw = reinterpret_cast<double2*>(p);    p += CWARPS * (NC + 3) * 16;
  uint64_t* full = reinterpret_cast<uint64_t*>(p);  p += C::MAX_STAGES * 8;
  uint64_t* empty = reinterpret_cast<uint64_t*>(p); p += C::MAX_STAGES * 8;
  int* sh_ctl = reinterpret_cast<int*>(p);          // [0] status, [1] phase

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < nstages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], SWARPS); }
    fence_mbar_init();
  }
  __syncthreads();

  const int set = (warp / SWARPS) % SETS, wset = warp % SWARPS;   // (the producer warp gets set 0; it never uses it)
  const int slice = wset % SLICES, g = wset / SLICES, rl = slice * 32 + lane;
  // rows of a thread: PAIRS of adjacent rows (one LDS.128 for 8-byte elements; 16-byte lane stride = conflict-free), further
  // pairs 2 * ROWL rows apart; 16-byte elements: single rows ROWL apart; Float32: quadruples
  constexpr int RADJ0 = (sizeof(T) == 16) ? 1 : 16 / (int)sizeof(T);        // adjacent rows per 16-byte shared-memory read
  constexpr int RADJ = RPT < RADJ0 ? RPT : RADJ0;
#define ROWIDX(q) (((q) / RADJ) * (ROWL * RADJ) + rl * RADJ + ((q) % RADJ))
  int it = 0;  // running tile counter of this CTA: the stage ring and its parities continue across phases
  int npre = 0;  // producer warp: tiles of the upcoming phase whose V columns are already in flight
  const int nslots = j + 3;
  const int my_tiles = (int)((ntiles - (i64)blockIdx.x + (i64)gridDim.x - 1) / (i64)gridDim.x);

  for (int ph = 0; ph < 5; ++ph) {
    const int mode = (ph == 0 || ph == 3) ? 0 : (ph == 1 ? 1 : 2);
    if (ph >= 1) {
      // verdict of the previous phase (written by CTA 0 before the barrier); identical in every CTA
      if (tid == 0) { sh_ctl[0] = __ldcg(&ctl->status); sh_ctl[1] = __ldcg(&ctl->phase); }
      __syncthreads();
      const int st = sh_ctl[0], phs = sh_ctl[1];
      __syncthreads();
      if (st != ST_OK) break;
      if (ph >= 2 && phs != (ph == 2 ? PH_ORTH1 : PH_ORTH2)) break;   // phases 3 and 4 both belong to the second refinement
    }
    if (mode != 0) {
      // coefficients of this sweep, zero beyond the basis (the branch-free column loops read up to NCT * TG of them)
      const T* hb = hbuf_ptr<T>(ctl);
      constexpr int HN = NCT * TG;
      if (sizeof(T) == 4) {
        for (int c = tid; c < HN; c += blockDim.x) reinterpret_cast<float*>(hs)[c] = c < j ? __ldcg(reinterpret_cast<const float*>(hb) + c) : 0.0f;
      } else if (sizeof(T) == 8) {
        for (int c = tid; c < HN; c += blockDim.x) reinterpret_cast<double*>(hs)[c] = c < j ? __ldcg(reinterpret_cast<const double*>(hb) + c) : 0.0;
      } else {
        for (int c = tid; c < HN; c += blockDim.x)
          reinterpret_cast<double2*>(hs)[c] = c < j ? __ldcg(reinterpret_cast<const double2*>(hb) + c) : make_double2(0.0, 0.0);
      }
      __syncthreads();
    }
    // position in the stage ring at the start of this phase (one division per phase, none per tile)
    int s0 = it % nstages;
    uint32_t par0 = (uint32_t)(it / nstages) & 1u;

    if (warp == CWARPS) {
      // ---------------- producer warp.  The first `npre` tiles of this phase already have their V
      // columns in flight (issued while the previous phase was being reduced); they only lack the
      // column of the vector being orthogonalised, which the previous phase has just rewritten.
      int s = s0;
      uint32_t par = par0;
      int k = 0;
      for (i64 t = blockIdx.x; t < ntiles; t += gridDim.x, ++k) {
        unsigned char* st = smem + (size_t)s * stage_bytes;
        const i64 row0 = t * R;
        if (k < npre) {
          if (lane == 0) {
            mbar_arrive_expect_tx(&full[s], SEGB);
            bulk_g2s(st + (size_t)j * SEGB, wr + row0, SEGB, &full[s]);
          }
        } else {
          mbar_wait(&empty[s], par ^ 1u);
          if (lane == 0) mbar_arrive_expect_tx(&full[s], stage_bytes);
          __syncwarp();
          if (TMAP) {
            if (lane == 0) {
              tma_load_2d(st, &vmap, (int)(row0 * XS), 0, &full[s]);
              bulk_g2s(st + (size_t)j * SEGB, wr + row0, SEGB, &full[s]);
            }
          } else {
            for (int c = lane; c <= j; c += 32) {
              const T* src = (c < j) ? (V + (size_t)c * ldv + row0) : (wr + row0);
              bulk_g2s(st + (size_t)c * SEGB, src, SEGB, &full[s]);
            }
          }
        }
        if (++s == nstages) { s = 0; par ^= 1u; }
      }
      // prefetch for the next phase: V does not change between the phases of a step
      npre = 0;
      if (ph < 4 && xprefetch) {
        for (i64 t = blockIdx.x; t < ntiles && npre < nstages; t += gridDim.x, ++npre) {
          mbar_wait(&empty[s], par ^ 1u);
          if (lane == 0) mbar_expect_tx(&full[s], (uint32_t)j * SEGB);
          __syncwarp();
          unsigned char* st = smem + (size_t)s * stage_bytes;
          const i64 row0 = t * R;
          if (TMAP) {
            if (lane == 0) tma_load_2d(st, &vmap, (int)(row0 * XS), 0, &full[s]);
          } else {
            for (int c = lane; c < j; c += 32) bulk_g2s(st + (size_t)c * SEGB, V + (size_t)c * ldv + row0, SEGB, &full[s]);
          }
          if (++s == nstages) { s = 0; par ^= 1u; }
        }
      }
    } else {
      // ---------------- consumers
      acc_t acc[NCT];
#pragma unroll
      for (int i = 0; i < NCT; ++i) acc[i] = acc_t(Elem<T>::zero());
      NormAcc nacc;
      nacc.init();
      // element offsets of this thread's columns inside a stage; columns beyond the basis alias the vector column j
      // (finite data; coefficient 0 in the update, accumulator ignored in the reduction)
      int coff[NCT];
#pragma unroll
      for (int i = 0; i < NCT; ++i) { const int c = g + TG * i; coff[i] = (c < j ? c : j) * R; }
      // this set's tiles: the set-th, (set + SETS)-th, ... of the CTA's sequence, i.e. every SETS-th ring stage
      int s = s0 + set;
      uint32_t par = par0;
      if (s >= nstages) { s -= nstages; par ^= 1u; }
      bool first = true;
      int kk = 0;   // tiles done in this phase (parity = which half of the partial-sum exchange buffer)
      T* prs = pr + (size_t)set * 2 * TG * R;
      for (i64 t = (i64)blockIdx.x + (i64)set * gridDim.x; t < ntiles; t += (i64)SETS * gridDim.x) {
        mbar_wait(&full[s], par);
        GA_TRACE(1 + 8 * ph, blockIdx.x == 0 && tid == 0 && first)
        first = false;
        const T* tile = reinterpret_cast<const T*>(smem + (size_t)s * stage_bytes);
        T rr[RPT];
#pragma unroll
        for (int q = 0; q < RPT; ++q) rr[q] = tile[(size_t)j * R + ROWIDX(q)];
        if (mode != 0) {
          T p0[RPT];
#pragma unroll
          for (int q = 0; q < RPT; ++q) p0[q] = Elem<T>::zero();
#pragma unroll
          for (int i = 0; i < NCT; ++i) {
            const T hc = hs[g + TG * i];
#pragma unroll
            for (int q = 0; q < RPT; ++q) Elem<T>::sub_mul(p0[q], tile[coff[i] + ROWIDX(q)], hc);
          }
          T* prb = prs + (size_t)(kk & 1) * TG * R;
#pragma unroll
          for (int q = 0; q < RPT; ++q) prb[g * R + ROWIDX(q)] = p0[q];
          named_bar_sync(1 + set * SLICES + slice, TG * 32);
#pragma unroll
          for (int g2 = 0; g2 < TG; ++g2) {
#pragma unroll
            for (int q = 0; q < RPT; ++q) Elem<T>::add(rr[q], prb[g2 * R + ROWIDX(q)]);
          }
          if (g == 0) {
#pragma unroll
            for (int q = 0; q < RPT; ++q) wr[t * R + ROWIDX(q)] = rr[q];
          }
        }
        if (mode != 2) {
#pragma unroll
          for (int i = 0; i < NCT; ++i) {
#pragma unroll
            for (int q = 0; q < RPT; ++q) Elem<T>::acc_conj_mul(acc[i], tile[coff[i] + ROWIDX(q)], rr[q]);
          }
        }
        if (g == 0) {
#pragma unroll
          for (int q = 0; q < RPT; ++q) nacc.add(rr[q]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        s += SETS;
        if (s >= nstages) { s -= nstages; par ^= 1u; }
        ++kk;
      }
      GA_TRACE(2 + 8 * ph, blockIdx.x == 0 && tid == 0)
      // Warp-level sums of all NCT accumulators.  A butterfly per value costs 5 double-double additions per lane and value
      // (with error-free dots that was 3 us of FP64 work per phase for the 16 warps of a CTA, on the critical path between
      // two sweeps); here the values are reduced TRANSPOSED: in the round with lane mask m a lane keeps one half of its
      // values and hands the other half to its partner, so after log2(NP) rounds every lane owns one value (index
      // lane / (32 / NP)) summed over a subset of lanes, and ordinary butterfly rounds finish it: NP - 1 + 5 - log2(NP)
      // additions per lane instead of 5 * NCT.
      if (mode != 2) {
        constexpr int NP = NCT <= 1 ? 1 : NCT <= 2 ? 2 : NCT <= 4 ? 4 : NCT <= 8 ? 8 : 16;
        double2 sl[NP];
#pragma unroll
        for (int i = 0; i < NP; ++i) sl[i] = i < NCT ? slot_from_acc(acc[i < NCT ? i : 0]) : make_double2(0.0, 0.0);
        int m = 16;
#pragma unroll
        for (int h = NP / 2; h >= 1; h >>= 1, m >>= 1) {
          const bool up = (lane & m) != 0;
#pragma unroll
          for (int i = 0; i < h; ++i) {
            const double2 send = up ? sl[i] : sl[i + h];
            const double2 keep = up ? sl[i + h] : sl[i];
            sl[i] = slot_add_recv(acc[0], keep, make_double2(shfl_xor_d(send.x, m), shfl_xor_d(send.y, m)));
          }
        }
#pragma unroll
        for (; m >= 1; m >>= 1) sl[0] = slot_add_recv(acc[0], sl[0], make_double2(shfl_xor_d(sl[0].x, m), shfl_xor_d(sl[0].y, m)));
        constexpr int LPV = 32 / NP;                   // lanes per value
        const int idx = lane / LPV;
        if ((lane % LPV) == 0 && idx < NCT && g + TG * idx < j) redw[warp * (NC + 3) + idx] = sl[0];
      }
      if (g == 0) {
        ddbl nn[3] = {nacc.med, nacc.big, nacc.sml};
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {
#pragma unroll
          for (int k = 0; k < 3; ++k) nn[k] = dd_add(nn[k], ddbl(shfl_xor_d(nn[k].hi, m), shfl_xor_d(nn[k].lo, m)));
        }
        if (lane == 0) {
#pragma unroll
          for (int k = 0; k < 3; ++k) redw[warp * (NC + 3) + NC + k] = make_double2(nn[k].hi, nn[k].lo);
        }
      }
      // r was written with generic-proxy stores; the next phase reads it with TMA (async proxy)
      if (mode != 0) asm volatile("fence.proxy.async;" ::: "memory");
    }
    // every warp advances the running tile counter identically
    it += my_tiles;
    __syncthreads();
    // CTA partials -> global
    for (int s = tid; s < nslots; s += blockDim.x) {
      double2 a = make_double2(0.0, 0.0);
      // warps of column group gg: set * SWARPS + gg * SLICES + slice, summed in ascending warp order
      if (s < j) {
        if (mode != 2) {
          const int gg = s % TG, i = s / TG;
          for (int st = 0; st < SETS; ++st)
            for (int sl = 0; sl < SLICES; ++sl) {
              double2 v = redw[(st * SWARPS + gg * SLICES + sl) * (NC + 3) + i];
              a = CPLX ? slot_add<true>(a, v) : slot_add<false>(a, v);
            }
        }
      } else {
        const int k = s - j;
        for (int st = 0; st < SETS; ++st)
          for (int sl = 0; sl < SLICES; ++sl) a = slot_add<false>(a, redw[(st * SWARPS + sl) * (NC + 3) + NC + k]);
      }
      partial[(size_t)blockIdx.x * kMaxSlots + s] = a;
    }
    GA_TRACE(3 + 8 * ph, blockIdx.x == 0 && tid == 0)
    grid_barrier(ctl);                       // all partials are in global memory
    GA_TRACE(4 + 8 * ph, blockIdx.x == 0 && tid == 0)
    if (blockIdx.x == 0) {
      // fixed-order sum over the CTAs, all threads busy: SUB sub-lanes per slot each sum a strided
      // subset of the CTAs (independent loads in flight), then one thread per slot adds the SUB
      // partial sums in order.  redw (the CTA's own scratch) is free at this point.
      constexpr int SUB = 12;
      const unsigned int G = gridDim.x;
      // (j + 3) * SUB slots of 16 bytes = at most 12.6 KiB.  pr is idle between the phases; with 1 KiB segments of 8-byte
      // elements it is only 8 KiB and the tail runs on into redw, which lies directly behind it and is idle as well (its
      // values went into `partial` before the barrier above): pr + redw >= 12.75 KiB for every element type.
      double2* scratch = reinterpret_cast<double2*>(pr);
      static_assert((size_t)2 * TG * SEGB + (size_t)CWARPS * (NC + 3) * 16 >= (size_t)(kMaxNcv + 3) * SUB * 16, "scratch of the CTA-0 reduction");
      for (int idx = tid; idx < nslots * SUB; idx += blockDim.x) {
        const int s = idx / SUB, u = idx % SUB;
        const bool cplx = CPLX && s < j;
        double2 v[13];   // 13 * SUB >= 148 CTAs; all loads are issued before the first add
#pragma unroll
        for (int k = 0; k < 13; ++k) {
          const unsigned int b = u + k * SUB;
          v[k] = b < G ? __ldcg(&partial[(size_t)b * kMaxSlots + s]) : make_double2(0.0, 0.0);
        }
        double2 a = v[0];
#pragma unroll
        for (int k = 1; k < 13; ++k) a = cplx ? slot_add<true>(a, v[k]) : slot_add<false>(a, v[k]);
        for (unsigned int b = u + 13 * SUB; b < G; b += SUB) {   // larger grids (not on B200)
          const double2 w2 = __ldcg(&partial[(size_t)b * kMaxSlots + s]);
          a = cplx ? slot_add<true>(a, w2) : slot_add<false>(a, w2);
        }
        scratch[idx] = a;
      }
      __syncthreads();
      if (tid < nslots) {
        const bool cplx = CPLX && tid < j;
        double2 a = scratch[tid * SUB];
        for (int u = 1; u < SUB; ++u) a = cplx ? slot_add<true>(a, scratch[tid * SUB + u]) : slot_add<false>(a, scratch[tid * SUB + u]);
        ctl->red[tid] = a;
      }
      __syncthreads();
      GA_TRACE(5 + 8 * ph, tid == 0)
      if (tid == 0) { ctl->prof_runs += 1u; ctl->prof_units += (unsigned long long)(j + 1 + (mode != 0 ? 1 : 0)); }
      if (cm) comm_exchange_slots<CPLX>(ctl, cm, j, nslots);
      gs_decide<T>(ctl, ctl->red, j, ncv, ph == 0 ? DK_DOT : ph == 1 ? DK_CGS : ph == 2 ? DK_ORTH1_NODOTS : ph == 3 ? DK_TAKE : DK_ORTH2);
    }
    GA_TRACE(6 + 8 * ph, blockIdx.x == 0 && tid == 0)
    grid_barrier(ctl);                       // verdict + coefficients published
    GA_TRACE(7 + 8 * ph, blockIdx.x == 0 && tid == 0)
  }
  // a gated-off phase leaves prefetched V columns in flight: close those barrier phases and wait
  // for the copies to land before the CTA (and its shared memory) goes away
  if (warp == CWARPS && npre > 0) {
    int s = it % nstages;
    uint32_t par = (uint32_t)(it / nstages) & 1u;
    for (int k = 0; k < npre; ++k) {
      if (lane == 0) mbar_arrive(&full[s]);
      __syncwarp();
      mbar_wait(&full[s], par);
      if (++s == nstages) { s = 0; par ^= 1u; }
    }
  }
#undef GA_TRACE
#undef ROWIDX
}

template <class T, int SEG, int NCT, bool PLAIN, int SETS>
static int launch_step_nct(ga_ctx* c, int j, int nstages) {
  typedef GsCfg<T, SEG> C;
  i64 ntiles = c->n_pad / C::R;
  const size_t stage_bytes = (size_t)(j + 1) * C::SEG_BYTES;
  size_t smem = (size_t)nstages * stage_bytes + C::FIXED_SMEM + (size_t)(SETS - 1) * 2 * C::TG * C::SEG_BYTES;
  static PerDeviceOnce attr_set;
  if (attr_set.need(c->device)) {
    GA_CUDA(c, cudaFuncSetAttribute(k_gs_step<T, SEG, NCT, PLAIN, SETS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGsSmemBytes));
    attr_set.done(c->device);
  }
  int grid = (int)(ntiles < (i64)c->sm_count ? ntiles : (i64)c->sm_count);
  const T* V = (const T*)c->V;
  T* wr = (T*)c->resid;
  i64 ldv = c->n_pad;
  DevCtl* ctl = c->ctl;
  double2* partial = c->partial;
  int ncv = c->ncv;
  const DevComm* cm = c->world > 1 ? c->devcomm : nullptr;
  unsigned long long* trace = c->trace;
  int xprefetch = c->gs_xprefetch ? 1 : 0;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (c->prof) {
    if (c->prof_used + 2 > c->prof_ev.size()) {
      for (int i = 0; i < 2; ++i) { cudaEvent_t e; GA_CUDA(c, cudaEventCreate(&e)); c->prof_ev.push_back(e); }
    }
    e0 = c->prof_ev[c->prof_used]; e1 = c->prof_ev[c->prof_used + 1];
    c->prof_used += 2;
    GA_CUDA(c, cudaEventRecord(e0, c->stream));
  }
  alignas(64) CUtensorMap vmap;
  memset(&vmap, 0, sizeof(vmap));
  if (SEG == 1024) {
    const CUtensorMap* m = gs_tensor_map(c, j, C::R);
    if (!m) return GA_ERR_CUDA;
    vmap = *m;
  }
  void* args[] = {(void*)&V, (void*)&ldv, (void*)&j, (void*)&wr, (void*)&ntiles, (void*)&ctl, (void*)&partial,
                  (void*)&nstages, (void*)&ncv, (void*)&cm, (void*)&trace, (void*)&xprefetch, (void*)&vmap};
  GA_CUDA(c, cudaLaunchCooperativeKernel((const void*)k_gs_step<T, SEG, NCT, PLAIN, SETS>, dim3(grid), dim3(C::THREADS), args, smem, c->stream));
  c->launches += 1;
  if (c->prof) GA_CUDA(c, cudaEventRecord(e1, c->stream));
  return 0;
}

// NCT = ceil(j / TG), one instantiation each in [LO, HI]
template <class T, int SEG, bool PLAIN, int SETS, int LO, int HI>
static int launch_step_pick(ga_ctx* c, int j, int nstages, int nct) {
  if constexpr (LO > HI) {
    return set_err(c, GA_ERR_ARG, "gs: no step-kernel instantiation for this many columns");
  } else {
    if (nct <= LO) return launch_step_nct<T, SEG, LO, PLAIN, SETS>(c, j, nstages);
    return launch_step_pick<T, SEG, PLAIN, SETS, LO + 1, HI>(c, j, nstages, nct);
  }
}

// two-set instantiations exist for NCT <= this (j <= 24 columns for 8-byte elements)
template <class T> constexpr int kGsTwoSetMaxNct() { return (sizeof(T) == 16) ? 3 : 6; }

template <class T, bool PLAIN>
static int launch_step_tp(ga_ctx* c, int j) {
  typedef GsCfg<T, 2048> C2;
  typedef GsCfg<T, 1024> C1;
  const int nct = (j + C2::TG - 1) / C2::TG;
  const size_t budget2 = kGsSmemBytes - C2::FIXED_SMEM;
  int nst = (int)(budget2 / ((size_t)(j + 1) * 2048));
  // 2 KiB segments need two stages: (j + 1) * 4 KiB <= budget, i.e. j <= 51 -> NCT <= NC2MAX
  constexpr int J2MAX = (int)((kGsSmemBytes - C2::FIXED_SMEM) / 4096) - 1;
  constexpr int NC2MAX = (J2MAX + C2::TG - 1) / C2::TG < C2::NC ? (J2MAX + C2::TG - 1) / C2::TG : C2::NC;
  // short bases: two consumer sets (they need the larger exchange buffer and at least four stages)
  constexpr int NCT2SETS = kGsTwoSetMaxNct<T>();
  if (c->gs_two_sets && nct <= NCT2SETS && j < c->gs_seg1024_from) {
    int nst2 = (int)((budget2 - 2 * C2::TG * 2048) / ((size_t)(j + 1) * 2048));
    if (nst2 >= 4) {
      if (nst2 > C2::MAX_STAGES) nst2 = C2::MAX_STAGES;
      return launch_step_pick<T, 2048, PLAIN, 2, 1, NCT2SETS>(c, j, nst2, nct);
    }
  }
  if (nst >= c->gs_min_stages_2k && j < c->gs_seg1024_from) {
    if (nst > C2::MAX_STAGES) nst = C2::MAX_STAGES;
    return launch_step_pick<T, 2048, PLAIN, 1, 1, NC2MAX>(c, j, nst, nct);
  }
  const size_t budget1 = kGsSmemBytes - C1::FIXED_SMEM;
  nst = (int)(budget1 / ((size_t)(j + 1) * 1024));
  if (nst > C1::MAX_STAGES) nst = C1::MAX_STAGES;
  if (nst < 2) return set_err(c, GA_ERR_ARG, "gs: too many columns for the shared-memory pipeline");
  // 1 KiB segments serve every j for which fewer than three 2 KiB stages fit: (j + 1) * 6 KiB > budget
  constexpr int J3MAX = (int)((kGsSmemBytes - C2::FIXED_SMEM) / (3 * 2048)) - 1;
  constexpr int NC1MIN = (J3MAX + 1 + C1::TG - 1) / C1::TG;
  return launch_step_pick<T, 1024, PLAIN, 1, NC1MIN, C1::NC>(c, j, nst, nct);
}


}  // namespace ga
