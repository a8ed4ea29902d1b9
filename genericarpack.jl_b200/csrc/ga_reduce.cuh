// Device helpers shared by the reduction kernels:
//   - NormAcc: overflow/underflow-safe, extended-precision sum of squares.  The reference's
//     norm2 (src/arpack-blas-nrm2.jl:272-299, 355-372) accumulates in x87 80-bit registers so it
//     neither overflows nor loses the low bits; here every square is formed exactly (TwoProd)
//     and accumulated in double-double, and Blue's three-range scaling (the LAPACK >= 3.10
//     dnrm2 constants) keeps huge/tiny inputs in range.
//   - slot reductions: every partial result is a 16-byte "slot" (hi,lo) or (re,im); CTAs write
//     their slots to a [cta][slot] array and the last CTA to finish sums them in a fixed order,
//     so results are bit-reproducible run to run.
#pragma once
#include "ga_ctx.cuh"

namespace ga {

// Blue's scaling constants for IEEE double (LAPACK la_constants: dtsml, dtbig, dssml, dsbig)
#define GA_TSML 1.4916681462400413e-154   /* 2^-511 */
#define GA_TBIG 1.9979190722022350e+146   /* 2^486  */
#define GA_SSML 4.4989137945431964e+161   /* 2^537  */
#define GA_SBIG 1.1113793747425387e-162   /* 2^-538 */

struct NormAcc {
  ddbl med, big, sml;
  __device__ __forceinline__ void init() { med = ddbl(0.0, 0.0); big = ddbl(0.0, 0.0); sml = ddbl(0.0, 0.0); }
  __device__ __forceinline__ void add(double x) {
    double ax = fabs(x);
    if (ax > GA_TBIG) {
      double s = ax * GA_SBIG, p, e;
      two_prod(s, s, p, e);
      big = dd_add(big, ddbl(p, e));
    } else if (ax < GA_TSML) {
      if (ax != 0.0) {
        double s = ax * GA_SSML, p, e;
        two_prod(s, s, p, e);
        sml = dd_add(sml, ddbl(p, e));
      }
    } else {
      double p, e;
      two_prod(x, x, p, e);
      // sloppy add is enough here: all terms are non-negative
      double sh, sl;
      two_sum(med.hi, p, sh, sl);
      sl += med.lo + e;
      fast_two_sum(sh, sl, med.hi, med.lo);
    }
  }
  __device__ __forceinline__ void add(float x) { add((double)x); }
  __device__ __forceinline__ void add(cdbl x) { add(x.re); add(x.im); }
  __device__ __forceinline__ void add(ddbl x) {
    double ax = fabs(x.hi);
    if (ax > GA_TBIG) {
      ddbl s(x.hi * GA_SBIG, x.lo * GA_SBIG);
      big = dd_add(big, dd_mul(s, s));
    } else if (ax < GA_TSML) {
      if (ax != 0.0) {
        ddbl s(x.hi * GA_SSML, x.lo * GA_SSML);
        sml = dd_add(sml, dd_mul(s, s));
      }
    } else {
      med = dd_add(med, dd_mul(x, x));
    }
  }
};

// sqrt of the combined accumulators, in double-double (LAPACK dnrm2 combination rule)
__device__ __forceinline__ ddbl norm_finalize(ddbl med, ddbl big, ddbl sml) {
  if (big.hi > 0.0) {
    if (med.hi > 0.0 || med.hi != med.hi) {
      ddbl m(med.hi * GA_SBIG, med.lo * GA_SBIG);
      m = ddbl(m.hi * GA_SBIG, m.lo * GA_SBIG);
      big = dd_add(big, m);
    }
    ddbl r = dd_sqrt(big);
    const double inv = 1.0 / GA_SBIG;
    return ddbl(r.hi * inv, r.lo * inv);
  } else if (sml.hi > 0.0) {
    if (med.hi > 0.0 || med.hi != med.hi) {
      ddbl a = dd_sqrt(med);
      ddbl b = dd_sqrt(sml);
      const double inv = 1.0 / GA_SSML;
      b = ddbl(b.hi * inv, b.lo * inv);
      ddbl ymin = a < b ? a : b, ymax = a < b ? b : a;
      ddbl q = dd_div(ymin, ymax);
      ddbl s = dd_mul(dd_mul(ymax, ymax), dd_add_d(dd_mul(q, q), 1.0));
      return dd_sqrt(s);
    }
    ddbl r = dd_sqrt(sml);
    const double inv = 1.0 / GA_SSML;
    return ddbl(r.hi * inv, r.lo * inv);
  }
  return dd_sqrt(med);
}

// ---- warp shuffles of 16-byte slots ------------------------------------------------------------
__device__ __forceinline__ double shfl_xor_d(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ ddbl warp_sum_dd(ddbl v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) {
    ddbl o(shfl_xor_d(v.hi, m), shfl_xor_d(v.lo, m));
    v = dd_add(v, o);
  }
  return v;
}
__device__ __forceinline__ double2 warp_sum_slot(double v) {  // compensated: returns (hi,lo)
  ddbl r = warp_sum_dd(ddbl(v, 0.0));
  return make_double2(r.hi, r.lo);
}
__device__ __forceinline__ double2 warp_sum_slot(ddbl v) {
  ddbl r = warp_sum_dd(v);
  return make_double2(r.hi, r.lo);
}
__device__ __forceinline__ double2 warp_sum_slot(cdbl v) {  // (re, im), plain tree
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) {
    v.re += shfl_xor_d(v.re, m);
    v.im += shfl_xor_d(v.im, m);
  }
  return make_double2(v.re, v.im);
}

// In-warp tree for the dot accumulators: plain adds for double/complex (32 partials of equal
// magnitude; the compensated arithmetic is kept for the cross-warp / cross-CTA / cross-GPU
// levels where it is free), double-double adds for Double64.
__device__ __forceinline__ double2 warp_sum_slot_fast(double v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v += shfl_xor_d(v, m);
  return make_double2(v, 0.0);
}
__device__ __forceinline__ double2 warp_sum_slot_fast(dot2 v) {   // Dot2 accumulator: renormalise, then double-double tree
  ddbl a;
  two_sum(v.hi, v.lo, a.hi, a.lo);
  return warp_sum_slot(a);
}
__device__ __forceinline__ double2 warp_sum_slot_fast(cdbl v) { return warp_sum_slot(v); }
__device__ __forceinline__ double2 warp_sum_slot_fast(ddbl v) { return warp_sum_slot(v); }

// slot kinds: pair = (hi,lo) compensated sum; cplx = independent (re,im)
template <bool CPLX>
__device__ __forceinline__ double2 slot_add(double2 a, double2 b) {
  if (CPLX) return make_double2(a.x + b.x, a.y + b.y);
  ddbl r = dd_add(ddbl(a.x, a.y), ddbl(b.x, b.y));
  return make_double2(r.hi, r.lo);
}

// Called by all threads of a CTA after it wrote partial[blockIdx.x][0..nslots).  Returns true in
// the last CTA to arrive (after summing all CTAs' slots into ctl->red in a fixed order).
// dot slots [0, ndot) use CPLX_DOT semantics, the remaining slots are (hi,lo) pairs.
template <bool CPLX_DOT>
__device__ __forceinline__ bool last_block_reduce(DevCtl* ctl, const double2* partial, int ndot, int nslots,
                                                  int* sh_flag) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int t = atomicAdd(&ctl->ticket, 1u);
    *sh_flag = (t == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!*sh_flag) return false;
  __threadfence();
  // One warp per slot: lane l sums CTAs l, l+32, ... (independent loads in flight), then a fixed
  // xor-shuffle tree combines the 32 lanes -> same order every run, ~2 us instead of a serial
  // chain of gridDim.x dependent L2 loads per slot.
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int s = warp; s < nslots; s += nwarps) {
    const bool cplx = CPLX_DOT && s < ndot;
    double2 v[8];
    int cnt = 0;
    for (unsigned int b = lane; b < gridDim.x && cnt < 8; b += 32) v[cnt++] = __ldcg(&partial[(size_t)b * kMaxSlots + s]);
    double2 acc = make_double2(0.0, 0.0);
    for (int i = 0; i < cnt; ++i) acc = cplx ? slot_add<true>(acc, v[i]) : slot_add<false>(acc, v[i]);
    for (unsigned int b = lane + 256; b < gridDim.x; b += 32) {  // grids larger than 256 CTAs
      double2 u = __ldcg(&partial[(size_t)b * kMaxSlots + s]);
      acc = cplx ? slot_add<true>(acc, u) : slot_add<false>(acc, u);
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
      double2 o = make_double2(shfl_xor_d(acc.x, m), shfl_xor_d(acc.y, m));
      acc = cplx ? slot_add<true>(acc, o) : slot_add<false>(acc, o);
    }
    if (lane == 0) ctl->red[s] = acc;
  }
  if (threadIdx.x == 0) ctl->ticket = 0u;
  __syncthreads();
  return true;
}

// ---- fused cross-GPU reduction of the slots (multi-GPU, peer memory over NVLink) ----------------
__device__ __forceinline__ void st_sys_u32(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_sys_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double2 ld_sys_d2(const double2* p) {
  double2 v;
  asm volatile("ld.volatile.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_sys_d2(double2* p, double2 v) {
  asm volatile("st.volatile.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}
// Halo entries are written by the peers with P2P stores while the consuming kernel is already resident and
// spinning on the sequence flag: they must not travel through the read-only (.nc) path, whose contract is
// "constant for the kernel's lifetime".  ld.global.cg reads L2 (the point of coherence for peer writes).
template <class T> __device__ __forceinline__ T ld_halo(const T* p) {
  if constexpr (sizeof(T) == 4) {
    const float v = __ldcg(reinterpret_cast<const float*>(p));
    return *reinterpret_cast<const T*>(&v);
  } else if constexpr (sizeof(T) == 8) {
    const double v = __ldcg(reinterpret_cast<const double*>(p));
    return *reinterpret_cast<const T*>(&v);
  } else {
    const double2 v = __ldcg(reinterpret_cast<const double2*>(p));
    return *reinterpret_cast<const T*>(&v);
  }
}
// bounded spin: ~seconds; on expiry the status word flags a communication failure instead of hanging
__device__ __forceinline__ bool spin_until(const unsigned int* p, unsigned int want) {
  for (long long it = 0; it < (1LL << 25); ++it) {
    if (ld_sys_u32(p) == want) return true;
    if (it > 1024) __nanosleep(64);
  }
  return false;
}

__device__ __forceinline__ void st_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Called by all threads of the last CTA after last_block_reduce left this GPU's sums in
// ctl->red.  Every rank stores its slots into every peer's mailbox (P2P stores over NVLink) and
// sums the world's slots in rank order, so ctl->red ends up bit-identical on all GPUs.  This is
// the collective of the Gram-Schmidt sweep, executed inside the sweep kernel itself.
// Low-latency protocol: each 16-byte slot is sent as four 8-byte words {32 data bits, sequence
// number}; the receiver spins on the words themselves (one warp per slot: lane = 4 * source rank +
// word), so one one-way NVLink latency separates the last sender from the result -- no
// system-scope fence, no separate flag round.
template <bool CPLX_DOT>
__device__ __forceinline__ void comm_exchange_slots(DevCtl* ctl, const DevComm* cm, int ndot, int nslots) {
  __shared__ unsigned int sh_seq;
  if (threadIdx.x == 0) sh_seq = ctl->seq + 1u;
  __syncthreads();
  const unsigned int seq = sh_seq, par = seq & 1u;
  const int rank = cm->rank, world = cm->world;
  const unsigned int* red32 = reinterpret_cast<const unsigned int*>(ctl->red);
  for (int i = threadIdx.x; i < 4 * nslots; i += blockDim.x) {
    const unsigned long long w = ((unsigned long long)seq << 32) | (unsigned long long)red32[i];
    for (int q = 0; q < world; ++q) st_sys_u64(&cm->peer[q]->ll[par][rank][i], w);
  }
  __syncthreads();   // everybody has read ctl->red before it is overwritten below
  const Mailbox* me = cm->peer[rank];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int q = lane >> 2, e = lane & 3;
  bool ok = true;
  for (int s = warp; s < nslots; s += nwarps) {
    unsigned int word = 0u;
    if (q < world) {
      const unsigned long long* p = &me->ll[par][q][4 * s + e];
      unsigned long long w = ld_sys_u64(p);
      for (long long it = 0; (unsigned int)(w >> 32) != seq; ++it) {
        if (it > (1LL << 24)) { ok = false; break; }
        if (it > 4096) __nanosleep(32);
        w = ld_sys_u64(p);
      }
      word = (unsigned int)w;
    }
    // lanes 4q..4q+3 hold x.lo32, x.hi32, y.lo32, y.hi32 of source rank q
    const unsigned int up = __shfl_down_sync(0xffffffffu, word, 1);
    const double val = __hiloint2double((int)up, (int)word);      // valid in lanes with e = 0 (x) and e = 2 (y)
    const bool cplx = CPLX_DOT && s < ndot;
    double2 acc = make_double2(0.0, 0.0);
    for (int r = 0; r < world; ++r) {                              // rank order: identical on every GPU
      const double vx = __shfl_sync(0xffffffffu, val, 4 * r), vy = __shfl_sync(0xffffffffu, val, 4 * r + 2);
      acc = cplx ? slot_add<true>(acc, make_double2(vx, vy)) : slot_add<false>(acc, make_double2(vx, vy));
    }
    if (lane == 0) ctl->red[s] = acc;
  }
  if (!ok) ctl->status = ST_COMM;
  if (threadIdx.x == 0) ctl->seq = seq;
  __syncthreads();
}

}  // namespace ga
