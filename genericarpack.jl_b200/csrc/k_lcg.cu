// dgetv0's random start vector: a bit-exact, embarrassingly parallel replica of LAPACK dlarnv
// (idist = 2) as restated by the reference in _dlarnv_idist_2! (src/arpack-blas.jl:457-544).
//
// The reference walks a 48-bit multiplicative congruential generator sequentially, 128 values
// per block using a table whose row i is a^i mod 2^48 (src/arpack-blas.jl:281-408,
// a = row 1 = (494,322,2508,2549) in 12-bit limbs = 0x1EE1429CC9F5).  Value g (1-based) of a
// call is therefore X_g = seed * a^g mod 2^48, x_g = 2*(X_g / 2^48) - 1, and the seed left behind
// is X_last.  On the device every thread jumps straight to its first value with a square-and-
// multiply over precomputed a^(2^k) and then strides with a^(stride).  The conversion
// R*(IT1 + R*(IT2 + R*(IT3 + R*IT4))), R = 2^-12, is exact in binary64 (48 significant bits), so
// (double)X * 2^-48 reproduces it bit for bit; for Double64 the low word is 0 and for ComplexF64
// element i takes values 2i+1 (re) and 2i+2 (im) (:531-536).
#include "ga_ctx.cuh"

namespace ga {

static constexpr unsigned long long kLcgA = (494ULL << 36) | (322ULL << 24) | (2508ULL << 12) | 2549ULL;
static constexpr unsigned long long kMask48 = (1ULL << 48) - 1ULL;

struct LcgPow { unsigned long long p[64]; };  // a^(2^k) mod 2^48

static LcgPow lcg_pow_table() {
  LcgPow t;
  unsigned long long v = kLcgA;
  for (int k = 0; k < 64; ++k) { t.p[k] = v; v = (v * v) & kMask48; }
  return t;
}
static unsigned long long lcg_apow(unsigned long long e) {
  unsigned long long r = 1, b = kLcgA;
  while (e) { if (e & 1ULL) r = (r * b) & kMask48; b = (b * b) & kMask48; e >>= 1; }
  return r;
}
static unsigned long long seed_to_u64(const i64 s[4]) {
  return (((unsigned long long)s[0] & 4095ULL) << 36) | (((unsigned long long)s[1] & 4095ULL) << 24) |
         (((unsigned long long)s[2] & 4095ULL) << 12) | ((unsigned long long)s[3] & 4095ULL);
}
// iseed <- limbs of seed * a^nvalues (what the sequential generator leaves behind, :543)
void lcg_advance_seed(i64 iseed[4], i64 nvalues) {
  if (nvalues <= 0) return;
  unsigned long long x = (seed_to_u64(iseed) * lcg_apow((unsigned long long)nvalues)) & kMask48;
  iseed[0] = (i64)((x >> 36) & 4095ULL); iseed[1] = (i64)((x >> 24) & 4095ULL);
  iseed[2] = (i64)((x >> 12) & 4095ULL); iseed[3] = (i64)(x & 4095ULL);
}

__device__ __forceinline__ unsigned long long dev_apow(const LcgPow& t, unsigned long long e) {
  unsigned long long r = 1;
#pragma unroll 1
  for (int k = 0; e; ++k, e >>= 1)
    if (e & 1ULL) r = (r * t.p[k]) & kMask48;
  return r;
}
__device__ __forceinline__ double lcg_to_unit(unsigned long long x) {
  const double rv = (double)x * 3.552713678800501e-15;  // X * 2^-48, exact
  return 2.0 * rv - 1.0;                                // :538
}

template <class T> __device__ __forceinline__ T lcg_elem(double a, double b);
template <> __device__ __forceinline__ double lcg_elem<double>(double a, double) { return a; }
template <> __device__ __forceinline__ ddbl lcg_elem<ddbl>(double a, double) { return ddbl(a, 0.0); }
template <> __device__ __forceinline__ cdbl lcg_elem<cdbl>(double a, double b) { cdbl z; z.re = a; z.im = b; return z; }

// x[i], i in [0,n) is global element g0 + i (0-based); PER = draws per element (1 real, 2 complex)
template <class T, int PER>
__global__ void __launch_bounds__(256) k_lcg_fill(T* __restrict__ x, i64 n, i64 g0, unsigned long long seed,
                                                  const __grid_constant__ LcgPow tab) {
  const i64 stride = (i64)gridDim.x * blockDim.x;
  const i64 i0 = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i0 >= n) return;
  // first draw of element (g0+i0) is value number PER*(g0+i0) + 1
  unsigned long long X = (seed * dev_apow(tab, (unsigned long long)(PER * (g0 + i0) + 1))) & kMask48;
  const unsigned long long astride = dev_apow(tab, (unsigned long long)(PER * stride));
  for (i64 i = i0; i < n; i += stride) {
    if (PER == 1) {
      x[i] = lcg_elem<T>(lcg_to_unit(X), 0.0);
    } else {
      const unsigned long long X2 = (X * tab.p[0]) & kMask48;
      x[i] = lcg_elem<T>(lcg_to_unit(X), lcg_to_unit(X2));
    }
    X = (X * astride) & kMask48;
  }
}

// ---- Float32 (src/arpack-blas.jl:457-544 with RT = Float32).  The conversion
// R*(RT(IT1) + R*(RT(IT2) + R*(RT(IT3) + R*RT(IT4)))) is evaluated in binary32: the multiplications by R = 2^-12
// are exact, the three additions round.  When the four limbs round up to exactly 1 (X >= 2^48 - 2^23, about one
// draw in 2^25) the reference adds 2 to every limb of the current block's BASE seed and redraws (:520-525): the
// rest of the stream then continues from the modified seed.  The stream is therefore "pure" (X = s * a^k) between
// such events.  The device fills from a given origin (seed s, first index g_first, k counted from the origin) and
// reports the first index whose conversion hits 1; the host then re-bases the stream there and refills the tail.
__device__ __forceinline__ float lcg_to_unit_f32(unsigned long long x) {
  const float R = 2.44140625e-4f;   // 1/4096
  const float t1 = (float)(unsigned)((x >> 36) & 4095ULL), t2 = (float)(unsigned)((x >> 24) & 4095ULL);
  const float t3 = (float)(unsigned)((x >> 12) & 4095ULL), t4 = (float)(unsigned)(x & 4095ULL);
  return __fmul_rn(R, __fadd_rn(t1, __fmul_rn(R, __fadd_rn(t2, __fmul_rn(R, __fadd_rn(t3, __fmul_rn(R, t4)))))));
}
// Global index i in [i_first, n): value number k0 + (i - i_first) + 1 of the pure stream from `seed`; this rank stores
// only its rows [row0, row0 + n_local).  Every rank scans the WHOLE index range: where the stream is re-based depends on
// every earlier draw, and 16M conversions cost less than exchanging the hits would.
__global__ void __launch_bounds__(256) k_lcg_fill_f32(float* __restrict__ x, i64 n, i64 row0, i64 n_local, i64 i_first, i64 k0,
                                                      unsigned long long seed, const __grid_constant__ LcgPow tab,
                                                      unsigned long long* first_hit) {
  const i64 stride = (i64)gridDim.x * blockDim.x;
  const i64 i0 = i_first + (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i0 >= n) return;
  unsigned long long X = (seed * dev_apow(tab, (unsigned long long)(k0 + (i0 - i_first) + 1))) & kMask48;
  const unsigned long long astride = dev_apow(tab, (unsigned long long)stride);
  for (i64 i = i0; i < n; i += stride) {
    const float rv = lcg_to_unit_f32(X);
    if (rv == 1.0f) atomicMin(first_hit, (unsigned long long)i);
    const i64 il = i - row0;
    if (il >= 0 && il < n_local) x[il] = __fadd_rn(__fmul_rn(2.0f, rv), -1.0f);   // :538
    X = (X * astride) & kMask48;
  }
}

// Float32 start vector (rows [row0, row0 + n) of the n_global-long vector).  `iseed` is advanced to what the sequential generator leaves behind.
static int lcg_fill_f32(ga_ctx* c, i64 iseed[4], const LcgPow& tab) {
  const i64 n = c->n_global;
  const unsigned long long kTwoC = 2ULL * ((1ULL << 36) | (1ULL << 24) | (1ULL << 12) | 1ULL);   // +2 on every limb
  unsigned long long* hit_dev = reinterpret_cast<unsigned long long*>(c->partial);   // scratch, idle here
  unsigned long long base = seed_to_u64(iseed);   // seed of the 128-value block that holds index i_first
  i64 i_first = 0, k0 = 0;                        // first index to (re)fill; its 0-based position inside that block
  for (int round = 0; round < 64; ++round) {
    const unsigned long long none = ~0ULL;
    GA_CUDA(c, cudaMemcpyAsync(hit_dev, &none, sizeof(none), cudaMemcpyHostToDevice, c->stream));
    const int threads = 256;
    i64 blocks = (n - i_first + threads - 1) / threads;
    const i64 maxb = (i64)c->sm_count * 8;
    if (blocks > maxb) blocks = maxb;
    if (blocks < 1) blocks = 1;
    k_lcg_fill_f32<<<(int)blocks, threads, 0, c->stream>>>((float*)c->resid, n, c->row0, c->n, i_first, k0, base, tab, hit_dev);
    GA_CUDA(c, cudaGetLastError());
    c->launches += 1;
    unsigned long long hit = none;
    GA_CUDA(c, cudaMemcpyAsync(&hit, hit_dev, sizeof(hit), cudaMemcpyDeviceToHost, c->stream));
    GA_CUDA(c, cudaStreamSynchronize(c->stream));
    // the pure stream from (base, k0) covers indices i_first .. (hit or n-1); block seeds advance by a^128
    const i64 last = hit == none ? n - 1 : (i64)hit;
    const i64 kk = k0 + (last - i_first);           // 0-based position of `last` counted from the block start of i_first
    const unsigned long long blk_seed = (base * lcg_apow((unsigned long long)((kk / 128) * 128))) & kMask48;
    if (hit == none) {
      // seed left behind = last value drawn (:543)
      const unsigned long long xl = (blk_seed * lcg_apow((unsigned long long)(kk % 128 + 1))) & kMask48;
      iseed[0] = (i64)((xl >> 36) & 4095ULL); iseed[1] = (i64)((xl >> 24) & 4095ULL);
      iseed[2] = (i64)((xl >> 12) & 4095ULL); iseed[3] = (i64)(xl & 4095ULL);
      return 0;
    }
    // redraw index `hit` and everything after it from the re-based block seed (the +2 may itself hit again: loop)
    base = (blk_seed + kTwoC) & kMask48;
    i_first = (i64)hit; k0 = kk % 128;
  }
  return set_err(c, GA_ERR_NUMERIC, "lcg: Float32 reseed loop did not terminate");
}

// resid[0:n) <- dlarnv values for this rank's rows [row0, row0+n)
int launch_lcg_fill(ga_ctx* c, const i64 iseed[4]) {
  static const LcgPow tab = lcg_pow_table();
  const unsigned long long seed = seed_to_u64(iseed);
  const int threads = 256;
  i64 blocks = (c->n + threads - 1) / threads;
  const i64 maxb = (i64)c->sm_count * 8;
  if (blocks > maxb) blocks = maxb;
  if (blocks < 1) blocks = 1;
  switch (c->dtype) {
    case GA_F64: k_lcg_fill<double, 1><<<(int)blocks, threads, 0, c->stream>>>((double*)c->resid, c->n, c->row0, seed, tab); break;
    case GA_C64: k_lcg_fill<cdbl, 2><<<(int)blocks, threads, 0, c->stream>>>((cdbl*)c->resid, c->n, c->row0, seed, tab); break;
    case GA_F32: return set_err(c, GA_ERR_ARG, "lcg: Float32 goes through launch_lcg_fill_advance");
    default: k_lcg_fill<ddbl, 1><<<(int)blocks, threads, 0, c->stream>>>((ddbl*)c->resid, c->n, c->row0, seed, tab); break;
  }
  GA_CUDA(c, cudaGetLastError());
  c->launches += 1;
  return 0;
}

// Fill resid and leave in iseed what the reference's sequential generator leaves behind (:543).
int launch_lcg_fill_advance(ga_ctx* c, i64 iseed[4]) {
  if (c->dtype == GA_F32) {
    static const LcgPow tab = lcg_pow_table();
    return lcg_fill_f32(c, iseed, tab);
  }
  int rc = launch_lcg_fill(c, iseed);
  if (rc) return rc;
  lcg_advance_seed(iseed, (c->dtype == GA_C64 ? 2 : 1) * c->n_global);
  return 0;
}

}  // namespace ga
