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
    default: k_lcg_fill<ddbl, 1><<<(int)blocks, threads, 0, c->stream>>>((ddbl*)c->resid, c->n, c->row0, seed, tab); break;
  }
  GA_CUDA(c, cudaGetLastError());
  c->launches += 1;
  return 0;
}

}  // namespace ga
