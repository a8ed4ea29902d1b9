// Micro-benchmark: streaming rate of 1-D TMA bulk copies (cp.async.bulk global->shared) as a
// function of the copy size, one producer warp per CTA, 148 persistent CTAs.  Informs the tile
// shape of the Gram-Schmidt sweep kernel (k_gs.cu).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_bulk_rate tma_bulk_rate.cu && ./tma_bulk_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../genericarpack.jl_b200/csrc/ga_tma.cuh"
using namespace ga;

// each "tile" = NCOPY copies of COPY bytes taken from NCOPY different columns (stride ld bytes)
__global__ void __launch_bounds__(160, 1) k_stream(const char* __restrict__ base, size_t ld_bytes, int ncopy, int copy_bytes,
                                               long long ntiles, int nstages, int issue_lanes, double* sink) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t full[16], empty[16];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t stage_bytes = (uint32_t)ncopy * copy_bytes;
  if (tid == 0) { for (int s = 0; s < nstages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 4); } fence_mbar_init(); }
  __syncthreads();
  double acc = 0;
  if (warp == 4) {
    int it = 0;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      const int s = it % nstages; const uint32_t ph = (it / nstages) & 1;
      mbar_wait(&empty[s], ph ^ 1);
      if (lane == 0) mbar_arrive_expect_tx(&full[s], stage_bytes);
      __syncwarp();
      for (int c = lane; c < ncopy; c += issue_lanes) {
        if (lane < issue_lanes)
          bulk_g2s(smem + (size_t)s * stage_bytes + (size_t)c * copy_bytes, base + (size_t)c * ld_bytes + (size_t)t * copy_bytes, copy_bytes, &full[s]);
      }
    }
  } else {
    int it = 0;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      const int s = it % nstages; const uint32_t ph = (it / nstages) & 1;
      mbar_wait(&full[s], ph);
      acc += reinterpret_cast<const double*>(smem + (size_t)s * stage_bytes)[tid];
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
  }
  if (acc == 12345.678) sink[0] = acc;
}

int main() {
  const size_t ncols = 32;
  const size_t col_bytes = 128ull << 20;  // 128 MiB per column -> 4 GiB total
  char* buf; double* sink;
  cudaMalloc(&buf, ncols * col_bytes); cudaMemset(buf, 0, ncols * col_bytes); cudaMalloc(&sink, 8);
  cudaFuncSetAttribute(k_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  printf("copy_bytes ncopy stages lanes  GB/s\n");
  for (int lanes : {32, 1}) for (int copy : {512, 1024, 2048, 4096, 8192, 16384}) for (int ncopy : {8, 32}) {
    size_t stage = (size_t)copy * ncopy;
    int nst = (int)((192 * 1024) / stage); if (nst > 12) nst = 12; if (nst < 1) continue;
    long long ntiles = col_bytes / copy;
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      k_stream<<<148, 160, nst * stage>>>(buf, col_bytes, ncopy, copy, ntiles, nst, lanes, sink);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    printf("%9d %5d %6d %5d  %8.1f %s\n", copy, ncopy, nst, lanes, (double)ncopy * col_bytes / best / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
