// Tile / thread-layout configuration shared by the Gram-Schmidt sweep kernels (k_gs.cu: one sweep
// per launch; k_gs_step.cu: all sweeps of a Lanczos step in one persistent cooperative launch).
#pragma once
#include "ga_decide.cuh"
#include "ga_tma.cuh"

namespace ga {

// dynamic shared memory per CTA (the SM has 227 KiB for one CTA incl. ~1 KiB static)
constexpr size_t kGsSmemBytes = 224 * 1024;

template <class T, int SEG> struct GsCfg {
  // A tile is R rows of (j columns + the vector being orthogonalised).  One TMA bulk copy moves
  // one column segment of the tile; measured on B200 (profiles/microbench/tma_bulk_rate.cu) the
  // per-SM bulk-copy issue rate caps 1 KiB copies at ~4.3 TB/s chip-wide while 2 KiB copies reach
  // ~6.8 TB/s, hence 2 KiB segments: R = 256 rows (F64) / 128 rows (C64, DD).
  // SEG = 1024 (tiles of half the rows) serves wide bases, where fewer than three 2 KiB-segment stages would
  // fit in shared memory; those tiles are moved by one tensor-map copy each (k_gs_step.cuh), which is
  // not subject to the per-copy issue rate.
  static constexpr int SEG_BYTES = SEG;
  static constexpr int R = SEG_BYTES / (int)sizeof(T);
  // Thread layout (16 consumer warps + 1 producer warp; 17 warps are allocated as 20, which caps
  // the kernel at 96 registers/thread):
  //   F64 : 128 row lanes x 4 column groups, 2 rows/thread at 2 KiB (1 at 1 KiB), 16 accumulators
  //   16 B: 64 row lanes x 8 column groups, 2 rows/thread at 2 KiB (1 at 1 KiB), 8 accumulators
  //   F32 : 128 row lanes x 4 column groups, 4 adjacent rows/thread at 2 KiB (one LDS.128; 2 at 1 KiB),
  //         16 binary64 accumulators
  static constexpr int ROWL = (sizeof(T) <= 8) ? 128 : 64;
  static constexpr int RPT = R / ROWL;
  static constexpr int TG = 512 / ROWL;
  static constexpr int NC = kMaxNcv / TG;             // columns per thread (register accumulators)
  static constexpr int SLICES = ROWL / 32;
  static constexpr int CWARPS = SLICES * TG;          // consumer warps
  static constexpr int THREADS = CWARPS * 32 + 32;    // + producer warp
  static constexpr int MAX_STAGES = 8;
  static constexpr int FIXED_SMEM = kMaxNcv * 16 /*hs*/ + 2 * TG * SEG_BYTES /*pr*/ + CWARPS * (NC + 3) * 16 /*red*/ +
                                    2 * MAX_STAGES * 8 /*mbar*/ + 64;
};

}  // namespace ga
