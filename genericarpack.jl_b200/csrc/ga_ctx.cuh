// Device-side control block and the per-problem context of the B200 Lanczos library.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <string>
#include <vector>

#include "../../include/garpack_b200.h"
#include "ga_types.cuh"

namespace ga {

constexpr int kMaxNcv = GA_MAX_NCV;
constexpr int kRowPad = 256;  // n_local is padded to a multiple of this many rows (rows beyond n stay 0)
constexpr int kMaxSlots = kMaxNcv + 4;  // dots + 3 norm accumulators (+1 spare)
constexpr int kMaxCtas = 1024;
constexpr int kMaxWorld = 8;  // one NVSwitch node

// Step phases written by the on-device decision logic (dsaitr's flags, src/dsaitr.jl:1324-1447)
enum : int { PH_DONE = 0, PH_ORTH1 = 1, PH_ORTH2 = 2 };
enum : int { ST_OK = 0, ST_BREAKDOWN = 1, ST_NUMERIC = 2, ST_COMM = 3,
              ST_CG_DONE = 100 /* the B-solve of the generalised problem converged: later launches of the batch are no-ops */ };

// What the last CTA of a reduction kernel does with the reduced slots.
enum : int {
  DK_NONE = 0,    // leave the reduced slots in ctl->red (multi-GPU: a later kernel decides)
  DK_DOT = 1,     // h <- dots, wnorm <- norm                         (src/dsaitr.jl:1208,1227)
  DK_CGS = 2,     // first CGS update done: H[j,:], rnorm, 0.717 test (:1243-1263,1304,1324)
  DK_ORTH1 = 3,   // first DGKS correction done                        (:1364-1367,1410,1424-1434)
  DK_ORTH2 = 4,   // second DGKS correction done                       (:1424-1446)
  DK_GETV0_DOT = 5,  // getv0 restart: h <- dots, rnorm0 <- norm       (src/dgetv0.jl:630,662)
  DK_GETV0_UPD = 6,  // getv0 restart: rnorm <- norm, h <- dots        (:665,700)
  DK_NORM = 7,    // rnorm <- norm                                    (src/dsaup2.jl:1414)
  DK_TAKE = 8,    // h <- dots, nothing else (bmat = :G: the host runs the step's scalar logic)
  DK_ORTH1_NODOTS = 9   // DK_ORTH1 after a sweep that did not compute the next dots (step kernel: they are only formed if the 0.717 test asks for them)
};

// Small control block in device memory.  Reals are stored as (hi, lo); lo = 0 for F64/C64.
struct DevCtl {
  int status;        // ST_*
  int phase;         // PH_*
  int j_break;       // 1-based step at which rnorm <= 0 was met
  int rstart_j;      // 1-based step that follows a dgetv0 restart (H[j,1] = 0), else 0
  int nrorth, nitref;
  int zero_resid;    // dsaitr set resid = 0, rnorm = 0 (src/dsaitr.jl:1442-1443)
  unsigned int ticket;  // last-block counter
  unsigned int gbar_count, gbar_gen;  // grid barrier of the persistent step kernel
  unsigned int seq;              // sequence number of the fused cross-GPU slot exchange
  unsigned int hseq;             // sequence number of the fused halo push
  unsigned int zseq;             // sequence number of the sharded normal operator's partial-result exchange
  unsigned int prof_runs;        // executed (not phase-gated) Gram-Schmidt sweep launches
  unsigned long long prof_units; // n-vector passes those launches moved: (j+1) reads + 1 write when updating
  double rnorm[2], wnorm[2], rnorm0[2], rnorm1[2];
  double2 red[kMaxSlots];    // reduced slots of the last reduction kernel
  double2 hbuf[kMaxNcv];     // coefficient vector h / s (element type T, padded to 16 B)
  double2 H[2 * kMaxNcv];    // H(:,1) sub-diagonal [0..ncv), H(:,2) diagonal [ncv..2ncv), real type
  // conjugate-gradient B-solve of the generalised problem (k_general.cu); reals as (hi, lo)
  double cg_rz[2], cg_pq[2], cg_bb[2], cg_rr[2], cg_beta[2], cg_tol2[2];
  int cg_iters, cg_pad;
};

// Peer-memory mailbox (one per GPU, IPC-mapped into every peer): the fused compute+collective
// path writes reduction slots / halo entries straight into the peers' memory over NVLink and
// signals with a sequence-numbered flag; no NCCL kernel, no extra launch.
struct Mailbox {
  double2 slots[2][kMaxWorld][kMaxSlots];   // [parity][source rank][slot]
  unsigned int flags[2][kMaxWorld];         // [parity][source rank] = seq when the slots are complete
  unsigned int hflags[kMaxWorld];           // [source rank] = hseq when that peer's halo entries landed
  unsigned int zflags[kMaxWorld];           // [source rank] = zseq when that peer's partial A^H(A x) is complete
  // low-latency slot exchange: every 16-byte slot travels as four 8-byte words {32 data bits, seq};
  // an aligned 8-byte store is a single NVLink transaction, so data and flag arrive together and the
  // receiver needs neither a system-scope fence nor a second round trip for a separate flag
  unsigned long long ll[2][kMaxWorld][4 * kMaxSlots];   // [parity][source rank][4 * slot + word]
};
struct DevComm {
  int rank, world;
  Mailbox* peer[kMaxWorld];                 // peer[rank] = own mailbox
  void* peer_halo[kMaxWorld];               // halo receive buffers of the peers
  int send_off[kMaxWorld + 1];              // my send list, grouped by destination rank
  int dest_off[kMaxWorld];                  // where my entries start in peer q's halo buffer
  int recv_counts[kMaxWorld];
  // sharded normal operator: byte offsets of the two partial-result buffers inside every rank's
  // exported halo allocation (identical on all ranks), and the element offset of the halo part
  long long zp_off[2];
  long long zp_halo_elem;                   // = n_pad: partial results are laid out [own rows | halo columns]
};

// Process-wide cache of large device blocks (>= 32 MiB, exact-size reuse, bounded): a solve creates a context (basis, CSR
// arrays: gigabytes) and destroys it; cudaMalloc / cudaFree of such blocks cost 15-250 ms on B200 (mapping + scrubbing) and sit
// inside the end-to-end time of every solve.  dev_free parks big blocks, dev_alloc hands them back to the next context
// of the same shape.  Blocks handed to other processes (cudaIpc: halo, mailboxes) never come from here.  ga_trim_cache()
// returns everything to the driver; GA_CACHE_BYTES (default 24 GiB) bounds what is kept.
cudaError_t dev_alloc(void** p, size_t bytes);
void dev_free(void* p);
void dev_cache_trim();

struct CsrDev {
  i64 nrows = 0, ncols = 0, nnz = 0;
  int* rowptr = nullptr;    // nrows+1 (32-bit: nnz < 2^31 per rank)
  int* col = nullptr;       // nnz, local column index into the x vector handed to the kernel
  void* val = nullptr;      // nnz elements of T
  int* rowblk = nullptr;    // nblk+1 row-block boundaries for the row-segment kernel (v1)
  int nblk = 0;
  // streamed kernel (k_spmv_stream): row blocks {r0, r1, k0, k1} whose non-zeros fit one pipeline
  // stage, plus the list of rows that are longer than a stage (handled by the v1 long-row path)
  int4* desc = nullptr;
  int ndesc = 0;
  int* longrow = nullptr;   // nlong row indices
  int nlong = 0;
  int tune_groups = 4, tune_stages = 6;  // consumer groups / pipeline stages picked by the upload-time autotune
  void free_all() {
    dev_free(rowptr); dev_free(col); dev_free(val); cudaFree(rowblk); cudaFree(desc); cudaFree(longrow);
    rowptr = col = rowblk = longrow = nullptr; val = nullptr; desc = nullptr; nblk = 0; nnz = 0; ndesc = 0; nlong = 0;
  }
};

}  // namespace ga

// The opaque context of include/garpack_b200.h
struct ga_ctx {
  int dtype = 0;
  size_t esize = 8;          // sizeof(T)
  ga::i64 n_global = 0, n = 0, row0 = 0, n_pad = 0;
  int ncv = 0, ncv_pad = 0;
  int device = 0, rank = 0, world = 1, sm_count = 148;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;

  // big arrays (element type T), all n_pad long with zero padding rows
  void* V = nullptr;       // n_pad x ncv, column-major, ld = n_pad
  void* resid = nullptr;   // n_pad   (also holds w = OP v_j before orthogonalisation, like the reference)
  // The tall-skinny kernels (Gram-Schmidt sweeps, norm, scale) run on the "active vector space":
  // a basis (ld = padded row count, zero padding rows) and the vector being orthogonalised.  It is
  // (V, resid) except while ga_svd_vectors orthonormalises the other side's singular vectors.
  struct VSpace { void* V = nullptr; void* w = nullptr; ga::i64 ld = 0, rows = 0; } vs;
  void* wbuf = nullptr; ga::i64 wbuf_ld = 0; int wbuf_cols = 0;   // other-side singular vectors (svds)
  void* xfull = nullptr;   // normal operator: scratch vector of length max(m, n)
  ga::i64 xfull_len = 0;

  ga::DevCtl* ctl = nullptr;       // device
  ga::DevCtl* ctl_host = nullptr;  // pinned mirror
  double2* partial = nullptr;      // kMaxCtas x kMaxSlots
  double2* gather = nullptr;       // world x kMaxSlots (multi-GPU all-gather target)
  void* qdev = nullptr;            // ncv_pad x ncv_pad reals for apply_shifts / ritz_vectors
  void* zdev = nullptr; ga::i64 zcols = 0;

  // operator
  int op_kind = 0;  // 0 none, 1 simple CSR, 2 normal (A^H A / A A^H), 3 generalised A x = lambda B x (mode 2, bmat = :G)
  ga::CsrDev A, At;  // At: explicit adjoint CSR for the normal operator
  // generalised problem (ArpackSymmetricGeneralizedOp, src/idonow_ops.jl:462-492): B, the A*v / B*r work vector
  // of the B-inner products, and the work vectors + Jacobi preconditioner of the CG solve with B
  ga::CsrDev Bm;
  void* gz = nullptr;                  // n_pad: A v_j (mode 2: WORKD(IVJ)) or B r (WORKD(IPJ))
  void* cg_r = nullptr; void* cg_p = nullptr; void* cg_q = nullptr; void* cg_x = nullptr;   // n_pad each
  void* cg_graph = nullptr;            // cudaGraphExec_t: 8 CG iterations captured once (k_general.cu)
  bool cg_graph_tried = false, cg_use_graph = true;   // GA_CG_GRAPH=0 at ga_create: plain launches (A/B switch)
  size_t cg_graph_launches = 0;
  void* b_dinv = nullptr;              // n_pad reals: 1 / diag(B)
  double cg_tol[2] = {0.0, 0.0};       // relative residual target of the solve (<= 0: 8 eps)
  int cg_maxit = 0;                    // 0: 1000
  ga_solve_fn solve_cb = nullptr; void* solve_user = nullptr;   // optional user solve (device pointers)
  long long solve_count = 0, solve_iters = 0; double solve_max_relres = 0.0;
  // function operators (ga_set_op_callback / ga_set_normal_op_callbacks): the caller's OP*x on device vectors
  ga_op_fn op_cb = nullptr, av_cb = nullptr, atv_cb = nullptr; void* op_user = nullptr;
  const void* gs_dotvec = nullptr;     // when set, the MODE 0 sweep forms V' * (this vector) instead of V' * resid
  ga::i64 normal_m = 0, normal_n = 0; bool normal_left = true;
  std::vector<ga::CsrDev> Wblk;      // normal operator, single GPU: the wide matrix of the second product in column blocks (k_spmv.cu:normal_wide)
  ga::i64 wblk_cols = 0;             // columns per block
  ga::i64 normal_colblock = 0;       // forced block width (GA_NORMAL_COLBLOCK at ga_create; 0 = automatic: 32 MiB of the long vector)

  // collectives (NCCL, loaded at run time)
  void* nccl_comm = nullptr;
  std::string nccl_id;       // 128-byte unique id; the communicator is created on first use
  // halo plan of the row-sharded SpMV (multi-GPU): which local entries each peer needs from me
  // (send_idx grouped by peer) and how many entries I receive from each peer (halo = concatenation
  // in peer order = ascending global column order).
  int nhalo = 0, nsend = 0;
  std::vector<int> send_counts, recv_counts;
  int* send_idx = nullptr;       // device, nsend local row indices
  void* send_buf = nullptr;      // device, nsend elements of T
  void* halo = nullptr;          // device, nhalo elements of T (x entries owned by peers)
  std::vector<int> dest_off;     // where my entries start in each peer's halo
  // sharded normal operator OP = T^H T (rows of the tall T partitioned, see ga_set_normal_csr_dist)
  bool normal_dist = false;
  bool general_dist = false;         // row-sharded generalised problem: A and Bm have the [own | halo] column layout
  ga::i64 mt_local = 0, mt_pad = 0;
  void* zp[2] = {nullptr, nullptr};   // partial results T_loc^H (T_loc x), inside the halo allocation (peer-readable)
  int* zinv = nullptr;                // [world][n_pad]: position of own row i in peer q's halo segment, or -1
  void* zrecv = nullptr;              // NCCL path: the peers' contributions, laid out like send_buf
  unsigned int zpar = 0;              // parity of the partial-result buffer used by the next OP*x
  // fused peer-memory collectives (cudaIpc): own mailbox, device-side descriptor, mapped peers
  ga::Mailbox* mbox = nullptr;
  ga::DevComm* devcomm = nullptr;   // device copy; nullptr -> NCCL fallback (or single GPU)
  ga::DevComm hostcomm{};
  void* ipc_mapped[2 * ga::kMaxWorld] = {nullptr};
  bool ipc_exported = false;        // the halo buffer's handle is out: the sharded operator can no longer be replaced

  // host-side solve state for ga_symeigs_begin/iterate/end
  void* solve = nullptr;

  ga_stats stats{};

  // measurement hooks (bench.py): CUDA events around every Gram-Schmidt sweep launch
  bool prof = false;
  std::vector<cudaEvent_t> prof_ev;  // pairs (start, stop)
  size_t prof_used = 0;
  long long launches = 0;            // kernels launched by this context
  cudaEvent_t tm0 = nullptr, tm1 = nullptr;
  bool coop = false;                    // device supports cooperative launches (persistent step kernel)
  bool use_step_kernel = true;
  bool dot_plain = false;               // GA_OPT_DOT_MODE = GA_DOT_PLAIN: one FMA per product in the Float64 step kernel (default: Dot2)
  int gs_min_stages_2k = 3;             // step kernel: 2 KiB column segments need this many ring stages, else 1 KiB tiles moved by tensor-map copies (GA_GS_MIN2K: 2 = the round-1 rule)
  std::vector<CUtensorMap> gs_maps;     // tensor maps of V, box = R rows x j columns, index j (lazily encoded; k_gs_step.cu)
  std::vector<char> gs_maps_ok;
  int upload_threads = 8;               // host threads of csr_upload's index passes (GA_UPLOAD_THREADS at ga_create)
  bool vq_stream = true;                // V*Q: streaming two-group kernel for up to 16 outputs (GA_VQ_STREAM=0 at ga_create: A/B switch)
  bool gs_two_sets = true;              // step kernel: two consumer sets for short bases (GA_GS_TWO_SETS=0 at ga_create: A/B switch)
  int gs_seg1024_from = 1 << 30;        // step kernel: 1 KiB column segments from this many columns on (default: only when 2 KiB stages no longer fit twice)
  bool gs_xprefetch = true;             // cross-phase V prefetch of the step kernel (GA_GS_XPREFETCH=0 at ga_create: debug A/B switch)
  bool use_spmv_stream = true;          // TMA-staged persistent SpMV (GA_SPMV_V1=1 selects the first kernel)
  unsigned int gs_flip = 0;             // consecutive sweeps traverse the row tiles in opposite directions (L2 reuse)
  unsigned int* spmv_chk = nullptr;     // device, 128 words: stale-stage self-check of the streamed SpMV (debug)
  unsigned long long* trace = nullptr;  // device, 16 globaltimer samples of the last traced sweep (debug)
  std::string err;
};

namespace ga {

int set_err(ga_ctx* c, int code, const std::string& msg);
#define GA_CUDA(ctx, call)                                                                       \
  do {                                                                                           \
    cudaError_t e_ = (call);                                                                     \
    if (e_ != cudaSuccess)                                                                       \
      return ga::set_err((ctx), GA_ERR_CUDA, std::string(__FILE__) + ":" + std::to_string(__LINE__) + " " + #call + ": " + cudaGetErrorString(e_)); \
  } while (0)

// ---- kernel launchers (each in its own .cu) -------------------------------------------------
// k_lcg.cu
int launch_lcg_fill(ga_ctx* c, const i64 iseed[4]);
void lcg_advance_seed(i64 iseed[4], i64 nvalues);
int launch_lcg_fill_advance(ga_ctx* c, i64 iseed[4]);   // fill + the seed the sequential generator leaves behind
// k_vec.cu
int launch_normalize(ga_ctx* c, int jcol, bool push_halo = true);  // vs.V[:,jcol] = vs.w / rnorm (device rnorm)
int launch_norm_resid(ga_ctx* c, int decide_kind);         // ||resid|| -> ctl
int launch_zero_resid(ga_ctx* c);
// k_spmv.cu
int launch_opx(ga_ctx* c, const void* x, void* y, bool gated);  // y = OP x (device vectors)
int spmv_autotune(ga_ctx* c, CsrDev& M, const void* x, void* y);  // time the streamed SpMV's configurations once
int launch_push_halo(ga_ctx* c, const void* x, bool gated = false);                 // fused multi-GPU path: x's halo entries -> the peers (k_vec.cu)
int launch_normal_half(ga_ctx* c, const void* x, void* y);       // y = A x or A^H x (svds post-processing)
int csr_upload(ga_ctx* c, CsrDev& M, i64 nrows, i64 ncols, const int64_t* rowptr, const int32_t* col,
               const void* val, i64 col_shift);
// k_gs.cu
int launch_gs(ga_ctx* c, int mode, int j, int gate_phase, int decide_kind);
int launch_decide(ga_ctx* c, int j, int gate_phase, int decide_kind);
// k_gs_step.cu
int launch_gs_step(ga_ctx* c, int j);
// k_spmv.cu: y = M x for any uploaded CSR matrix (device vectors)
int launch_spmv_mat(ga_ctx* c, const CsrDev& M, const void* x, void* y, bool gated, bool wait_halo = false);
// k_general.cu: generalised problem A x = lambda B x (mode 2, bmat = :G), host-driven scalar logic
int general_upload(ga_ctx* c, i64 ncols, const int64_t* arp, const int32_t* acol, const void* aval, const int64_t* brp,
                   const int32_t* bcol, const void* bval);
void general_free(ga_ctx* c);
int general_opx(ga_ctx* c, const void* x, void* y, void* ax_out);   // y = inv(B) A x; ax_out (may be null) <- A x
int general_bx(ga_ctx* c, const void* x, void* y);                  // y = B x
int launch_dotc(ga_ctx* c, const void* a, const void* b);           // ctl->red[0..1] <- sum conj(a_i) b_i (all ranks)
int general_getv0(ga_ctx* c, int64_t iseed[4], int initv, int j, double* rnorm);
int general_lanczos(ga_ctx* c, int k, int np, void* H, int ldh, double* rnorm, int64_t iseed[4]);
int general_resid_bnorm(ga_ctx* c, double* rnorm_out);              // sqrt|resid' B resid| (src/dsaup2.jl:1383-1411)
// k_update.cu
int launch_vq(ga_ctx* c, int kplusp, int nout, bool do_resid, const void* sigma, const void* beta, int kev);

// cudaFuncSetAttribute applies to the current device only, and contexts may live on several devices and be
// driven from several host threads (no globals: src/README "multithread safe"): one flag per device, set atomically.
struct PerDeviceOnce {
  std::atomic<unsigned long long> mask{0ULL};
  bool need(int dev) const { return ((mask.load(std::memory_order_acquire) >> (dev & 63)) & 1ULL) == 0ULL; }
  void done(int dev) { mask.fetch_or(1ULL << (dev & 63), std::memory_order_release); }
};

const CUtensorMap* gs_tensor_map(ga_ctx* c, int j, int rows);   // k_gs_step.cu
int sync_ctl(ga_ctx* c);  // copy DevCtl to the pinned mirror and synchronise the stream
int status_error(ga_ctx* c, const char* where);  // after sync_ctl: ST_COMM -> GA_ERR_COMM, ST_NUMERIC -> GA_ERR_NUMERIC

}  // namespace ga
