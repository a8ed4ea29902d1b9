/*
 * garpack_b200.h -- C ABI of the B200 (sm_100a) implicitly-restarted-Lanczos library.
 *
 * This is the drop-in boundary for GenericArpack.jl's hot path.  Every entry point replaces
 * one reference routine (citations: /root/reference/<file>:<lines>, GenericArpack.jl v0.2.1)
 * and is what the reference's Julia side binds with `ccall` from a method specialised on a
 * `B200ArpackState <: AbstractArpackState` / `B200ArpackOp <: ArpackOp` (INTEGRATION.md shows
 * the exact stubs).  Plain pointers and sizes only; no C++/torch types; no exceptions cross the
 * boundary.  All state lives in the opaque context (the reference keeps all formerly-static
 * Fortran state in ArpackState, src/macros.jl:358-366), so independent contexts are re-entrant.
 *
 * Conventions
 *   - return value: 0 = ok; >0 informational (ARPACK `info`); <0 error.  -1..-13 and -9999
 *     mirror dsaupd's codes (src/dsaupd.jl:885-917); GA_ERR_* are library errors.
 *   - element types: GA_F64 (double), GA_C64 (re,im pairs), GA_DD (hi,lo pairs = Julia Double64),
 *     GA_F32 (float vectors with a Float64 factorisation: the reference's mixed-precision
 *     symeigs(Float32, Float64, op, nev), src/interface.jl:26-34, README.md:28-35; V, resid, the CSR
 *     values and every host vector/matrix of elements are `float`; modes 1 and normal op, one or several GPUs --
 *     local blocks are padded to 512 rows instead of 256, which the halo plan's column numbering follows).
 *   - "real" scalars (rnorm, H, Q, tol, shifts, Ritz values) are `double` for GA_F64/GA_C64/GA_F32 and
 *     (hi,lo) pairs of double for GA_DD; pointers to them are `void*`/`double*` documented per call.
 *   - host pointers are caller-owned; the library never keeps them after the call returns.
 *   - matrices are column-major (Julia layout) with explicit leading dimensions.
 */
#ifndef GARPACK_B200_H
#define GARPACK_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ga_ctx ga_ctx;

enum { GA_F64 = 0, GA_C64 = 1, GA_DD = 2, GA_F32 = 3 };
enum { GA_LM = 0, GA_SM = 1, GA_LA = 2, GA_SA = 3, GA_BE = 4 }; /* which::Symbol */

enum {
  GA_ERR_CUDA = -100,      /* a CUDA runtime call failed; see ga_last_error */
  GA_ERR_ARG = -101,       /* invalid argument (sizes, NULL, unsupported ncv) */
  GA_ERR_NOOP = -102,      /* no operator uploaded */
  GA_ERR_COMM = -103,      /* collective layer failure */
  GA_ERR_NUMERIC = -104,   /* non-finite value met on the device */
  GA_ERR_SOLVE = -105,     /* generalised problem: the solve with B failed (B not positive definite / no convergence) */
  GA_ERR_TRIDIAG = -8      /* dsaupd's code for a failed tridiagonal eigen-solve */
};

#define GA_MAX_NCV 64

/* Counters and timers of ArpackStats (src/macros.jl:173-198).  Times are seconds measured
 * with CUDA events / host clocks inside the library. */
typedef struct ga_stats {
  int64_t nopx, nbx, nrorth, nitref, nrstrt;
  double taupd, taup2, taitr, teigt, tgets, tapps, tconv, tmvopx, tmvbx, tgetv0, titref, trvec;
} ga_stats;

/* ---- context ------------------------------------------------------------------------------
 * One context = one eigenproblem's device state on one GPU: V (n_local x ncv), resid, scratch,
 * the operator.  Replaces the allocations of symeigs (src/interface.jl:245-251).
 *   n_local   rows owned by this rank;   row_offset  first global row;   n_global  total rows.
 *   Single GPU: n_local = n_global, row_offset = 0, rank 0 of 1.
 */
int ga_create(ga_ctx** out, int dtype, int64_t n_global, int64_t n_local, int64_t row_offset, int ncv,
              int device, int rank, int world_size);
void ga_destroy(ga_ctx* ctx);
const char* ga_last_error(const ga_ctx* ctx);
int ga_device_sm_count(const ga_ctx* ctx);

/* Per-context options (no process-global state).
 *   GA_OPT_DOT_MODE  how the Float64 Gram-Schmidt sweeps accumulate h = V'w (the reference calls LinearAlgebra.mul! ->
 *                    OpenBLAS dgemv, src/dsaitr.jl:1227,1352, whose summation order is unspecified):
 *       GA_DOT_EXACT (default)  error-free products and sums (Dot2): h is the exact dot product rounded once, so it does
 *                               not depend on the row partition -- Lanczos trajectories and restart counts are identical
 *                               on 1/2/4/8 GPUs and equal to the oracle's (oracle/oracle.hpp ARITH_EXACT).
 *       GA_DOT_PLAIN            one fused multiply-add per product (what a BLAS dgemv does); ~15 % faster sweeps, last
 *                               bits of h depend on the partition.  Other element types are not affected. */
enum { GA_OPT_DOT_MODE = 1 };
enum { GA_DOT_EXACT = 0, GA_DOT_PLAIN = 1 };
/* The library keeps large device blocks (basis, CSR arrays; >= 32 MiB) of destroyed contexts for the next context of the same
 * shape (cudaMalloc / cudaFree of gigabytes cost 15-250 ms per solve otherwise); at most GA_CACHE_BYTES (environment,
 * default 24 GiB) are kept.  ga_trim_cache returns all of them to the driver. */
void ga_trim_cache(void);
int ga_set_option(ga_ctx* ctx, int key, int64_t value);
int ga_get_option(const ga_ctx* ctx, int key, int64_t* value);

/* Multi-GPU plumbing.  The caller owns rendezvous (torch.distributed / MPI / Julia Distributed):
 * rank 0 calls ga_comm_unique_id, broadcasts the 128 bytes, every rank calls ga_comm_init.
 * Collectives are then issued by the library on its own stream (NCCL over NVLink 5 / NVSwitch). */
int ga_comm_unique_id(void* id128);
int ga_comm_init(ga_ctx* ctx, const void* id128);
/* Fused compute+collective path over NVSwitch peer memory (optional, <= 8 GPUs of one node):
 * every rank exports 128 bytes (two cudaIpc handles: its mailbox and its halo buffer; call after
 * ga_set_csr_dist), the caller all-gathers them, every rank imports the world's handles.  From then
 * on the Gram-Schmidt sweep kernel itself writes its reduction slots into the peers' mailboxes and
 * the normalisation kernel pushes the halo entries into the peers' halo buffers (P2P stores +
 * sequence flags); NCCL is no longer on the per-step path. */
int ga_comm_ipc_export(ga_ctx* ctx, void* out128);
int ga_comm_ipc_import(ga_ctx* ctx, const void* all_handles /* world x 128 bytes */);

/* ---- operator (OP*x of dsaitr step 3: src/dsaitr.jl:1120-1133 -> opx!, src/idonow_ops.jl:74-76)
 * ga_set_csr: ArpackSimpleOp(A) (src/idonow_ops.jl:115-123).  Full (both triangles) CSR of the
 * n_local rows owned by this rank, global column indices, 0-based.  val has the ctx dtype.
 * ga_set_normal_csr: ArpackNormalOp(A) for svds (src/idonow_ops.jl:216-277): A is m x n CSR,
 * OP = A^H A (m > n, problem size n) or A A^H (m <= n, problem size m).  When the long side exceeds 64 MiB of vector the
 * wide matrix of the second product is kept in column blocks whose slice of the intermediate stays L2-resident; the
 * result is bit-identical to the unblocked product.
 * Limits (all ga_set_*csr* calls; GA_ERR_ARG with a message in ga_last_error): per GPU at most 2^31 - 1 rows and
 * 2^31 - 1 stored entries (the device copy of rowptr is 32-bit); rowptr must be non-decreasing and every column index
 * inside [0, ncols) (checked at upload: the kernels gather x[col] unchecked).  The basis is limited to
 * ncv <= GA_MAX_NCV columns (ga_create; the Gram-Schmidt kernels keep a thread's accumulators in registers). */
int ga_set_csr(ga_ctx* ctx, const int64_t* rowptr, const int32_t* col, const void* val);
/* Row-sharded ArpackSimpleOp (world_size > 1).  The caller (who owns rendezvous) remaps the
 * columns of this rank's rows: own column g -> g - row_offset; remote column -> n_pad + its
 * position in the halo, n_pad = n_block rounded up to 256 rows (512 for GA_F32), n_block = ceil(n_global/world).
 * The halo lists the needed remote columns in ascending global order; recv_counts[q] of them are
 * owned by rank q.  send_idx (grouped by destination, send_counts[q] each) are the LOCAL row
 * indices whose x entries rank q needs.  Per OP*x the library packs and exchanges exactly these
 * entries with one grouped ncclSend/ncclRecv round (NVLink), then runs the local SpMV. */
int ga_set_csr_dist(ga_ctx* ctx, const int64_t* rowptr, const int32_t* col_local, const void* val,
                    int32_t nhalo, const int32_t* recv_counts, const int32_t* send_counts,
                    const int32_t* send_idx, const int32_t* send_dest_off /* [world]: where my entries
                    start in peer q's halo; may be NULL when the fused path is not used */);
int ga_set_normal_csr(ga_ctx* ctx, int64_t m, int64_t n, const int64_t* rowptr, const int32_t* col,
                      const void* val);
/* Row-sharded ArpackNormalOp (world_size > 1): OP = T^H T, T = the tall orientation of A (T = A for
 * m > n, At_side = :left; T = A^H otherwise).  The Lanczos vectors (length min(m, n)) are block-
 * partitioned as usual; this rank additionally owns mt_local consecutive rows of T.  The columns of
 * those rows are remapped exactly as for ga_set_csr_dist ([own block | halo]; for a matrix with
 * scattered columns the halo is nearly all of x, i.e. an all-gather) and the halo-plan arguments mean
 * the same.  Per OP*x: (1) the halo of x moves (pushed by the normalisation kernel over NVLink peer
 * memory, or one grouped ncclSend/ncclRecv round); (2) y_loc = T_loc [x_own | halo]; (3) the partial
 * result z = T_loc^H y_loc is formed in the same [own | halo] layout; (4) reduce-scatter: every rank
 * adds, in rank order, the halo segments of its peers' z that belong to its rows -- read straight out
 * of the peers' memory (fused path) or exchanged with ncclSend/ncclRecv -- to its own part.
 * Fixed order = bit-reproducible, and correct for Double64 (a NCCL sum over hi/lo words is not). */
int ga_set_normal_csr_dist(ga_ctx* ctx, int64_t mt_local, const int64_t* rowptr, const int32_t* col_local,
                           const void* val, int32_t nhalo, const int32_t* recv_counts,
                           const int32_t* send_counts, const int32_t* send_idx, const int32_t* send_dest_off);
/* Function operators: the caller's own OP*x on DEVICE vectors, the counterpart of ArpackSimpleFunctionOp(F, n)
 * (src/idonow_ops.jl:360-368, opx!(y, op, x) = op.F(y, x)) and ArpackNormalFunctionOp(av, atv, m, n) (:406-431,
 * y = atv(av(x)) with the library's scratch vector in between; svds then uses av / atv for the other side's vectors).
 * The callback must ENQUEUE its work on `cuda_stream` (a cudaStream_t; the stream every kernel of this context runs
 * on) and return 0; it is called once per OP*x while the library enqueues a batch of Lanczos steps, so it must not
 * synchronise the device.  x_dev has nx elements, y_dev ny elements of the context's element type; rows beyond ny
 * of y_dev are padding and must not be written.  Single GPU.  ga_set_normal_op_callbacks needs a context created
 * with n_global = min(m, n), like ga_set_normal_csr. */
typedef int (*ga_op_fn)(void* user, const void* x_dev, void* y_dev, int64_t nx, int64_t ny, void* cuda_stream);
int ga_set_op_callback(ga_ctx* ctx, ga_op_fn opx, void* user);
int ga_set_normal_op_callbacks(ga_ctx* ctx, int64_t m, int64_t n, ga_op_fn av, ga_op_fn atv, void* user);
/* Generalised symmetric problem A x = lambda B x in ARPACK's "regular inverse" mode 2 with bmat = :G:
 * the reference's ArpackSymmetricGeneralizedOp(A, S, B) (src/idonow_ops.jl:462-492), reached by
 * eigs(Symmetric(A), Symmetric(B), k) (src/interface.jl:143-163).  A and B are n x n full (both triangles)
 * CSR matrices, B Hermitian positive definite.  OP = inv(B) A; the Krylov basis is B-orthonormal and every
 * norm of dgetv0!/dsaitr!/dsaup2! becomes sqrt(|x' B x|) (src/dsaitr.jl:1195-1204,1298-1300,1405-1407,
 * src/dgetv0.jl:625-627,696-698, src/dsaup2.jl:1383-1411).  The reference's S is any factorisation object the
 * caller brings; here the solve with B is a Jacobi-preconditioned conjugate-gradient iteration on the device
 * (relative residual cg_tol, NULL -> 8 eps of the real type; at most cg_maxit iterations, 0 -> 1000), or the
 * caller's own solver on device pointers via ga_set_solve_callback.  One GPU (several: the _dist variant).  After this call
 * ga_getv0 / ga_lanczos / ga_apply_shifts / ga_symeigs* run mode 2; ga_opx applies inv(B) A. */
int ga_set_generalized_csr(ga_ctx* ctx, const int64_t* a_rowptr, const int32_t* a_col, const void* a_val,
                           const int64_t* b_rowptr, const int32_t* b_col, const void* b_val,
                           const double* cg_tol, int cg_maxit);
/* The same problem row-sharded over the GPUs of a multi-GPU context (world_size > 1): this rank's rows of A and of B
 * with columns remapped as for ga_set_csr_dist against ONE halo plan shared by both matrices (the union of the remote
 * columns A's and B's rows reference), then ga_comm_init / ga_comm_ipc_export / ga_comm_ipc_import as usual (this path
 * needs the peer-memory mailboxes; there is no NCCL-only variant).  Every product pushes its operand's halo to the
 * peers, every B-inner product and CG scalar is reduced over all GPUs inside the kernel that computes it, so all ranks
 * take the same branches (two-GPU check: same restart count and counters as one GPU, tests/dist_gpu_check_general.py). */
int ga_set_generalized_csr_dist(ga_ctx* ctx, const int64_t* a_rowptr, const int32_t* a_col_local, const void* a_val,
                                const int64_t* b_rowptr, const int32_t* b_col_local, const void* b_val,
                                const double* cg_tol, int cg_maxit, int32_t nhalo, const int32_t* recv_counts,
                                const int32_t* send_counts, const int32_t* send_idx, const int32_t* send_dest_off);
/* The S of ArpackSymmetricGeneralizedOp(A, S, B) as a callback: x_dev <- inv(B) rhs_dev, both device vectors of
 * n elements, work enqueued on `cuda_stream` (a cudaStream_t); return 0 on success.  NULL restores the built-in CG. */
typedef int (*ga_solve_fn)(void* user, const void* rhs_dev, void* x_dev, int64_t n, void* cuda_stream);
int ga_set_solve_callback(ga_ctx* ctx, ga_solve_fn fn, void* user);
/* y = B x on host vectors (bx!, src/idonow_ops.jl:490-492). */
int ga_bx(ga_ctx* ctx, const void* x_host, void* y_host);
/* solves with B so far, CG iterations they took, and the largest final relative residual */
int ga_get_solve_stats(const ga_ctx* ctx, int64_t* solves, int64_t* iterations, double* max_relres);
/* y = OP*x on host vectors of the problem size (test / svds post-processing helper). */
int ga_opx(ga_ctx* ctx, const void* x_host, void* y_host);

/* ---- data movement (copyto!(resid, v0) src/interface.jl:260; results of ArpackEigen) --------- */
int ga_upload_resid(ga_ctx* ctx, const void* resid_host);
int ga_download_resid(ga_ctx* ctx, void* resid_host);
int ga_upload_v(ga_ctx* ctx, int col0, int ncols, const void* v_host, int64_t ldv);
int ga_download_v(ga_ctx* ctx, int col0, int ncols, void* v_host, int64_t ldv);

/* ---- dgetv0! (src/dgetv0.jl:504-753) ---------------------------------------------------------
 * iseed[4]: dlarnv seed limbs, in/out (src/macros.jl:291).  initv != 0: resid already holds v0.
 * j: 1-based column being generated; j > 1 orthogonalises against V[:,1:j-1] with <= 5
 * refinements.  rnorm out (1 or 2 doubles).  Returns ierr: 0 or -1 (src/dgetv0.jl:725-729). */
int ga_getv0(ga_ctx* ctx, int64_t iseed[4], int initv, int j, double* rnorm);

/* ---- dsaitr! (src/dsaitr.jl:962-1535), mode 1, bmat = :I --------------------------------------
 * Extends a k-step Lanczos factorisation by np steps entirely on the device: per step
 * v_j = r/rnorm, w = OP v_j, classical Gram-Schmidt + DGKS refinement with the 0.717 tests
 * decided on the device, overflow-safe norms; no host round trip inside the call.
 * H: host (ldh x 2) reals, column 0 = sub-diagonal, column 1 = diagonal; rows k..k+np-1 written.
 * rnorm in/out.  iseed in/out (used only if an invariant subspace forces a restart).
 * Returns info: 0, or j-1 = size of the invariant subspace found (src/dsaitr.jl:1515-1523). */
int ga_lanczos(ga_ctx* ctx, int k, int np, void* H, int ldh, double* rnorm, int64_t iseed[4]);

/* ---- device tail of dsapps! (src/dsapps.jl:813-854) + dsaup2's residual norm (src/dsaup2.jl:1397-1414)
 * Q: host (ldq x kplusp) reals from the host bulge chase.  beta = H[kev+1,1] after the chase.
 * V[:,1:kev] <- V[:,1:kplusp] Q[:,1:kev]; if beta > 0: V[:,kev+1] <- V Q[:,kev+1];
 * resid <- Q[kplusp,kev] resid + beta V[:,kev+1]; rnorm_out = ||resid||. */
int ga_apply_shifts(ga_ctx* ctx, int kev, int np, const void* Q, int ldq, const double* beta,
                    double* rnorm_out);

/* ---- device tail of simple_dseupd! (src/dseupd.jl:1413-1435) ------------------------------------
 * Qh: host (ldq x ncv) reals = H(1)...H(nconv) accumulated on the host (the product the
 * reference applies reflector by reflector with dorm2r!).  V <- V Qh (all ncv columns, like the
 * reference leaves V), Z <- V[:,1:nconv] copied to z_host (ldz >= n_local) when z_host != NULL. */
int ga_ritz_vectors(ga_ctx* ctx, int nconv, const void* Qh, int ldq, void* z_host, int64_t ldz);

/* ---- svds post-processing: eigenvecs_to_singvecs! + orth! (src/idonow_ops.jl:279-333) -----------
 * For an ArpackNormalOp context whose basis columns V[:,perm[i]], i < nvec, hold the wanted
 * eigenvectors of A^H A (At_side = :left) or A A^H (:right) -- e.g. after ga_symeigs_end / ga_ritz_vectors:
 *   W[:,i] = A V[:,perm[i]]  (or A^H V[:,perm[i]]),  norms_out[i] = ||W[:,i]||,  W[:,i] /= norms_out[i]
 *   (_scale_from_to),  then orth!(W).
 * The reference's orth! is Householder QR (_dgeqr2!) + dorg2r! + a column sign fix that makes R's
 * diagonal positive, i.e. the unique thin-QR Q factor; here the same Q is formed on the device by
 * classical Gram-Schmidt with one full reorthogonalisation per column (the fused tall-skinny sweeps
 * of the Lanczos step, overflow-safe norms), all nvec columns without a host round trip.
 * w_host: (other dimension) x nvec, ldw >= other dimension (m for :left, n for :right).
 * norms_out: nvec reals (the caller applies the reference's check_norm thresholds, :300-306).
 * Returns 0, GA_ERR_NOOP for a non-normal context, GA_ERR_NUMERIC if a column vanished. */
int ga_svd_vectors(ga_ctx* ctx, int nvec, const int32_t* perm, void* w_host, int64_t ldw, double* norms_out);

/* ---- small device services used by tests and by the Julia shim ---------------------------------
 * ga_norm2: norm2(TR, resid) (src/arpack-blas-nrm2.jl:355-372) of the device resid. */
int ga_norm2_resid(ga_ctx* ctx, double* out);
/* The norm dsaup2! recomputes after a restart (src/dsaup2.jl:1378-1414): norm2(resid) for bmat = :I,
 * sqrt(|resid' B resid|) for bmat = :G (the reference's ido = 2 exchange + dot + sqrt(abs(.))). */
int ga_bnorm_resid(ga_ctx* ctx, double* out);
int ga_get_stats(const ga_ctx* ctx, ga_stats* out);
int ga_reset_stats(ga_ctx* ctx);

/* ---- host-only pieces of the control flow -------------------------------------------------------
 * The O(ncv^2) scalar work BASELINE.json's north_star keeps on the host (with Julia it is the reference's own
 * code; here csrc/host/ (tridiag.hpp, irl.hpp) mirrors it for the ga_symeigs family).  Exported without any device dependency so that the
 * `-m "not gpu"` tests can hold it against the oracle and the reference's analytic known answers.
 * use_dd = 0: reals are double; != 0: (hi, lo) pairs.  Return 0, GA_ERR_ARG, or GA_ERR_TRIDIAG (-8).
 *   ga_host_tridiag_eig     dstqrb! (src/dstqrb.jl:607-1052; z[n] <- last row of the eigenvector matrix) and, with
 *                           want_full, simple_dsteqr! (src/simple.jl; z[ldz x n] <- eigenvectors); d, e overwritten
 *   ga_host_ritz_and_bounds dseigt! (src/simple.jl:968-1035): H is ldh x 2 (col 0 sub-diagonal, col 1 diagonal)
 *   ga_host_sort_pair       dsortr (src/simple.jl:255-308), companion array always applied
 *   ga_host_select_shifts   dsgets (src/simple.jl:603-702)
 *   ga_host_count_converged dsconv (src/simple.jl:37-80); returns nconv
 *   ga_host_chase_shifts    the bulge chase of dsapps! (src/dsapps.jl:609-805): H updated, Q (ldq x kplusp) formed */
int ga_host_tridiag_eig(int use_dd, int n, void* d, void* e, void* z, int want_full, int ldz);
int ga_host_ritz_and_bounds(int use_dd, const void* rnorm, int n, const void* H, int ldh, void* eig, void* bounds, void* work);
int ga_host_sort_pair(int use_dd, int which, int n, void* x1, void* x2);
int ga_host_select_shifts(int use_dd, int ishift, int which, int kev, int np, void* ritz, void* bounds, void* shifts);
int ga_host_count_converged(int use_dd, int n, const void* ritz, const void* bounds, const void* tol);
int ga_host_chase_shifts(int use_dd, int kev, int np, const void* shift, void* H, int ldh, void* Q, int ldq);

/* ---- measurement hooks (bench.py) -------------------------------------------------------------
 * ga_timer_start/stop: CUDA events recorded on the library's own stream (the stream every kernel of
 * this context is launched on); stop synchronises and returns milliseconds.
 * ga_profile(ctx, 1): additionally bracket every Gram-Schmidt sweep launch with CUDA events.
 * ga_profile_read: gs_ms = summed device time of those launches, gs_runs = launches that were not
 * phase-gated off, gs_units = n-vector passes they moved (algorithmic bytes = units * n * sizeof(T)),
 * launches = kernels launched by this context since ga_profile(ctx, *). */
int ga_timer_start(ga_ctx* ctx);
int ga_timer_stop(ga_ctx* ctx, double* ms);
int ga_profile(ga_ctx* ctx, int enable);
/* ga_bench_opx: `reps` back-to-back OP*x on device vectors (V[:,0] -> resid), CUDA events on the
 * library stream; *ms = average per application.  Kernel-level measurement hook only. */
int ga_bench_opx(ga_ctx* ctx, int reps, double* ms);
/* debug: out16 == NULL enables an in-kernel globaltimer timeline of the Gram-Schmidt sweep; a later
 * call with out16 != NULL reads the 64 samples (ns) of the most recent sweep launch. */
int ga_debug_trace(ga_ctx* ctx, uint64_t* out16);
/* Override the per-matrix autotune of the streamed SpMV (tuning aid): groups in {4, 6} consumer groups,
 * groups <= stages <= 8 pipeline stages; groups == 0 selects the first, non-streamed kernel.  Every
 * configuration computes the same bits. */
int ga_tune_spmv(ga_ctx* ctx, int groups, int stages);
/* debug: out128 == NULL arms a self-check inside the streamed SpMV (every consumer group verifies that the
 * rowptr slice staged by TMA belongs to its row block); a later call with out128 != NULL reads the counter
 * (out128[0]) and up to 15 records of 8 words from out128[8] on. */
int ga_debug_spmv_check(ga_ctx* ctx, uint32_t* out128);
int ga_profile_read(ga_ctx* ctx, double* gs_ms, int64_t* gs_runs, int64_t* gs_units, int64_t* launches);
/* number of Gram-Schmidt kernel launches bracketed since ga_profile(ctx, 1) (the persistent step kernel
 * runs all sweeps of a Lanczos step in one launch, so this is <= gs_runs). */
int ga_profile_gs_launches(ga_ctx* ctx, int64_t* n);
/* per-launch milliseconds of the first min(n, launches) bracketed Gram-Schmidt launches (tuning aid) */
int ga_profile_gs_times(ga_ctx* ctx, float* ms_out, int64_t n);

/* ---- whole solve: symeigs (src/interface.jl:215-290) = dsaupd! + simple_dseupd! -----------------
 * Host control (dsaup2 loop, dsgets, dsconv, dseigt/dstqrb, dsapps bulge chase, dseupd's small
 * eigenproblem) runs on the host inside this call and drives the entry points above; it exists
 * because Julia is not available in the build image -- with Julia the reference's own dsaupd!/
 * dsaup2! call ga_getv0/ga_lanczos/ga_apply_shifts/ga_ritz_vectors directly.
 *   tol: 1 (or 2 for DD) doubles, <= 0 -> eps/2.   v0_host NULL -> dlarnv start vector.
 *   values[nev] (ascending), bounds[nev]: reals.  vectors_host: n_local x nev (ldz) or NULL.
 *   iparam_out[11]: ARPACK iparam (iparam[2] = restarts, iparam[4] = nconv, iparam[8..10] counters).
 * Returns dsaupd/dseupd info: 0 ok, 1 maxiter reached, <0 error. */
int ga_symeigs(ga_ctx* ctx, int which, int nev, const double* tol, int maxiter, const void* v0_host,
               int ritzvec, void* values, void* bounds, void* vectors_host, int64_t ldz,
               int64_t iparam_out[11]);

/* Restartable form used by bench.py: run at most `max_restarts` more outer iterations of the
 * dsaup2 loop of a solve started with ga_symeigs_begin; returns 0 while unconverged, 99 when
 * finished (then call ga_symeigs_end to post-process). */
int ga_symeigs_begin(ga_ctx* ctx, int which, int nev, const double* tol, int maxiter, const void* v0_host);
int ga_symeigs_iterate(ga_ctx* ctx, int max_restarts);
int ga_symeigs_end(ga_ctx* ctx, int ritzvec, void* values, void* bounds, void* vectors_host, int64_t ldz,
                   int64_t iparam_out[11]);

#ifdef __cplusplus
}
#endif
#endif /* GARPACK_B200_H */
