// C ABI of the hot path (include/garpack_b200.h): context, operator upload, dgetv0!, dsaitr!,
// the device tails of dsapps! and simple_dseupd!.  Host logic here is only what the reference
// routine itself does around its vector operations (loop over steps, restart handling).
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "ga_ctx.cuh"

namespace ga {
int comm_unique_id(void* id128);
int comm_init(ga_ctx* c, const void* id128);
int comm_ipc_export(ga_ctx* c, void* out128);
int comm_ipc_import(ga_ctx* c, const void* all_handles);
void comm_destroy(ga_ctx* c);
void solve_free(ga_ctx* c);  // ga_host.cu

// ---- cache of large device blocks (see ga_ctx.cuh)
namespace {
struct DevBlock { void* p; size_t bytes; int dev; };
std::mutex g_cache_mu;
std::vector<DevBlock> g_cache_free;                  // parked blocks
std::vector<DevBlock> g_cache_live;                  // big blocks handed out (so that dev_free knows their size)
size_t g_cache_bytes = 0;
constexpr size_t kCacheMinBlock = 32u << 20;
size_t cache_cap() {
  static size_t cap = [] { const char* e = getenv("GA_CACHE_BYTES"); return e ? (size_t)strtoull(e, nullptr, 10) : ((size_t)24 << 30); }();
  return cap;
}
}  // namespace
cudaError_t dev_alloc(void** p, size_t bytes) {
  *p = nullptr;
  int dev = 0;
  cudaGetDevice(&dev);
  if (bytes >= kCacheMinBlock) {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    for (size_t i = 0; i < g_cache_free.size(); ++i)
      if (g_cache_free[i].bytes == bytes && g_cache_free[i].dev == dev) {
        *p = g_cache_free[i].p;
        g_cache_bytes -= bytes;
        g_cache_live.push_back(g_cache_free[i]);
        g_cache_free.erase(g_cache_free.begin() + (long)i);
        return cudaSuccess;
      }
  }
  cudaError_t e = cudaMalloc(p, bytes);
  if (e != cudaSuccess && bytes >= kCacheMinBlock) {   // out of memory with blocks parked: give them back and retry once
    cudaGetLastError();
    dev_cache_trim();
    e = cudaMalloc(p, bytes);
  }
  if (e == cudaSuccess && bytes >= kCacheMinBlock) {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    g_cache_live.push_back(DevBlock{*p, bytes, dev});
  }
  return e;
}
void dev_free(void* p) {
  if (!p) return;
  bool big = false;
  {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    for (const DevBlock& b : g_cache_live) big = big || b.p == p;
  }
  if (big) cudaDeviceSynchronize();   // cudaFree would have waited for work still using the block; a parked block must be idle too
  {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    for (size_t i = 0; i < g_cache_live.size(); ++i)
      if (g_cache_live[i].p == p) {
        const DevBlock b = g_cache_live[i];
        g_cache_live.erase(g_cache_live.begin() + (long)i);
        if (g_cache_bytes + b.bytes <= cache_cap()) {
          g_cache_free.push_back(b);
          g_cache_bytes += b.bytes;
          return;
        }
        break;
      }
  }
  cudaFree(p);
}
void dev_cache_trim() {
  std::vector<DevBlock> blocks;
  {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    blocks.swap(g_cache_free);
    g_cache_bytes = 0;
  }
  int cur = 0;
  cudaGetDevice(&cur);
  for (const DevBlock& b : blocks) { cudaSetDevice(b.dev); cudaFree(b.p); }
  cudaSetDevice(cur);
}

int set_err(ga_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg;
  return code;
}
static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static void clear_operator(ga_ctx* c) {   // before a new operator is installed
  for (auto& w : c->Wblk) w.free_all();
  c->Wblk.clear();
  c->wblk_cols = 0;
  c->general_dist = false;
  c->op_kind = 0;
  c->op_cb = c->av_cb = c->atv_cb = nullptr;
  c->op_user = nullptr;
}
static size_t real_size(const ga_ctx* c) { return c->dtype == GA_DD ? 16 : 8; }
// rows are padded to whole Gram-Schmidt tiles: 2 KiB column segments = 256 rows of 8 bytes, 512 rows of Float32
static i64 row_pad(const ga_ctx* c) { return c->dtype == GA_F32 ? 2 * kRowPad : kRowPad; }

int sync_ctl(ga_ctx* c) {
  GA_CUDA(c, cudaMemcpyAsync(c->ctl_host, c->ctl, sizeof(DevCtl), cudaMemcpyDeviceToHost, c->stream));
  GA_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}
// Library-level failures recorded by the kernels in the control block (call after sync_ctl): a peer that never
// raised its flag / a broken grid barrier is GA_ERR_COMM, a non-finite norm GA_ERR_NUMERIC.  ST_BREAKDOWN (an
// invariant subspace, the reference's rnorm == 0 restart) is not an error and is left to the caller.
int status_error(ga_ctx* c, const char* where) {
  const int st = c->ctl_host->status;
  if (st == ST_COMM) return set_err(c, GA_ERR_COMM, std::string(where) + ": a peer GPU did not answer in time (halo flag / mailbox / grid barrier timeout)");
  if (st == ST_NUMERIC) return set_err(c, GA_ERR_NUMERIC, std::string(where) + ": non-finite value met on the device");
  return 0;
}
static void host_real_to_pair(const ga_ctx* c, const double* in, double out[2]) {
  out[0] = in[0];
  out[1] = c->dtype == GA_DD ? in[1] : 0.0;
}
static void pair_to_host_real(const ga_ctx* c, const double in[2], double* out) {
  out[0] = in[0];
  if (c->dtype == GA_DD) out[1] = in[1];
}
}  // namespace ga

using namespace ga;

extern "C" {

int ga_create(ga_ctx** out, int dtype, int64_t n_global, int64_t n_local, int64_t row_offset, int ncv, int device,
              int rank, int world_size) {
  if (!out) return GA_ERR_ARG;
  *out = nullptr;
  if (dtype < 0 || dtype > 3 || n_global <= 0 || n_local <= 0 || ncv <= 0 || ncv > kMaxNcv || world_size < 1 ||
      rank < 0 || rank >= world_size || row_offset < 0 || row_offset + n_local > n_global)
    return GA_ERR_ARG;
  const double t_create0 = now_s();
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return GA_ERR_CUDA;  // no silent CPU fallback
  ga_ctx* c = new ga_ctx();
  c->dtype = dtype;
  c->esize = dtype == GA_F64 ? 8 : (dtype == GA_F32 ? 4 : 16);
  c->n_global = n_global; c->n = n_local; c->row0 = row_offset;
  c->ncv = ncv; c->ncv_pad = kMaxNcv;
  c->device = device; c->rank = rank; c->world = world_size;
  i64 blk = n_local;
  if (world_size > 1) {
    blk = (n_global + world_size - 1) / world_size;  // uniform block partition
    i64 expect = n_global - (i64)rank * blk;
    if (expect > blk) expect = blk;
    if (row_offset != (i64)rank * blk || n_local != expect) { delete c; return GA_ERR_ARG; }
  }
  c->n_pad = (blk + row_pad(c) - 1) / row_pad(c) * row_pad(c);
#define CK(call)                                            \
  do {                                                      \
    cudaError_t e_ = (call);                                \
    if (e_ != cudaSuccess) {                                \
      fprintf(stderr, "ga_create: %s: %s\n", #call, cudaGetErrorString(e_)); \
      ga_destroy(c);                                        \
      return GA_ERR_CUDA;                                   \
    }                                                       \
  } while (0)
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  c->coop = prop.cooperativeLaunch != 0;
  if (const char* e = getenv("GA_NO_STEP_KERNEL")) c->use_step_kernel = !(e[0] == '1');
  if (const char* e = getenv("GA_SPMV_V1")) c->use_spmv_stream = !(e[0] == '1');
  if (const char* e = getenv("GA_GS_XPREFETCH")) c->gs_xprefetch = !(e[0] == '0');
  if (const char* e = getenv("GA_CG_GRAPH")) c->cg_use_graph = !(e[0] == '0');
  if (const char* e = getenv("GA_UPLOAD_THREADS")) c->upload_threads = atoi(e) >= 1 ? atoi(e) : c->upload_threads;
  if (const char* e = getenv("GA_VQ_STREAM")) c->vq_stream = !(e[0] == '0');
  if (const char* e = getenv("GA_GS_TWO_SETS")) c->gs_two_sets = !(e[0] == '0');
  if (const char* e = getenv("GA_GS_MIN2K")) c->gs_min_stages_2k = atoi(e) >= 2 ? atoi(e) : c->gs_min_stages_2k;
  if (const char* e = getenv("GA_NORMAL_COLBLOCK")) c->normal_colblock = atoll(e) > 0 ? atoll(e) : 0;
  if (const char* e = getenv("GA_GS_SEG1024_FROM")) c->gs_seg1024_from = atoi(e) > 0 ? atoi(e) : c->gs_seg1024_from;
  CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CK(cudaEventCreate(&c->ev0));
  CK(cudaEventCreate(&c->ev1));
  CK(dev_alloc(&c->V, (size_t)c->n_pad * ncv * c->esize));
  CK(dev_alloc(&c->resid, (size_t)c->n_pad * c->esize));
  CK(cudaMemsetAsync(c->V, 0, (size_t)c->n_pad * ncv * c->esize, c->stream));
  CK(cudaMemsetAsync(c->resid, 0, (size_t)c->n_pad * c->esize, c->stream));
  c->vs.V = c->V; c->vs.w = c->resid; c->vs.ld = c->n_pad; c->vs.rows = c->n;
  CK(cudaMalloc(&c->ctl, sizeof(DevCtl)));
  CK(cudaMemsetAsync(c->ctl, 0, sizeof(DevCtl), c->stream));
  CK(cudaMallocHost(&c->ctl_host, sizeof(DevCtl)));
  memset(c->ctl_host, 0, sizeof(DevCtl));
  CK(cudaMalloc(&c->partial, sizeof(double2) * (size_t)kMaxCtas * kMaxSlots));
  CK(cudaMalloc(&c->gather, sizeof(double2) * (size_t)world_size * kMaxSlots));
  CK(cudaMalloc(&c->qdev, (size_t)kMaxNcv * kMaxNcv * 16));
  CK(cudaStreamSynchronize(c->stream));
#undef CK
  if (getenv("GA_TIMING")) fprintf(stderr, "ga_create: %.1f ms\n", (now_s() - t_create0) * 1e3);
  *out = c;
  return 0;
}

void ga_destroy(ga_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  solve_free(c);
  comm_destroy(c);
  c->A.free_all();
  c->At.free_all();
  for (auto& w : c->Wblk) w.free_all();
  general_free(c);
  dev_free(c->V); dev_free(c->resid); cudaFree(c->xfull); cudaFree(c->ctl); cudaFree(c->partial);
  cudaFree(c->gather); cudaFree(c->qdev); cudaFree(c->zdev); cudaFree(c->wbuf);
  cudaFree(c->send_idx); cudaFree(c->send_buf); cudaFree(c->halo); cudaFree(c->trace); cudaFree(c->spmv_chk); cudaFree(c->zinv); cudaFree(c->zrecv);
  if (c->ctl_host) cudaFreeHost(c->ctl_host);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->tm0) cudaEventDestroy(c->tm0);
  if (c->tm1) cudaEventDestroy(c->tm1);
  for (cudaEvent_t e : c->prof_ev) cudaEventDestroy(e);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

void ga_trim_cache(void) { dev_cache_trim(); }

int ga_set_option(ga_ctx* c, int key, int64_t value) {
  if (!c) return GA_ERR_ARG;
  switch (key) {
    case GA_OPT_DOT_MODE:
      if (value != GA_DOT_EXACT && value != GA_DOT_PLAIN) return set_err(c, GA_ERR_ARG, "ga_set_option: GA_OPT_DOT_MODE takes GA_DOT_EXACT or GA_DOT_PLAIN");
      c->dot_plain = value == GA_DOT_PLAIN;
      return 0;
    default:
      return set_err(c, GA_ERR_ARG, "ga_set_option: unknown key");
  }
}
int ga_get_option(const ga_ctx* c, int key, int64_t* value) {
  if (!c || !value) return GA_ERR_ARG;
  switch (key) {
    case GA_OPT_DOT_MODE: *value = c->dot_plain ? GA_DOT_PLAIN : GA_DOT_EXACT; return 0;
    default: return GA_ERR_ARG;
  }
}

const char* ga_last_error(const ga_ctx* c) { return c ? c->err.c_str() : "null context"; }
int ga_device_sm_count(const ga_ctx* c) { return c ? c->sm_count : 0; }

int ga_comm_unique_id(void* id128) { return comm_unique_id(id128); }
int ga_comm_init(ga_ctx* c, const void* id128) { return c ? comm_init(c, id128) : GA_ERR_ARG; }
int ga_comm_ipc_export(ga_ctx* c, void* out128) { return (c && out128) ? comm_ipc_export(c, out128) : GA_ERR_ARG; }
int ga_comm_ipc_import(ga_ctx* c, const void* all) { return (c && all) ? comm_ipc_import(c, all) : GA_ERR_ARG; }

// ------------------------------------------------------------------ operator
int ga_set_csr(ga_ctx* c, const int64_t* rowptr, const int32_t* col, const void* val) {
  if (!c || !rowptr || !col || !val) return GA_ERR_ARG;
  GA_CUDA(c, cudaSetDevice(c->device));
  clear_operator(c);
  if (c->world != 1) return set_err(c, GA_ERR_ARG, "multi-GPU contexts take ga_set_csr_dist (local columns + halo plan)");
  const double t0 = now_s();
  int rc = csr_upload(c, c->A, c->n, c->n_global, rowptr, col, val, 0);
  if (rc) return rc;
  c->op_kind = 1;
  const double t1 = now_s();
  rc = spmv_autotune(c, c->A, c->V, c->resid);
  if (getenv("GA_TIMING")) fprintf(stderr, "ga_set_csr: upload %.1f ms, autotune %.1f ms\n", (t1 - t0) * 1e3, (now_s() - t1) * 1e3);
  return rc;
}

// The halo plan of a row-sharded operator (ga_set_csr_dist, ga_set_generalized_csr_dist) indexes device memory directly.
static int validate_halo_plan(ga_ctx* c, const char* who, const int32_t* recv_counts, const int32_t* send_counts, const int32_t* send_idx) {
  int ns = 0;
  for (int q = 0; q < c->world; ++q) {
    if (send_counts[q] < 0 || recv_counts[q] < 0) return set_err(c, GA_ERR_ARG, std::string(who) + ": negative count in the halo plan");
    ns += send_counts[q];
  }
  for (int k = 0; send_idx && k < ns; ++k)
    if (send_idx[k] < 0 || send_idx[k] >= c->n) return set_err(c, GA_ERR_ARG, std::string(who) + ": send_idx outside the local rows");
  return 0;
}
// counts, offsets, send list and the halo buffer the peers push into (exported by ga_comm_ipc_export afterwards)
static int install_halo_plan(ga_ctx* c, int32_t nhalo, const int32_t* recv_counts, const int32_t* send_counts, const int32_t* send_idx,
                             const int32_t* send_dest_off) {
  c->send_counts.assign(send_counts, send_counts + c->world);
  c->recv_counts.assign(recv_counts, recv_counts + c->world);
  if (send_dest_off) c->dest_off.assign(send_dest_off, send_dest_off + c->world);
  else c->dest_off.assign(c->world, 0);
  int nsend = 0, nrecv = 0;
  for (int q = 0; q < c->world; ++q) { nsend += send_counts[q]; nrecv += recv_counts[q]; }
  if (nrecv != nhalo || (nsend > 0 && !send_idx)) return set_err(c, GA_ERR_ARG, "halo plan is inconsistent");
  c->nhalo = nhalo; c->nsend = nsend;
  cudaFree(c->send_idx); cudaFree(c->send_buf); cudaFree(c->halo);
  c->send_idx = nullptr; c->send_buf = nullptr; c->halo = nullptr;
  GA_CUDA(c, cudaMalloc(&c->send_idx, sizeof(int) * (size_t)(nsend > 0 ? nsend : 1)));
  GA_CUDA(c, cudaMalloc(&c->send_buf, c->esize * (size_t)(nsend > 0 ? nsend : 1)));
  GA_CUDA(c, cudaMalloc(&c->halo, c->esize * (size_t)(nhalo > 0 ? nhalo : 1)));
  if (nsend > 0) GA_CUDA(c, cudaMemcpy(c->send_idx, send_idx, sizeof(int) * (size_t)nsend, cudaMemcpyHostToDevice));
  GA_CUDA(c, cudaMemset(c->halo, 0, c->esize * (size_t)(nhalo > 0 ? nhalo : 1)));
  return 0;
}

// Row-sharded operator: this rank's rows with columns already remapped by the caller:
//   own column g (row0 <= g < row0 + n_local)  ->  g - row0
//   remote column                               ->  n_pad + position in the halo
// where the halo lists the needed remote columns in ascending global order (= grouped by owner
// rank), recv_counts[q] of them owned by rank q; send_idx (grouped by destination rank, send_counts[q]
// each) are the LOCAL row indices whose x entries rank q needs from this rank.
int ga_set_csr_dist(ga_ctx* c, const int64_t* rowptr, const int32_t* col_local, const void* val, int32_t nhalo,
                    const int32_t* recv_counts, const int32_t* send_counts, const int32_t* send_idx,
                    const int32_t* send_dest_off) {
  if (!c || !rowptr || !col_local || !val || !recv_counts || !send_counts || nhalo < 0) return GA_ERR_ARG;
  GA_CUDA(c, cudaSetDevice(c->device));
  if (c->ipc_exported || c->devcomm) return set_err(c, GA_ERR_ARG, "the sharded operator cannot be replaced after ga_comm_ipc_export (peers hold the halo buffer's handle): create a new context");
  int rc = validate_halo_plan(c, "ga_set_csr_dist", recv_counts, send_counts, send_idx);
  if (rc) return rc;
  clear_operator(c);
  c->normal_dist = false;
  if ((rc = csr_upload(c, c->A, c->n, c->n_pad + nhalo, rowptr, col_local, val, 0))) return rc;
  if ((rc = install_halo_plan(c, nhalo, recv_counts, send_counts, send_idx, send_dest_off))) return rc;
  c->op_kind = 1;
  return spmv_autotune(c, c->A, c->V, c->resid);
}

// Host transpose (conjugate transpose for ComplexF64) of rows [r0, r1) of a CSR matrix with `ncols` columns: the result
// has `ncols` rows and r1 - r0 columns (local indices).  Counting sort, entries of a row in ascending column order.
static void csr_slab_adjoint(const ga_ctx* c, int64_t r0, int64_t r1, int64_t ncols, const int64_t* rowptr, const int32_t* col,
                             const void* val, std::vector<int64_t>& tp, std::vector<int32_t>& tc, std::vector<char>& tv) {
  const size_t es = c->esize;
  const int64_t k0 = rowptr[r0], nnz = rowptr[r1] - k0;
  tp.assign((size_t)ncols + 1, 0);
  for (int64_t k = 0; k < nnz; ++k) tp[(size_t)col[k0 + k] + 1]++;
  for (int64_t j = 0; j < ncols; ++j) tp[(size_t)j + 1] += tp[(size_t)j];
  tc.resize((size_t)std::max<int64_t>(nnz, 1));
  tv.resize((size_t)std::max<int64_t>(nnz, 1) * es);
  std::vector<int64_t> pos(tp.begin(), tp.end() - 1);
  for (int64_t i = r0; i < r1; ++i)
    for (int64_t k = rowptr[i]; k < rowptr[i + 1]; ++k) {
      const int64_t d = pos[(size_t)col[k]]++;
      tc[(size_t)d] = (int32_t)(i - r0);
      memcpy(&tv[(size_t)d * es], (const char*)val + (size_t)k * es, es);
      if (c->dtype == GA_C64) ((double*)&tv[(size_t)d * es])[1] *= -1.0;
    }
}

int ga_set_normal_csr(ga_ctx* c, int64_t m, int64_t n, const int64_t* rowptr, const int32_t* col, const void* val) {
  if (!c || !rowptr || !col || !val || m <= 0 || n <= 0) return GA_ERR_ARG;
  if (c->world != 1) return set_err(c, GA_ERR_ARG, "multi-GPU contexts take ga_set_normal_csr_dist (this rank's rows of the tall orientation + halo plan)");
  const bool left = !(m <= n);  // src/idonow_ops.jl:222-235
  if (c->n != (left ? n : m)) return set_err(c, GA_ERR_ARG, "normal operator: context size must be min(m, n) side");
  GA_CUDA(c, cudaSetDevice(c->device));
  clear_operator(c);
  // The long vector of the normal operator (max(m, n) entries) is the x of the second product.  Beyond ~64 MiB its
  // random gathers miss L2: the wide matrix is then kept in column blocks of 32 MiB of that vector (k_spmv.cu:normal_wide)
  const i64 longside = left ? m : n;
  i64 blockw = c->normal_colblock > 0 ? c->normal_colblock : ((i64)longside * (i64)c->esize > (64LL << 20) ? (32LL << 20) / (i64)c->esize : 0);
  if (blockw > 0) blockw = (blockw + row_pad(c) - 1) / row_pad(c) * row_pad(c);
  const bool blocked = blockw > 0 && blockw < longside;
  // A and A^H.  Blocked: only the TALL one of the two is uploaded whole (it is the first product, and the half
  // ga_svd_vectors applies); the wide one exists as column blocks = adjoints of row slabs of the tall one.
  int rc = 0;
  std::vector<int64_t> tp;
  std::vector<int32_t> tc;
  std::vector<char> tv;
  if (!blocked || left) {
    if ((rc = csr_upload(c, c->A, m, n, rowptr, col, val, 0))) return rc;
  } else {
    c->A.free_all();
  }
  if (!blocked || !left) {
    csr_slab_adjoint(c, 0, m, n, rowptr, col, val, tp, tc, tv);            // A^H: n x m
    if ((rc = csr_upload(c, c->At, n, m, tp.data(), tc.data(), tv.data(), 0))) return rc;
  } else {
    c->At.free_all();
  }
  if (blocked) {
    const i64 nb = (longside + blockw - 1) / blockw;
    c->Wblk.resize((size_t)nb);
    c->wblk_cols = blockw;
    if (left) {
      // wide = A^H (n x m): column block b = adjoint of rows [b w, (b+1) w) of A
      std::vector<int64_t> bp;
      std::vector<int32_t> bc;
      std::vector<char> bv;
      for (i64 b = 0; b < nb; ++b) {
        const i64 r0 = b * blockw, r1 = std::min<i64>(m, r0 + blockw);
        csr_slab_adjoint(c, r0, r1, n, rowptr, col, val, bp, bc, bv);
        if ((rc = csr_upload(c, c->Wblk[(size_t)b], n, r1 - r0, bp.data(), bc.data(), bv.data(), 0))) return rc;
      }
    } else {
      // wide = A (m x n): column block b = adjoint of rows [b w, (b+1) w) of A^H (held in tp / tc / tv)
      std::vector<int64_t> bp;
      std::vector<int32_t> bc;
      std::vector<char> bv;
      for (i64 b = 0; b < nb; ++b) {
        const i64 r0 = b * blockw, r1 = std::min<i64>(n, r0 + blockw);
        csr_slab_adjoint(c, r0, r1, m, tp.data(), tc.data(), tv.data(), bp, bc, bv);
        if ((rc = csr_upload(c, c->Wblk[(size_t)b], m, r1 - r0, bp.data(), bc.data(), bv.data(), 0))) return rc;
      }
    }
  }
  c->normal_m = m; c->normal_n = n; c->normal_left = left; c->normal_dist = false;
  const i64 extra = left ? m : n;
  const i64 extra_pad = (extra + row_pad(c) - 1) / row_pad(c) * row_pad(c);
  cudaFree(c->xfull);
  c->xfull = nullptr;
  GA_CUDA(c, cudaMalloc(&c->xfull, (size_t)extra_pad * c->esize));
  GA_CUDA(c, cudaMemset(c->xfull, 0, (size_t)extra_pad * c->esize));
  c->xfull_len = extra_pad;
  c->op_kind = 2;
  // autotune the streamed SpMV per matrix.  V's first column / resid / xfull serve as operands.
  // left: x (n) -> A -> xfull (m) -> A^H -> resid;  right: x (m) -> A^H -> xfull (n) -> A -> resid
  CsrDev& tall = left ? c->A : c->At;
  if ((rc = spmv_autotune(c, tall, c->V, c->xfull))) return rc;
  if (!blocked) return spmv_autotune(c, left ? c->At : c->A, c->xfull, c->resid);
  if ((rc = spmv_autotune(c, c->Wblk[0], c->xfull, c->resid))) return rc;
  for (size_t b = 1; b < c->Wblk.size(); ++b) {
    c->Wblk[b].tune_groups = c->Wblk[0].tune_groups;
    c->Wblk[b].tune_stages = c->Wblk[0].tune_stages;
  }
  return 0;
}

// Row-sharded normal operator (see the header): T_loc and its adjoint in the [own | halo] column space.
int ga_set_normal_csr_dist(ga_ctx* c, int64_t mt_local, const int64_t* rowptr, const int32_t* col_local,
                           const void* val, int32_t nhalo, const int32_t* recv_counts, const int32_t* send_counts,
                           const int32_t* send_idx, const int32_t* send_dest_off) {
  if (!c || !rowptr || !col_local || !val || !recv_counts || !send_counts || nhalo < 0 || mt_local < 0) return GA_ERR_ARG;
  if (c->world < 2) return set_err(c, GA_ERR_ARG, "single-GPU contexts take ga_set_normal_csr");
  GA_CUDA(c, cudaSetDevice(c->device));
  if (c->ipc_exported || c->devcomm) return set_err(c, GA_ERR_ARG, "the sharded operator cannot be replaced after ga_comm_ipc_export (peers hold the halo buffer's handle): create a new context");
  clear_operator(c);
  const i64 ncol = c->n_pad + nhalo;                       // local column space [own | halo]
  int rc = csr_upload(c, c->A, mt_local, ncol, rowptr, col_local, val, 0);
  if (rc) return rc;
  // adjoint of T_loc: rows = local column space, columns = local rows of T
  const i64 nnz = rowptr[mt_local] - rowptr[0];
  {
    std::vector<int64_t> tp((size_t)ncol + 1, 0);
    const int64_t base = rowptr[0];
    for (i64 k = 0; k < nnz; ++k) {
      const int32_t cc = col_local[base + k];
      if (cc < 0 || cc >= ncol) return set_err(c, GA_ERR_ARG, "normal operator: column outside [own | halo]");
      tp[(size_t)cc + 1]++;
    }
    for (i64 j = 0; j < ncol; ++j) tp[(size_t)j + 1] += tp[(size_t)j];
    std::vector<int32_t> tc((size_t)std::max<i64>(nnz, 1));
    std::vector<char> tv((size_t)std::max<i64>(nnz, 1) * c->esize);
    std::vector<int64_t> pos(tp.begin(), tp.end() - 1);
    for (i64 i = 0; i < mt_local; ++i)
      for (i64 k = rowptr[i]; k < rowptr[i + 1]; ++k) {
        const i64 d = pos[(size_t)col_local[k]]++;
        tc[(size_t)d] = (int32_t)i;
        memcpy(&tv[(size_t)d * c->esize], (const char*)val + (size_t)k * c->esize, c->esize);
        if (c->dtype == GA_C64) ((double*)&tv[(size_t)d * c->esize])[1] *= -1.0;
      }
    rc = csr_upload(c, c->At, ncol, mt_local, tp.data(), tc.data(), tv.data(), 0);
    if (rc) return rc;
  }
  // halo plan (same bookkeeping as ga_set_csr_dist)
  c->send_counts.assign(send_counts, send_counts + c->world);
  c->recv_counts.assign(recv_counts, recv_counts + c->world);
  if (send_dest_off) c->dest_off.assign(send_dest_off, send_dest_off + c->world);
  else c->dest_off.assign(c->world, 0);
  int nsend = 0, nrecv = 0;
  for (int q = 0; q < c->world; ++q) { nsend += send_counts[q]; nrecv += recv_counts[q]; }
  if (nrecv != nhalo || (nsend > 0 && !send_idx)) return set_err(c, GA_ERR_ARG, "halo plan is inconsistent");
  c->nhalo = nhalo; c->nsend = nsend;
  cudaFree(c->send_idx); cudaFree(c->send_buf); cudaFree(c->halo); cudaFree(c->zinv); cudaFree(c->zrecv); cudaFree(c->xfull);
  c->send_idx = nullptr; c->send_buf = nullptr; c->halo = nullptr; c->zinv = nullptr; c->zrecv = nullptr; c->xfull = nullptr;
  // One peer-readable allocation with offsets that are the same on every rank:
  //   [halo: cap | z0: n_pad + cap | z1: n_pad + cap],  cap = (world - 1) * n_pad >= any rank's halo
  const i64 cap = (i64)(c->world - 1) * c->n_pad;
  if (nhalo > cap) return set_err(c, GA_ERR_ARG, "halo larger than the rest of the vector");
  const size_t total = (size_t)(cap + 2 * (c->n_pad + cap)) * c->esize;
  GA_CUDA(c, cudaMalloc(&c->halo, total));
  GA_CUDA(c, cudaMemset(c->halo, 0, total));
  c->zp[0] = (char*)c->halo + (size_t)cap * c->esize;
  c->zp[1] = (char*)c->zp[0] + (size_t)(c->n_pad + cap) * c->esize;
  GA_CUDA(c, cudaMalloc(&c->send_idx, sizeof(int) * (size_t)std::max(nsend, 1)));
  GA_CUDA(c, cudaMalloc(&c->send_buf, c->esize * (size_t)std::max(nsend, 1)));
  GA_CUDA(c, cudaMalloc(&c->zrecv, c->esize * (size_t)std::max(nsend, 1)));
  if (nsend > 0) GA_CUDA(c, cudaMemcpy(c->send_idx, send_idx, sizeof(int) * (size_t)nsend, cudaMemcpyHostToDevice));
  // inverse of the send lists: own row i sits at position zinv[q][i] of the segment rank q holds for me
  {
    std::vector<int> inv((size_t)c->world * c->n_pad, -1);
    size_t off = 0;
    for (int q = 0; q < c->world; ++q) {
      for (int k = 0; k < send_counts[q]; ++k) {
        const int row = send_idx[off + k];
        if (row < 0 || row >= c->n) return set_err(c, GA_ERR_ARG, "send_idx outside the local rows");
        inv[(size_t)q * c->n_pad + row] = k;
      }
      off += (size_t)send_counts[q];
    }
    GA_CUDA(c, cudaMalloc(&c->zinv, sizeof(int) * inv.size()));
    GA_CUDA(c, cudaMemcpy(c->zinv, inv.data(), sizeof(int) * inv.size(), cudaMemcpyHostToDevice));
  }
  c->mt_local = mt_local;
  c->mt_pad = (std::max<i64>(mt_local, 1) + kRowPad - 1) / kRowPad * kRowPad;
  GA_CUDA(c, cudaMalloc(&c->xfull, (size_t)c->mt_pad * c->esize));
  GA_CUDA(c, cudaMemset(c->xfull, 0, (size_t)c->mt_pad * c->esize));
  c->xfull_len = c->mt_pad;
  c->normal_dist = true; c->normal_left = true;
  c->normal_m = mt_local; c->normal_n = c->n;   // local view: T_loc is mt_local x (own | halo)
  c->zpar = 0;
  c->op_kind = 2;
  if ((rc = spmv_autotune(c, c->A, c->V, c->xfull))) return rc;
  return spmv_autotune(c, c->At, c->xfull, c->zp[0]);
}

// ArpackSymmetricGeneralizedOp(A, S, B): mode 2, bmat = :G (see the header and k_general.cu)
int ga_set_generalized_csr(ga_ctx* c, const int64_t* a_rowptr, const int32_t* a_col, const void* a_val,
                           const int64_t* b_rowptr, const int32_t* b_col, const void* b_val, const double* cg_tol,
                           int cg_maxit) {
  if (!c || !a_rowptr || !a_col || !a_val || !b_rowptr || !b_col || !b_val || cg_maxit < 0) return GA_ERR_ARG;
  GA_CUDA(c, cudaSetDevice(c->device));
  clear_operator(c);
  if (c->world != 1) return set_err(c, GA_ERR_ARG, "multi-GPU contexts take ga_set_generalized_csr_dist (local columns of A and B + their common halo plan)");
  if (c->dtype == GA_F32) return set_err(c, GA_ERR_ARG, "generalised problem: Float32 vectors are not supported (mode 1 / normal operator only)");
  c->cg_tol[0] = cg_tol ? cg_tol[0] : 0.0;
  c->cg_tol[1] = 0.0;
  c->cg_maxit = cg_maxit;
  c->solve_count = c->solve_iters = 0;
  c->solve_max_relres = 0.0;
  int rc = general_upload(c, c->n, a_rowptr, a_col, a_val, b_rowptr, b_col, b_val);
  if (rc) return rc;
  c->op_kind = 3;
  return 0;
}

// The same problem with rows of A and B block-partitioned over the GPUs (world_size > 1).  A and B share one halo plan:
// the caller remaps the columns of both matrices against the UNION of their remote columns (ga_set_csr_dist's layout).
int ga_set_generalized_csr_dist(ga_ctx* c, const int64_t* a_rowptr, const int32_t* a_col_local, const void* a_val,
                                const int64_t* b_rowptr, const int32_t* b_col_local, const void* b_val, const double* cg_tol,
                                int cg_maxit, int32_t nhalo, const int32_t* recv_counts, const int32_t* send_counts,
                                const int32_t* send_idx, const int32_t* send_dest_off) {
  if (!c || !a_rowptr || !a_col_local || !a_val || !b_rowptr || !b_col_local || !b_val || cg_maxit < 0 || !recv_counts ||
      !send_counts || nhalo < 0)
    return GA_ERR_ARG;
  GA_CUDA(c, cudaSetDevice(c->device));
  if (c->world < 2) return set_err(c, GA_ERR_ARG, "single-GPU contexts take ga_set_generalized_csr");
  if (c->dtype == GA_F32) return set_err(c, GA_ERR_ARG, "generalised problem: Float32 vectors are not supported (mode 1 / normal operator only)");
  if (c->ipc_exported || c->devcomm) return set_err(c, GA_ERR_ARG, "the sharded operator cannot be replaced after ga_comm_ipc_export (peers hold the halo buffer's handle): create a new context");
  int rc = validate_halo_plan(c, "ga_set_generalized_csr_dist", recv_counts, send_counts, send_idx);
  if (rc) return rc;
  clear_operator(c);
  c->normal_dist = false;
  c->cg_tol[0] = cg_tol ? cg_tol[0] : 0.0;
  c->cg_tol[1] = 0.0;
  c->cg_maxit = cg_maxit;
  c->solve_count = c->solve_iters = 0;
  c->solve_max_relres = 0.0;
  if ((rc = general_upload(c, c->n_pad + nhalo, a_rowptr, a_col_local, a_val, b_rowptr, b_col_local, b_val))) return rc;
  if ((rc = install_halo_plan(c, nhalo, recv_counts, send_counts, send_idx, send_dest_off))) return rc;
  c->general_dist = true;
  c->op_kind = 3;
  return 0;
}

int ga_set_op_callback(ga_ctx* c, ga_op_fn opx, void* user) {
  if (!c || !opx) return GA_ERR_ARG;
  if (c->world != 1) return set_err(c, GA_ERR_ARG, "function operators: single GPU only");
  c->A.free_all(); c->At.free_all();
  c->op_cb = opx; c->av_cb = c->atv_cb = nullptr; c->op_user = user;
  c->op_kind = 1;
  return 0;
}

int ga_set_normal_op_callbacks(ga_ctx* c, int64_t m, int64_t n, ga_op_fn av, ga_op_fn atv, void* user) {
  if (!c || !av || !atv || m <= 0 || n <= 0) return GA_ERR_ARG;
  if (c->world != 1) return set_err(c, GA_ERR_ARG, "function operators: single GPU only");
  const bool left = !(m <= n);  // src/idonow_ops.jl:222-235
  if (c->n != (left ? n : m)) return set_err(c, GA_ERR_ARG, "normal operator: context size must be min(m, n) side");
  GA_CUDA(c, cudaSetDevice(c->device));
  clear_operator(c);
  c->A.free_all(); c->At.free_all();
  c->normal_m = m; c->normal_n = n; c->normal_left = left; c->normal_dist = false;
  const i64 extra = left ? m : n;
  const i64 extra_pad = (extra + row_pad(c) - 1) / row_pad(c) * row_pad(c);
  cudaFree(c->xfull);
  c->xfull = nullptr;
  GA_CUDA(c, cudaMalloc(&c->xfull, (size_t)extra_pad * c->esize));
  GA_CUDA(c, cudaMemset(c->xfull, 0, (size_t)extra_pad * c->esize));
  c->xfull_len = extra_pad;
  c->op_cb = nullptr; c->av_cb = av; c->atv_cb = atv; c->op_user = user;
  c->op_kind = 2;
  return 0;
}

int ga_set_solve_callback(ga_ctx* c, ga_solve_fn fn, void* user) {
  if (!c) return GA_ERR_ARG;
  c->solve_cb = fn;
  c->solve_user = user;
  return 0;
}

int ga_get_solve_stats(const ga_ctx* c, int64_t* solves, int64_t* iterations, double* max_relres) {
  if (!c) return GA_ERR_ARG;
  if (solves) *solves = c->solve_count;
  if (iterations) *iterations = c->solve_iters;
  if (max_relres) *max_relres = c->solve_max_relres;
  return 0;
}

int ga_bx(ga_ctx* c, const void* x_host, void* y_host) {
  if (!c || !x_host || !y_host) return GA_ERR_ARG;
  if (c->op_kind != 3) return set_err(c, GA_ERR_NOOP, "ga_bx: not a generalised-problem context (ga_set_generalized_csr)");
  GA_CUDA(c, cudaSetDevice(c->device));
  void *dx = nullptr, *dy = nullptr;
  GA_CUDA(c, cudaMalloc(&dx, (size_t)c->n_pad * c->esize));
  GA_CUDA(c, cudaMalloc(&dy, (size_t)c->n_pad * c->esize));
  GA_CUDA(c, cudaMemsetAsync(dx, 0, (size_t)c->n_pad * c->esize, c->stream));
  GA_CUDA(c, cudaMemsetAsync(dy, 0, (size_t)c->n_pad * c->esize, c->stream));
  GA_CUDA(c, cudaMemcpyAsync(dx, x_host, (size_t)c->n * c->esize, cudaMemcpyHostToDevice, c->stream));
  int rc = general_bx(c, dx, dy);
  // several GPUs: a reduction's mailbox exchange keeps the next halo push of any peer behind this product (k_general.cu)
  if (rc == 0 && c->world > 1) rc = launch_dotc(c, dy, dy);
  if (rc == 0) {
    cudaError_t e = cudaMemcpyAsync(y_host, dy, (size_t)c->n * c->esize, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) rc = set_err(c, GA_ERR_CUDA, cudaGetErrorString(e));
  }
  cudaFree(dx); cudaFree(dy);
  return rc;
}

int ga_opx(ga_ctx* c, const void* x_host, void* y_host) {
  if (!c || !x_host || !y_host) return GA_ERR_ARG;
  GA_CUDA(c, cudaSetDevice(c->device));
  void *dx = nullptr, *dy = nullptr;
  GA_CUDA(c, cudaMalloc(&dx, (size_t)c->n_pad * c->esize));
  GA_CUDA(c, cudaMalloc(&dy, (size_t)c->n_pad * c->esize));
  GA_CUDA(c, cudaMemsetAsync(dx, 0, (size_t)c->n_pad * c->esize, c->stream));
  GA_CUDA(c, cudaMemsetAsync(dy, 0, (size_t)c->n_pad * c->esize, c->stream));
  GA_CUDA(c, cudaMemcpyAsync(dx, x_host, (size_t)c->n * c->esize, cudaMemcpyHostToDevice, c->stream));
  int rc = launch_opx(c, dx, dy, false);
  if (rc == 0) {
    cudaError_t e = cudaMemcpyAsync(y_host, dy, (size_t)c->n * c->esize, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) rc = set_err(c, GA_ERR_CUDA, cudaGetErrorString(e));
  }
  cudaFree(dx); cudaFree(dy);
  return rc;
}

// ------------------------------------------------------------------ data movement
int ga_upload_resid(ga_ctx* c, const void* h) {
  if (!c || !h) return GA_ERR_ARG;
  GA_CUDA(c, cudaMemcpyAsync(c->resid, h, (size_t)c->n * c->esize, cudaMemcpyHostToDevice, c->stream));
  GA_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}
int ga_download_resid(ga_ctx* c, void* h) {
  if (!c || !h) return GA_ERR_ARG;
  GA_CUDA(c, cudaMemcpyAsync(h, c->resid, (size_t)c->n * c->esize, cudaMemcpyDeviceToHost, c->stream));
  GA_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}
int ga_upload_v(ga_ctx* c, int col0, int ncols, const void* h, int64_t ldv) {
  if (!c || !h || col0 < 0 || ncols < 0 || col0 + ncols > c->ncv || ldv < c->n) return GA_ERR_ARG;
  if (ncols == 0) return 0;
  GA_CUDA(c, cudaMemcpy2DAsync((char*)c->V + (size_t)col0 * c->n_pad * c->esize, (size_t)c->n_pad * c->esize, h,
                               (size_t)ldv * c->esize, (size_t)c->n * c->esize, ncols, cudaMemcpyHostToDevice,
                               c->stream));
  GA_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}
int ga_download_v(ga_ctx* c, int col0, int ncols, void* h, int64_t ldv) {
  if (!c || !h || col0 < 0 || ncols < 0 || col0 + ncols > c->ncv || ldv < c->n) return GA_ERR_ARG;
  if (ncols == 0) return 0;
  GA_CUDA(c, cudaMemcpy2DAsync(h, (size_t)ldv * c->esize, (char*)c->V + (size_t)col0 * c->n_pad * c->esize,
                               (size_t)c->n_pad * c->esize, (size_t)c->n * c->esize, ncols, cudaMemcpyDeviceToHost,
                               c->stream));
  GA_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}

// ------------------------------------------------------------------ dgetv0!
static int reset_status(ga_ctx* c) {
  static const int zeros[4] = {0, 0, 0, 0};  // status, phase, j_break, rstart_j
  GA_CUDA(c, cudaMemcpyAsync(c->ctl, zeros, sizeof(zeros), cudaMemcpyHostToDevice, c->stream));
  return 0;
}
static bool real_gt_717(const ga_ctx* c, const double a[2], const double b[2]) {  // a > 0.717 * b
  if (c->dtype == GA_DD) {
    ddbl x(a[0], a[1]), y(b[0], b[1]);
    return x > dd_mul_d(y, 0.717);
  }
  return a[0] > 0.717 * b[0];
}

int ga_getv0(ga_ctx* c, int64_t iseed[4], int initv, int j, double* rnorm) {
  if (!c || !iseed || !rnorm || j < 1 || j > c->ncv) return GA_ERR_ARG;
  GA_CUDA(c, cudaSetDevice(c->device));
  if (c->op_kind == 3) return general_getv0(c, iseed, initv, j, rnorm);   // bmat = :G
  const double t0 = now_s();
  int rc = reset_status(c);
  if (rc) return rc;
  int ierr = 0;
  if (!initv) {                                                // src/dgetv0.jl:542-553
    i64 seed[4] = {iseed[0], iseed[1], iseed[2], iseed[3]};
    rc = launch_lcg_fill_advance(c, seed);
    if (rc) return rc;
    for (int i = 0; i < 4; ++i) iseed[i] = seed[i];
  }
  if (j == 1) {                                                // :630-637
    rc = launch_norm_resid(c, DK_NORM);
    if (rc) return rc;
    rc = sync_ctl(c);
    if (rc) return rc;
    if ((rc = status_error(c, "ga_getv0"))) return rc;
    pair_to_host_real(c, c->ctl_host->rnorm, rnorm);
  } else {
    // s = V[:,1:j-1]' r ; r -= V s ; repeat while ||r|| <= 0.717 ||r_prev||, at most 5 refinements
    rc = launch_gs(c, 0, j - 1, -1, DK_GETV0_DOT);             // :630,662 (norm + first dots)
    if (rc) return rc;
    int iter = 0;
    double rnorm0[2], rn[2];
    rc = sync_ctl(c);
    if (rc) return rc;
    if ((rc = status_error(c, "ga_getv0"))) return rc;
    rnorm0[0] = c->ctl_host->rnorm0[0]; rnorm0[1] = c->ctl_host->rnorm0[1];
    while (true) {
      rc = launch_gs(c, 1, j - 1, -1, DK_GETV0_UPD);           // :665,700 (+ next dots)
      if (rc) return rc;
      rc = sync_ctl(c);
      if (rc) return rc;
      if ((rc = status_error(c, "ga_getv0"))) return rc;
      rn[0] = c->ctl_host->rnorm[0]; rn[1] = c->ctl_host->rnorm[1];
      if (real_gt_717(c, rn, rnorm0)) break;                   // :711
      iter += 1;
      if (iter <= 5) { rnorm0[0] = rn[0]; rnorm0[1] = rn[1]; } // :716-720
      else {                                                   // :725-729
        rc = launch_zero_resid(c);
        if (rc) return rc;
        rn[0] = rn[1] = 0.0;
        GA_CUDA(c, cudaMemcpyAsync(c->ctl->rnorm, rn, sizeof(rn), cudaMemcpyHostToDevice, c->stream));
        GA_CUDA(c, cudaStreamSynchronize(c->stream));
        ierr = -1;
        break;
      }
    }
    pair_to_host_real(c, rn, rnorm);
  }
  c->stats.tgetv0 += now_s() - t0;
  return ierr;
}

// ------------------------------------------------------------------ dsaitr!
static int enqueue_steps(ga_ctx* c, int jfirst, int jlast) {
  // One Lanczos step = 6 launches, no host synchronisation (1-based step j, column j-1)
  for (int j = jfirst; j <= jlast; ++j) {
    int rc;
    if ((rc = launch_normalize(c, j - 1))) return rc;                       // STEP 2 (:1097-1105)
    if ((rc = launch_opx(c, (const char*)c->V + (size_t)(j - 1) * c->n_pad * c->esize, c->resid, true))) return rc;  // STEP 3
    if (c->coop && c->use_step_kernel && (c->world == 1 || c->devcomm != nullptr)) {
      // all sweeps of the step in one persistent cooperative launch (k_gs_step.cu)
      if ((rc = launch_gs_step(c, j))) return rc;
      continue;
    }
    if ((rc = launch_gs(c, 0, j, -1, DK_DOT))) return rc;                   // wnorm, h = V'w (:1208,1227)
    if ((rc = launch_gs(c, 1, j, -1, DK_CGS))) return rc;                   // r = w - V h, s = V'r, ||r|| (:1240-1324)
    if ((rc = launch_gs(c, 1, j, PH_ORTH1, DK_ORTH1))) return rc;           // DGKS round 1 (:1352-1434)
    if ((rc = launch_gs(c, 2, j, PH_ORTH2, DK_ORTH2))) return rc;           // DGKS round 2 (:1424-1446)
  }
  return 0;
}

int ga_lanczos(ga_ctx* c, int k, int np, void* H, int ldh, double* rnorm, int64_t iseed[4]) {
  if (!c || !H || !rnorm || !iseed || k < 0 || np < 0 || k + np > c->ncv || ldh < k + np) return GA_ERR_ARG;
  if (c->op_kind == 0) return set_err(c, GA_ERR_NOOP, "no operator uploaded (ga_set_csr)");
  if (np == 0) return 0;
  GA_CUDA(c, cudaSetDevice(c->device));
  if (c->op_kind == 3) return general_lanczos(c, k, np, H, ldh, rnorm, iseed);   // mode 2, bmat = :G
  const double t0 = now_s();
  const size_t rs = real_size(c);
  // control block: status/phase cleared, counters cleared, rnorm from the caller
  DevCtl* hc = c->ctl_host;
  int rc = reset_status(c);
  if (rc) return rc;
  {
    int zeros[3] = {0, 0, 0};  // nrorth, nitref, zero_resid
    GA_CUDA(c, cudaMemcpyAsync(&c->ctl->nrorth, zeros, sizeof(zeros), cudaMemcpyHostToDevice, c->stream));
    double rn[2];
    host_real_to_pair(c, rnorm, rn);
    GA_CUDA(c, cudaMemcpyAsync(c->ctl->rnorm, rn, sizeof(rn), cudaMemcpyHostToDevice, c->stream));
  }
  int info = 0;
  int jfirst = k + 1;
  const int jlast = k + np;
  while (jfirst <= jlast) {
    if ((rc = enqueue_steps(c, jfirst, jlast))) return rc;
    if ((rc = sync_ctl(c))) return rc;
    c->stats.nrorth += hc->nrorth;
    c->stats.nitref += hc->nitref;
    if (hc->status == ST_OK) {
      c->stats.nopx += jlast - jfirst + 1;
      break;
    }
    if ((rc = status_error(c, "ga_lanczos"))) return rc;
    if (hc->status != ST_BREAKDOWN) return set_err(c, GA_ERR_CUDA, "ga_lanczos: unknown device status " + std::to_string(hc->status));
    // ---- invariant subspace at step jb: restart with dgetv0 (src/dsaitr.jl:1061-1087,1497-1525)
    const int jb = hc->j_break;
    c->stats.nopx += jb - jfirst;
    c->stats.nrstrt += 1;
    int itry = 1, ierr;
    double rn[2] = {0.0, 0.0};
    while (true) {
      ierr = ga_getv0(c, iseed, 0, jb, rn);
      if (ierr < GA_ERR_CUDA + 1 && ierr != -1) return ierr;  // library error
      if (ierr < 0) {
        itry += 1;
        if (itry <= 3) continue;
        info = jb - 1;                                         // :1520
        break;
      }
      break;
    }
    if (info > 0) { jfirst = jb; break; }
    // resume at step jb with rstart = true (H[jb,1] = 0), counters restarted
    int st[4] = {ST_OK, PH_DONE, 0, jb};
    GA_CUDA(c, cudaMemcpyAsync(c->ctl, st, sizeof(st), cudaMemcpyHostToDevice, c->stream));
    int zeros[3] = {0, 0, 0};
    GA_CUDA(c, cudaMemcpyAsync(&c->ctl->nrorth, zeros, sizeof(zeros), cudaMemcpyHostToDevice, c->stream));
    jfirst = jb;
  }
  // results: H rows k..k+np-1 (both columns), rnorm
  if ((rc = sync_ctl(c))) return rc;
  if ((rc = status_error(c, "ga_lanczos"))) return rc;
  if (hc->zero_resid) {                                         // :1442-1443 on the last step
    if ((rc = launch_zero_resid(c))) return rc;
    GA_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  const int jdone = info > 0 ? info : jlast;
  const char* Hd = (const char*)hc->H;
  for (int col = 0; col < 2; ++col)
    for (int r = k; r < jdone; ++r)
      memcpy((char*)H + ((size_t)col * ldh + r) * rs, Hd + ((size_t)col * c->ncv + r) * rs, rs);
  pair_to_host_real(c, hc->rnorm, rnorm);
  c->stats.taitr += now_s() - t0;
  return info;
}

// ------------------------------------------------------------------ dsapps! tail
int ga_apply_shifts(ga_ctx* c, int kev, int np, const void* Q, int ldq, const double* beta, double* rnorm_out) {
  if (!c || !Q || !beta || !rnorm_out || kev < 1 || np < 0 || kev + np > c->ncv || ldq < kev + np) return GA_ERR_ARG;
  GA_CUDA(c, cudaSetDevice(c->device));
  const double t0 = now_s();
  const int kplusp = kev + np;
  const size_t rs = real_size(c);
  int rc = reset_status(c);
  if (rc) return rc;
  // Q columns 1..kev+1 -> device (ld = ncv_pad)
  const int ncolq = kev + 1 <= kplusp ? kev + 1 : kplusp;
  GA_CUDA(c, cudaMemcpy2DAsync(c->qdev, (size_t)c->ncv_pad * rs, Q, (size_t)ldq * rs, (size_t)kplusp * rs, ncolq,
                               cudaMemcpyHostToDevice, c->stream));
  const bool bpos = c->dtype == GA_DD ? (beta[0] > 0.0 || (beta[0] == 0.0 && beta[1] > 0.0)) : beta[0] > 0.0;
  const int nout = (bpos && kev + 1 <= kplusp) ? kev + 1 : kev;       // src/dsapps.jl:813,839
  rc = launch_vq(c, kplusp, nout, true, nullptr, beta, kev);
  if (rc) return rc;
  if ((rc = sync_ctl(c))) return rc;
  if ((rc = status_error(c, "ga_apply_shifts"))) return rc;
  pair_to_host_real(c, c->ctl_host->rnorm, rnorm_out);
  if (c->op_kind == 3 && (rc = general_resid_bnorm(c, rnorm_out))) return rc;   // bmat = :G: the B-norm (src/dsaup2.jl:1383-1411)
  c->stats.tapps += now_s() - t0;
  return 0;
}

// ------------------------------------------------------------------ dseupd! tail
int ga_ritz_vectors(ga_ctx* c, int nconv, const void* Qh, int ldq, void* z_host, int64_t ldz) {
  if (!c || !Qh || nconv < 0 || nconv > c->ncv || ldq < c->ncv) return GA_ERR_ARG;
  if (z_host && ldz < c->n) return GA_ERR_ARG;
  GA_CUDA(c, cudaSetDevice(c->device));
  const double t0 = now_s();
  const size_t rs = real_size(c);
  int rc = reset_status(c);
  if (rc) return rc;
  GA_CUDA(c, cudaMemcpy2DAsync(c->qdev, (size_t)c->ncv_pad * rs, Qh, (size_t)ldq * rs, (size_t)c->ncv * rs, c->ncv,
                               cudaMemcpyHostToDevice, c->stream));
  rc = launch_vq(c, c->ncv, c->ncv, false, nullptr, nullptr, 0);      // V <- V Qh (src/dseupd.jl:1431)
  if (rc) return rc;
  GA_CUDA(c, cudaStreamSynchronize(c->stream));
  if (z_host && nconv > 0) rc = ga_download_v(c, 0, nconv, z_host, ldz);  // Z <- V[:,1:nconv] (:1434)
  c->stats.trvec += now_s() - t0;
  return rc;
}

// ------------------------------------------------------------------ svds: eigenvecs_to_singvecs! + orth!
int ga_svd_vectors(ga_ctx* c, int nvec, const int32_t* perm, void* w_host, int64_t ldw, double* norms_out) {
  if (!c || !perm || !norms_out || nvec < 1 || nvec > c->ncv) return GA_ERR_ARG;
  if (c->op_kind != 2) return set_err(c, GA_ERR_NOOP, "ga_svd_vectors needs a normal-operator context");
  for (int i = 0; i < nvec; ++i)
    if (perm[i] < 0 || perm[i] >= c->ncv) return GA_ERR_ARG;
  const i64 other = c->normal_left ? c->normal_m : c->normal_n;   // rows of the other side's vectors
  if (w_host && ldw < other) return GA_ERR_ARG;
  GA_CUDA(c, cudaSetDevice(c->device));
  const double t0 = now_s();
  const size_t rs = real_size(c);
  const i64 ld = (other + row_pad(c) - 1) / row_pad(c) * row_pad(c);
  if (!c->wbuf || c->wbuf_ld != ld || c->wbuf_cols < nvec) {
    cudaFree(c->wbuf);
    c->wbuf = nullptr;
    GA_CUDA(c, cudaMalloc(&c->wbuf, (size_t)ld * nvec * c->esize));
    c->wbuf_ld = ld; c->wbuf_cols = nvec;
  }
  GA_CUDA(c, cudaMemsetAsync(c->wbuf, 0, (size_t)ld * nvec * c->esize, c->stream));  // padding rows stay zero
  void* norms_dev = nullptr;
  GA_CUDA(c, cudaMalloc(&norms_dev, 16 * (size_t)nvec));
  int rc = reset_status(c);
  const ga_ctx::VSpace saved = c->vs;
  for (int i = 0; i < nvec && rc == 0; ++i) {
    char* wi = (char*)c->wbuf + (size_t)i * ld * c->esize;
    rc = launch_normal_half(c, (const char*)c->V + (size_t)perm[i] * c->n_pad * c->esize, wi);   // :311 / :322
    if (rc) break;
    c->vs.V = c->wbuf; c->vs.w = wi; c->vs.ld = ld; c->vs.rows = other;
    if ((rc = launch_norm_resid(c, DK_NORM))) break;                       // nrm = norm(U[:,i])  (:313)
    if (cudaMemcpyAsync((char*)norms_dev + 16 * (size_t)i, c->ctl->rnorm, 16, cudaMemcpyDeviceToDevice, c->stream) != cudaSuccess) { rc = GA_ERR_CUDA; break; }
    if ((rc = launch_normalize(c, i, false))) break;                      // _scale_from_to(nrm, 1, U[:,i])  (:315)
    if (i == 0) continue;                                                 // orth!: first column is already unit
    if ((rc = launch_gs(c, 0, i, -1, DK_GETV0_DOT))) break;               // h = W_i' w
    if ((rc = launch_gs(c, 1, i, -1, DK_GETV0_UPD))) break;               // w -= W_i h ; s = W_i' w
    if ((rc = launch_gs(c, 2, i, -1, DK_NORM))) break;                    // w -= W_i s ; ||w||
    if ((rc = launch_normalize(c, i, false))) break;                      // unit column, R[i,i] > 0
  }
  c->vs = saved;
  if (rc == 0) rc = sync_ctl(c);
  if (rc == 0) rc = status_error(c, "ga_svd_vectors");
  if (rc == 0 && c->ctl_host->status != ST_OK)
    rc = set_err(c, GA_ERR_NUMERIC, "svd vectors: a column of A*V vanished or is not finite");
  if (rc == 0) {
    std::vector<double> nb(2 * (size_t)nvec);
    cudaError_t e = cudaMemcpy(nb.data(), norms_dev, 16 * (size_t)nvec, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = set_err(c, GA_ERR_CUDA, cudaGetErrorString(e));
    for (int i = 0; i < nvec && rc == 0; ++i) memcpy((char*)norms_out + rs * (size_t)i, &nb[2 * (size_t)i], rs);
  }
  cudaFree(norms_dev);
  if (rc == 0 && w_host) {
    cudaError_t e = cudaMemcpy2DAsync(w_host, (size_t)ldw * c->esize, c->wbuf, (size_t)ld * c->esize,
                                      (size_t)other * c->esize, nvec, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) rc = set_err(c, GA_ERR_CUDA, cudaGetErrorString(e));
  }
  c->stats.trvec += now_s() - t0;
  return rc;
}

int ga_norm2_resid(ga_ctx* c, double* out) {
  if (!c || !out) return GA_ERR_ARG;
  GA_CUDA(c, cudaSetDevice(c->device));
  int rc = reset_status(c);
  if (rc) return rc;
  if ((rc = launch_norm_resid(c, DK_NORM))) return rc;
  if ((rc = sync_ctl(c))) return rc;
  if ((rc = status_error(c, "ga_norm2_resid"))) return rc;
  pair_to_host_real(c, c->ctl_host->rnorm, out);
  return 0;
}

int ga_bnorm_resid(ga_ctx* c, double* out) {
  if (!c || !out) return GA_ERR_ARG;
  if (c->op_kind != 3) return ga_norm2_resid(c, out);        // bmat = :I: the B-norm is the 2-norm
  GA_CUDA(c, cudaSetDevice(c->device));
  return general_resid_bnorm(c, out);
}

int ga_timer_start(ga_ctx* c) {
  if (!c) return GA_ERR_ARG;
  GA_CUDA(c, cudaSetDevice(c->device));
  if (!c->tm0) { GA_CUDA(c, cudaEventCreate(&c->tm0)); GA_CUDA(c, cudaEventCreate(&c->tm1)); }
  GA_CUDA(c, cudaStreamSynchronize(c->stream));
  GA_CUDA(c, cudaEventRecord(c->tm0, c->stream));
  return 0;
}
int ga_timer_stop(ga_ctx* c, double* ms) {
  if (!c || !ms || !c->tm0) return GA_ERR_ARG;
  GA_CUDA(c, cudaEventRecord(c->tm1, c->stream));
  GA_CUDA(c, cudaEventSynchronize(c->tm1));
  float f = 0.f;
  GA_CUDA(c, cudaEventElapsedTime(&f, c->tm0, c->tm1));
  *ms = (double)f;
  return 0;
}
int ga_profile(ga_ctx* c, int enable) {
  if (!c) return GA_ERR_ARG;
  GA_CUDA(c, cudaSetDevice(c->device));
  GA_CUDA(c, cudaStreamSynchronize(c->stream));
  c->prof = enable != 0;
  c->prof_used = 0;
  c->launches = 0;
  GA_CUDA(c, cudaMemsetAsync(&c->ctl->prof_runs, 0, sizeof(unsigned int), c->stream));
  GA_CUDA(c, cudaMemsetAsync(&c->ctl->prof_units, 0, sizeof(unsigned long long), c->stream));
  GA_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}
int ga_profile_read(ga_ctx* c, double* gs_ms, int64_t* gs_runs, int64_t* gs_units, int64_t* launches) {
  if (!c) return GA_ERR_ARG;
  int rc = sync_ctl(c);
  if (rc) return rc;
  double tot = 0.0;
  for (size_t i = 0; i + 1 < c->prof_used; i += 2) {
    float f = 0.f;
    GA_CUDA(c, cudaEventElapsedTime(&f, c->prof_ev[i], c->prof_ev[i + 1]));
    tot += (double)f;
  }
  if (gs_ms) *gs_ms = tot;
  if (gs_runs) *gs_runs = (int64_t)c->ctl_host->prof_runs;
  if (gs_units) *gs_units = (int64_t)c->ctl_host->prof_units;
  if (launches) *launches = (int64_t)c->launches;
  return 0;
}

int ga_bench_opx(ga_ctx* c, int reps, double* ms) {
  if (!c || !ms || reps < 1) return GA_ERR_ARG;
  if (c->op_kind == 0) return set_err(c, GA_ERR_NOOP, "no operator uploaded (ga_set_csr)");
  GA_CUDA(c, cudaSetDevice(c->device));
  int rc = launch_opx(c, c->V, c->resid, false);  // warm-up
  if (rc) return rc;
  if ((rc = ga_timer_start(c))) return rc;
  for (int i = 0; i < reps; ++i)
    if ((rc = launch_opx(c, c->V, c->resid, false))) return rc;
  double t = 0.0;
  if ((rc = ga_timer_stop(c, &t))) return rc;
  *ms = t / reps;
  return 0;
}

int ga_profile_gs_times(ga_ctx* c, float* ms_out, int64_t n) {
  if (!c || !ms_out || n < 0) return GA_ERR_ARG;
  GA_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int64_t i = 0; i < n && 2 * (size_t)i + 1 < c->prof_used; ++i)
    GA_CUDA(c, cudaEventElapsedTime(&ms_out[i], c->prof_ev[2 * (size_t)i], c->prof_ev[2 * (size_t)i + 1]));
  return 0;
}

int ga_profile_gs_launches(ga_ctx* c, int64_t* n) {
  if (!c || !n) return GA_ERR_ARG;
  *n = (int64_t)(c->prof_used / 2);
  return 0;
}

// debug: enable (out == NULL) or read back (out = 16 x uint64 ns) the in-kernel timeline of the
// most recent Gram-Schmidt sweep launch
int ga_debug_trace(ga_ctx* c, uint64_t* out16) {  // 64 samples
  if (!c) return GA_ERR_ARG;
  GA_CUDA(c, cudaSetDevice(c->device));
  if (!c->trace) {
    GA_CUDA(c, cudaMalloc(&c->trace, 64 * sizeof(unsigned long long)));
    GA_CUDA(c, cudaMemset(c->trace, 0, 64 * sizeof(unsigned long long)));
  }
  if (out16) {
    GA_CUDA(c, cudaStreamSynchronize(c->stream));
    GA_CUDA(c, cudaMemcpy(out16, c->trace, 64 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  }
  return 0;
}

// Override the upload-time autotune of the streamed SpMV for this context's operator: groups in {4, 6},
// groups <= stages <= 8; groups == 0 selects the first (non-streamed) row-segment kernel.
int ga_tune_spmv(ga_ctx* c, int groups, int stages) {
  if (!c) return GA_ERR_ARG;
  if (groups == 0) { c->use_spmv_stream = false; return 0; }
  if ((groups != 4 && groups != 6) || stages < groups || stages > 8) return set_err(c, GA_ERR_ARG, "ga_tune_spmv: groups in {0,4,6}, groups <= stages <= 8");
  c->use_spmv_stream = true;
  c->A.tune_groups = c->At.tune_groups = groups;
  c->A.tune_stages = c->At.tune_stages = stages;
  return 0;
}

// debug: enable (out == NULL) or read back (out = 128 x uint32) the streamed SpMV's stage self-check:
// out[0] = number of row blocks whose staged rowptr slice did not belong to the block, then 15 records
// of 8 words from out[8] on (CTA, ring iteration, block-in-CTA, expected k0, found, expected k1, found, 16*stages+groups)
int ga_debug_spmv_check(ga_ctx* c, uint32_t* out128) {
  if (!c) return GA_ERR_ARG;
  GA_CUDA(c, cudaSetDevice(c->device));
  if (!c->spmv_chk) {
    GA_CUDA(c, cudaMalloc(&c->spmv_chk, 128 * sizeof(unsigned int)));
    GA_CUDA(c, cudaMemset(c->spmv_chk, 0, 128 * sizeof(unsigned int)));
  }
  if (out128) {
    GA_CUDA(c, cudaStreamSynchronize(c->stream));
    GA_CUDA(c, cudaMemcpy(out128, c->spmv_chk, 128 * sizeof(unsigned int), cudaMemcpyDeviceToHost));
  }
  return 0;
}

int ga_get_stats(const ga_ctx* c, ga_stats* out) {
  if (!c || !out) return GA_ERR_ARG;
  *out = c->stats;
  return 0;
}
int ga_reset_stats(ga_ctx* c) {
  if (!c) return GA_ERR_ARG;
  c->stats = ga_stats{};
  return 0;
}

}  // extern "C"
