// Collective layer: one process per GPU, NCCL over NVLink 5 / NVSwitch.
//
// The Lanczos data distribution is P_ARPACK's (src/dsaup2.jl:911-926): rows of A, V, resid are
// block-partitioned, H/Q/ritz are replicated.  The exchanges the path needs are tiny:
//   - per Gram-Schmidt sweep: the (j+3) 16-byte reduction slots of every rank are ALL-GATHERED and
//     summed in rank order by every rank (bit-identical results on all ranks, and a correct
//     double-double sum, which a NCCL sum of hi/lo words would not be);
//   - per OP*x: only the HALO of the operand vector moves: each rank packs the entries its peers'
//     rows reference and one grouped ncclSend/ncclRecv round delivers them (32 KB per neighbour for
//     the 2-D Laplacian at n = 16M); the plan comes from the caller (ga_set_csr_dist).
// libnccl is opened at run time (the one torch already loaded, if any) so the library has no
// link-time dependency and a single-GPU user never touches NCCL.
#include <dlfcn.h>
#include <nccl.h>

#include "ga_ctx.cuh"

namespace ga {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

static NcclApi& nccl() {
  static NcclApi api;
  if (api.lib) return api;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) break;
  }
  if (!api.lib) return api;
  api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.lib, "ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.lib, "ncclCommInitRank");
  api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.lib, "ncclCommDestroy");
  api.AllGather = (decltype(api.AllGather))dlsym(api.lib, "ncclAllGather");
  api.Send = (decltype(api.Send))dlsym(api.lib, "ncclSend");
  api.Recv = (decltype(api.Recv))dlsym(api.lib, "ncclRecv");
  api.GroupStart = (decltype(api.GroupStart))dlsym(api.lib, "ncclGroupStart");
  api.GroupEnd = (decltype(api.GroupEnd))dlsym(api.lib, "ncclGroupEnd");
  api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.lib, "ncclGetErrorString");
  api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.Send && api.Recv &&
           api.GroupStart && api.GroupEnd && api.GetErrorString;
  return api;
}

#define GA_NCCL(ctx, call)                                                                          \
  do {                                                                                              \
    ncclResult_t r_ = (call);                                                                       \
    if (r_ != ncclSuccess)                                                                          \
      return set_err((ctx), GA_ERR_COMM, std::string(#call) + ": " + nccl().GetErrorString(r_));    \
  } while (0)

int comm_unique_id(void* id128) {
  NcclApi& a = nccl();
  if (!a.ok) return GA_ERR_COMM;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  return a.GetUniqueId((ncclUniqueId*)id128) == ncclSuccess ? 0 : GA_ERR_COMM;
}

int comm_init(ga_ctx* c, const void* id128) {
  NcclApi& a = nccl();
  if (!a.ok) return set_err(c, GA_ERR_COMM, "libnccl.so.2 could not be loaded");
  if (c->world <= 1) return 0;
  // The communicator is created on first use (ncclCommInitRank costs ~1 s and the fused
  // peer-memory path never needs it on the per-step path); creation is collective, and so is every
  // code path that reaches it.
  c->nccl_id.assign((const char*)id128, (const char*)id128 + sizeof(ncclUniqueId));
  return 0;
}

static int comm_ensure(ga_ctx* c) {
  if (c->nccl_comm) return 0;
  if (c->nccl_id.size() != sizeof(ncclUniqueId)) return set_err(c, GA_ERR_COMM, "ga_comm_init was not called");
  NcclApi& a = nccl();
  ncclUniqueId id;
  memcpy(&id, c->nccl_id.data(), sizeof(id));
  ncclComm_t comm;
  GA_CUDA(c, cudaSetDevice(c->device));
  GA_NCCL(c, a.CommInitRank(&comm, c->world, id, c->rank));
  c->nccl_comm = (void*)comm;
  return 0;
}

// ---- peer-memory (cudaIpc) setup for the fused collectives --------------------------------------
// export: 2 x 64-byte handles (mailbox, halo buffer); import: world x 128 bytes from every rank.
int comm_ipc_export(ga_ctx* c, void* out128) {
  GA_CUDA(c, cudaSetDevice(c->device));
  if (!c->mbox) {
    GA_CUDA(c, cudaMalloc(&c->mbox, sizeof(Mailbox)));
    GA_CUDA(c, cudaMemset(c->mbox, 0, sizeof(Mailbox)));
  }
  // The halo buffer is allocated by ga_set_csr_dist / ga_set_normal_csr_dist (at least one element even without
  // remote columns): the operator comes first, so that the exported allocation is the one the peers will push into.
  if (!c->halo) return set_err(c, GA_ERR_ARG, "ga_comm_ipc_export: set the sharded operator first (ga_set_csr_dist / ga_set_normal_csr_dist)");
  c->ipc_exported = true;
  cudaIpcMemHandle_t h[2];
  GA_CUDA(c, cudaIpcGetMemHandle(&h[0], c->mbox));
  GA_CUDA(c, cudaIpcGetMemHandle(&h[1], c->halo));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(out128, h, 128);
  return 0;
}

int comm_ipc_import(ga_ctx* c, const void* all_handles) {
  if (c->world <= 1) return 0;
  if (c->world > kMaxWorld) return set_err(c, GA_ERR_ARG, "fused collectives support at most 8 GPUs (one node)");
  if (!c->mbox) return set_err(c, GA_ERR_ARG, "call ga_comm_ipc_export first");
  GA_CUDA(c, cudaSetDevice(c->device));
  DevComm& hc = c->hostcomm;
  memset(&hc, 0, sizeof(hc));
  hc.rank = c->rank; hc.world = c->world;
  for (int q = 0; q < c->world; ++q) {
    if (q == c->rank) { hc.peer[q] = c->mbox; hc.peer_halo[q] = c->halo; continue; }
    cudaIpcMemHandle_t h[2];
    memcpy(h, (const char*)all_handles + (size_t)q * 128, 128);
    void *pm = nullptr, *ph = nullptr;
    GA_CUDA(c, cudaIpcOpenMemHandle(&pm, h[0], cudaIpcMemLazyEnablePeerAccess));
    GA_CUDA(c, cudaIpcOpenMemHandle(&ph, h[1], cudaIpcMemLazyEnablePeerAccess));
    c->ipc_mapped[2 * q] = pm; c->ipc_mapped[2 * q + 1] = ph;
    hc.peer[q] = (Mailbox*)pm; hc.peer_halo[q] = ph;
  }
  int off = 0;
  for (int q = 0; q < c->world; ++q) {
    hc.send_off[q] = off;
    off += c->send_counts.empty() ? 0 : c->send_counts[q];
    hc.dest_off[q] = c->dest_off.empty() ? 0 : c->dest_off[q];
    hc.recv_counts[q] = c->recv_counts.empty() ? 0 : c->recv_counts[q];
  }
  hc.send_off[c->world] = off;
  if (c->normal_dist) {
    hc.zp_off[0] = (long long)((const char*)c->zp[0] - (const char*)c->halo);
    hc.zp_off[1] = (long long)((const char*)c->zp[1] - (const char*)c->halo);
    hc.zp_halo_elem = c->n_pad;
  }
  if (!c->devcomm) GA_CUDA(c, cudaMalloc(&c->devcomm, sizeof(DevComm)));
  GA_CUDA(c, cudaMemcpy(c->devcomm, &hc, sizeof(DevComm), cudaMemcpyHostToDevice));
  return 0;
}

void comm_destroy(ga_ctx* c) {
  for (int i = 0; i < 2 * kMaxWorld; ++i)
    if (c->ipc_mapped[i]) { cudaIpcCloseMemHandle(c->ipc_mapped[i]); c->ipc_mapped[i] = nullptr; }
  if (c->devcomm) { cudaFree(c->devcomm); c->devcomm = nullptr; }
  if (c->mbox) { cudaFree(c->mbox); c->mbox = nullptr; }
  if (c->nccl_comm) {
    nccl().CommDestroy((ncclComm_t)c->nccl_comm);
    c->nccl_comm = nullptr;
  }
}

// gather[rank][0..kMaxSlots) <- every rank's ctl->red
int comm_allgather_slots(ga_ctx* c, int nslots) {
  (void)nslots;
  if (c->world <= 1) return 0;
  if (int rc = comm_ensure(c)) return rc;
  GA_NCCL(c, nccl().AllGather(c->ctl->red, c->gather, (size_t)kMaxSlots * 2, ncclFloat64, (ncclComm_t)c->nccl_comm,
                              c->stream));
  return 0;
}

// pack kernel: sendbuf[i] = x[send_idx[i]]
template <class T>
__global__ void k_pack(const T* __restrict__ x, const int* __restrict__ idx, T* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = x[idx[i]];
}

// Halo exchange of the row-sharded SpMV: every rank packs the entries of its local x that its
// peers' rows reference, then one grouped ncclSend/ncclRecv round moves them over NVLink.
int comm_halo_exchange(ga_ctx* c, const void* x_local) {
  if (c->world <= 1) return 0;
  if (int rc = comm_ensure(c)) return rc;
  NcclApi& a = nccl();
  if (c->nsend > 0) {
    const int threads = 256, blocks = (c->nsend + threads - 1) / threads;
    if (c->esize == 4) k_pack<float><<<blocks, threads, 0, c->stream>>>((const float*)x_local, c->send_idx, (float*)c->send_buf, c->nsend);
    else if (c->esize == 8) k_pack<double><<<blocks, threads, 0, c->stream>>>((const double*)x_local, c->send_idx, (double*)c->send_buf, c->nsend);
    else k_pack<double2><<<blocks, threads, 0, c->stream>>>((const double2*)x_local, c->send_idx, (double2*)c->send_buf, c->nsend);
    GA_CUDA(c, cudaGetLastError());
    c->launches += 1;
  }
  const size_t per = c->esize == 4 ? 1 : c->esize / 8;                           // wire words per element
  const ncclDataType_t wire = c->esize == 4 ? ncclFloat32 : ncclFloat64;
  GA_NCCL(c, a.GroupStart());
  size_t soff = 0, roff = 0;
  for (int q = 0; q < c->world; ++q) {
    if (c->send_counts[q] > 0)
      GA_NCCL(c, a.Send((const char*)c->send_buf + soff * c->esize, (size_t)c->send_counts[q] * per, wire, q,
                        (ncclComm_t)c->nccl_comm, c->stream));
    if (c->recv_counts[q] > 0)
      GA_NCCL(c, a.Recv((char*)c->halo + roff * c->esize, (size_t)c->recv_counts[q] * per, wire, q,
                        (ncclComm_t)c->nccl_comm, c->stream));
    soff += (size_t)c->send_counts[q];
    roff += (size_t)c->recv_counts[q];
  }
  GA_NCCL(c, a.GroupEnd());
  return 0;
}

// NCCL path of the sharded normal operator's reduce-scatter: the segment of my partial result that
// belongs to rank q (my halo columns owned by q) goes to q; q's segment for me lands in zrecv laid
// out like send_buf (grouped by source rank, send_counts[q] entries each).
int comm_zp_exchange(ga_ctx* c, const void* zp) {
  if (c->world <= 1) return 0;
  if (int rc = comm_ensure(c)) return rc;
  NcclApi& a = nccl();
  const size_t per = c->esize == 4 ? 1 : c->esize / 8;                           // wire words per element
  const ncclDataType_t wire = c->esize == 4 ? ncclFloat32 : ncclFloat64;
  GA_NCCL(c, a.GroupStart());
  size_t soff = 0, roff = 0;
  for (int q = 0; q < c->world; ++q) {
    if (c->recv_counts[q] > 0)
      GA_NCCL(c, a.Send((const char*)zp + ((size_t)c->n_pad + roff) * c->esize, (size_t)c->recv_counts[q] * per, wire, q,
                        (ncclComm_t)c->nccl_comm, c->stream));
    if (c->send_counts[q] > 0)
      GA_NCCL(c, a.Recv((char*)c->zrecv + soff * c->esize, (size_t)c->send_counts[q] * per, wire, q,
                        (ncclComm_t)c->nccl_comm, c->stream));
    soff += (size_t)c->send_counts[q];
    roff += (size_t)c->recv_counts[q];
  }
  GA_NCCL(c, a.GroupEnd());
  return 0;
}

}  // namespace ga
