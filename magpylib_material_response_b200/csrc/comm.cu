// NCCL plumbing for the multi-GPU paths (one process per GPU).  NCCL is dlopen'ed at first use
// so that (a) the single-GPU path has no link-time NCCL dependency and (b) inside a torch
// process we bind to the libnccl.so.2 torch already loaded (same SONAME) instead of a second copy.
#include <dlfcn.h>

#include "common.cuh"
#include "comm.h"

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct {
  char internal[128];
} ncclUniqueId;
typedef int ncclResult_t;
// ncclDataType_t / ncclRedOp_t values (nccl.h): ncclChar/ncclInt8 = 0, ncclFloat64 = 8; ncclSum = 0
constexpr int kNcclChar = 0;
constexpr int kNcclFloat64 = 8;

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_nccl;

int load_nccl() {
  if (g_nccl.lib) return 0;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* lib = nullptr;
  for (const char* nm : names) {
    lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (lib) break;
  }
  if (!lib) {
    mmr_set_error("cannot dlopen libnccl.so.2: %s", dlerror());
    return 2000 + 255;
  }
#define LOAD(field, sym)                                          \
  g_nccl.field = (decltype(g_nccl.field))dlsym(lib, sym);         \
  if (!g_nccl.field) {                                            \
    mmr_set_error("NCCL symbol %s not found", sym);               \
    return 2000 + 254;                                            \
  }
  LOAD(GetUniqueId, "ncclGetUniqueId");
  LOAD(CommInitRank, "ncclCommInitRank");
  LOAD(CommDestroy, "ncclCommDestroy");
  LOAD(AllGather, "ncclAllGather");
  LOAD(Broadcast, "ncclBroadcast");
  LOAD(GetErrorString, "ncclGetErrorString");
#undef LOAD
  g_nccl.lib = lib;
  return 0;
}

int fail_nccl(ncclResult_t r, const char* what) {
  mmr_set_error("NCCL error %d (%s) in %s", (int)r, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?",
                what);
  return 2000 + (int)r;
}

}  // namespace

extern "C" int mmr_nccl_unique_id(void* id128_out) {
  MMR_ARG(id128_out != nullptr, "id buffer is NULL");
  int rc = load_nccl();
  if (rc) return rc;
  ncclUniqueId id;
  ncclResult_t r = g_nccl.GetUniqueId(&id);
  if (r != 0) return fail_nccl(r, "ncclGetUniqueId");
  memcpy(id128_out, id.internal, 128);
  return 0;
}

extern "C" int mmr_comm_init(mmr_ctx* ctx, const void* id128, int nranks, int rank) {
  MMR_ARG(ctx != nullptr && id128 != nullptr, "NULL ctx/id");
  MMR_ARG(nranks >= 1 && rank >= 0 && rank < nranks, "bad nranks/rank");
  MMR_ARG(ctx->nccl_comm == nullptr, "communicator already initialised");
  int rc = load_nccl();
  if (rc) return rc;
  DeviceGuard g(ctx->device);
  ncclUniqueId id;
  memcpy(id.internal, id128, 128);
  ncclComm_t comm = nullptr;
  ncclResult_t r = g_nccl.CommInitRank(&comm, nranks, id, rank);
  if (r != 0) return fail_nccl(r, "ncclCommInitRank");
  ctx->nccl_comm = comm;
  ctx->nranks = nranks;
  ctx->rank = rank;
  return 0;
}

// ---- peer-memory buffers (CUDA IPC) for the flag-synchronised triangular sweeps ----------------------------
// Every rank allocates one buffer, exports its 64-byte IPC handle (mmr_comm_p2p_export), the host layer gathers
// the handles of all ranks (torch.distributed is plumbing only) and hands the table to mmr_comm_p2p_open, which
// maps every peer's buffer into this process.  All GPUs of one NVSwitch box reach each other directly.
extern "C" int mmr_comm_p2p_export(mmr_ctx* ctx, void* handle64_out) {
  MMR_ARG(ctx != nullptr && handle64_out != nullptr, "NULL ctx/handle");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  DeviceGuard g(ctx->device);
  if (!ctx->p2p_local) {
    MMR_CUDA(cudaMalloc(&ctx->p2p_local, mmr_p2p_bytes()));
    MMR_CUDA(cudaMemset(ctx->p2p_local, 0, mmr_p2p_bytes()));
  }
  cudaIpcMemHandle_t h;
  MMR_CUDA(cudaIpcGetMemHandle(&h, ctx->p2p_local));
  memcpy(handle64_out, &h, 64);
  return 0;
}

extern "C" int mmr_comm_p2p_open(mmr_ctx* ctx, const void* handles, int nranks, int rank) {
  MMR_ARG(ctx != nullptr && handles != nullptr, "NULL ctx/handles");
  MMR_ARG(nranks >= 1 && nranks <= 16 && rank >= 0 && rank < nranks, "bad nranks/rank (at most 16 ranks)");
  MMR_ARG(ctx->p2p_local != nullptr, "mmr_comm_p2p_export has not been called");
  DeviceGuard g(ctx->device);
  for (int r = 0; r < nranks; ++r) {
    if (r == rank) {
      ctx->p2p_peer[r] = ctx->p2p_local;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + 64 * r, 64);
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      mmr_set_error("cudaIpcOpenMemHandle for rank %d failed (%s); the triangular sweeps fall back to NCCL broadcasts",
                    r, cudaGetErrorString(e));
      ctx->p2p_ready = false;
      return 1000 + (int)e;
    }
    ctx->p2p_peer[r] = ptr;
  }
  if (!ctx->p2p_dev_table) MMR_CUDA(cudaMalloc((void**)&ctx->p2p_dev_table, sizeof(void*) * 16));
  MMR_CUDA(cudaMemcpy(ctx->p2p_dev_table, ctx->p2p_peer, sizeof(void*) * 16, cudaMemcpyHostToDevice));
  ctx->p2p_ready = true;
  ctx->p2p_epoch = 0;
  return 0;
}

extern "C" int mmr_comm_p2p_disable(mmr_ctx* ctx) {
  if (ctx) ctx->p2p_ready = false;
  return 0;
}

static void p2p_release(mmr_ctx* ctx) {
  for (int r = 0; r < 16; ++r) {
    if (ctx->p2p_peer[r] && ctx->p2p_peer[r] != ctx->p2p_local) cudaIpcCloseMemHandle(ctx->p2p_peer[r]);
    ctx->p2p_peer[r] = nullptr;
  }
  if (ctx->p2p_dev_table) cudaFree(ctx->p2p_dev_table);
  if (ctx->p2p_local) cudaFree(ctx->p2p_local);
  ctx->p2p_dev_table = nullptr;
  ctx->p2p_local = nullptr;
  ctx->p2p_ready = false;
  cudaGetLastError();
}

extern "C" int mmr_comm_destroy(mmr_ctx* ctx) {
  if (!ctx) return 0;
  {
    DeviceGuard g0(ctx->device);
    p2p_release(ctx);
  }
  if (!ctx->nccl_comm) return 0;
  DeviceGuard g(ctx->device);
  g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm);
  ctx->nccl_comm = nullptr;
  ctx->nranks = 1;
  ctx->rank = 0;
  return 0;
}

int mmr_comm_allgather_f64(mmr_ctx* ctx, const double* send, double* recv, size_t count, cudaStream_t s) {
  if (!ctx->nccl_comm) {
    mmr_set_error("multi-GPU call without mmr_comm_init");
    return -6;
  }
  ncclResult_t r = g_nccl.AllGather(send, recv, count, kNcclFloat64, (ncclComm_t)ctx->nccl_comm, s);
  if (r != 0) return fail_nccl(r, "ncclAllGather");
  return 0;
}

int mmr_comm_bcast_bytes(mmr_ctx* ctx, void* buf, size_t bytes, int root, cudaStream_t s) {
  if (!ctx->nccl_comm) {
    mmr_set_error("multi-GPU call without mmr_comm_init");
    return -6;
  }
  ncclResult_t r = g_nccl.Broadcast(buf, buf, bytes, kNcclChar, root, (ncclComm_t)ctx->nccl_comm, s);
  if (r != 0) return fail_nccl(r, "ncclBroadcast");
  return 0;
}
