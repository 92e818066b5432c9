// Shared context, error plumbing and launch accounting for libmmr_b200.so (see include/mmr_b200.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <unordered_map>
#include <vector>

#include "../../include/mmr_b200.h"

#define MMR_PI 3.14159265358979323846
#define MMR_MU0 1.25663706127e-06  // scipy.constants.mu_0 == magpy.mu_0 (demag.py:451)

struct mmr_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  int sm_count = 148;
  int64_t launches = 0;
  // grow-only device workspace
  void* ws = nullptr;
  size_t ws_bytes = 0;
  // second grow-only workspace (LU panels / GMRES basis) so both can be live at once
  void* ws2 = nullptr;
  size_t ws2_bytes = 0;
  // pinned host scratch for small device->host reads
  void* hpin = nullptr;
  size_t hpin_bytes = 0;
  // NCCL (dlopen'ed lazily, comm.cu)
  void* nccl_comm = nullptr;
  int nranks = 1;
  int rank = 0;
  // peer-memory exchange buffers of the multi-GPU triangular sweeps (comm.cu: mmr_comm_p2p_*): one buffer per
  // rank, cudaMalloc'ed by its owner and opened by every peer through CUDA IPC.  Layout of a buffer:
  // [2 x P2P_CAP doubles (forward / backward solution blocks) | P2P_FLAGS ints (epoch of every block)]
  void* p2p_local = nullptr;        // my buffer
  void* p2p_peer[16] = {nullptr};   // every rank's buffer as seen from this device ([rank] = p2p_local)
  void** p2p_dev_table = nullptr;   // device copy of p2p_peer
  bool p2p_ready = false;
  int p2p_epoch = 0;
  // optional profiling records (mmr_ctx_set_profiling)
  bool profiling = false;
  unsigned prof_mask = ~0u;  // bit k: record kernel class k (mmr_ctx_set_profiling)
  struct ProfRec {
    int kind;
    int side;  // 0: main stream, 1: side (look-ahead) stream, 2: communication stream
    cudaEvent_t e0, e1;
    double units;
  };
  std::vector<ProfRec> prof;
  std::vector<cudaEvent_t> prof_pool;
  cudaEvent_t prof_base = nullptr;  // recorded by mmr_ctx_set_profiling(on): origin of the timeline
  // side stream + events for look-ahead in the LU
  cudaStream_t side = nullptr;
  bool lookahead = true;
  void* laswp_buf = nullptr;  // LaswpPerm scratch (lu.cu)
  // kernel attributes are per device: what this context has already raised / probed (never process-wide
  // statics, so that contexts on several devices of one process work)
  std::unordered_map<const void*, size_t> func_smem;  // MaxDynamicSharedMemorySize set so far, per kernel
  int panel_cluster_max = -1;                          // largest launchable strip cluster (-1: not probed)
  bool panel_reg_ok = false;                           // non-portable cluster sizes allowed for the register strips
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  // third stream + events of the distributed LU: broadcasts run beside the panel chain
  cudaStream_t comm = nullptr;
  cudaStream_t aux = nullptr;  // low priority: deferred work of the single-GPU LU (U12 rows finalised off the critical path)
  cudaEvent_t evd[10] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};

void mmr_set_error(const char* fmt, ...);
int mmr_fail_cuda(cudaError_t e, const char* what, const char* file, int line);

#define MMR_CUDA(call)                                                 \
  do {                                                                 \
    cudaError_t _e = (call);                                           \
    if (_e != cudaSuccess) return mmr_fail_cuda(_e, #call, __FILE__, __LINE__); \
  } while (0)

#define MMR_CHECK_LAUNCH(ctx)                                          \
  do {                                                                 \
    (ctx)->launches++;                                                 \
    cudaError_t _e = cudaGetLastError();                               \
    if (_e != cudaSuccess) return mmr_fail_cuda(_e, "kernel launch", __FILE__, __LINE__); \
  } while (0)

#define MMR_ARG(cond, msg)                          \
  do {                                              \
    if (!(cond)) {                                  \
      mmr_set_error("invalid argument: %s", msg);   \
      return -1;                                    \
    }                                               \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

int mmr_ws_reserve(mmr_ctx* ctx, size_t bytes);    // ctx->ws
int mmr_ws2_reserve(mmr_ctx* ctx, size_t bytes);   // ctx->ws2
int mmr_hpin_reserve(mmr_ctx* ctx, size_t bytes);  // ctx->hpin
// raise a kernel's dynamic shared-memory limit to `bytes` on this context's device (once per size)
int mmr_func_smem(mmr_ctx* ctx, const void* func, size_t bytes);

// profiling scope: records an event pair around the launches enqueued in between
int mmr_prof_begin(mmr_ctx* ctx, int kind, cudaStream_t s);
void mmr_prof_end(mmr_ctx* ctx, int idx, double units, cudaStream_t s);

constexpr int64_t MMR_P2P_CAP = 1 << 18;  // doubles per direction (N <= 262144)
constexpr int MMR_P2P_FLAGS = 4096;       // blocks per sweep
static inline size_t mmr_p2p_bytes() { return sizeof(double) * 2 * MMR_P2P_CAP + sizeof(int) * 2 * MMR_P2P_FLAGS; }

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
