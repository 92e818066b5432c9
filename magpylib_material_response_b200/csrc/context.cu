// Context, error strings and workspace management for libmmr_b200.so.
#include "common.cuh"

static thread_local char g_err[1024] = "";

void mmr_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int mmr_fail_cuda(cudaError_t e, const char* what, const char* file, int line) {
  mmr_set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
  return 1000 + (int)e;
}

static int reserve(void** p, size_t* have, size_t bytes, bool pinned) {
  if (*have >= bytes) return 0;
  if (*p) {
    cudaError_t e = pinned ? cudaFreeHost(*p) : cudaFree(*p);
    if (e != cudaSuccess) return mmr_fail_cuda(e, "free workspace", __FILE__, __LINE__);
    *p = nullptr;
    *have = 0;
  }
  cudaError_t e = pinned ? cudaMallocHost(p, bytes) : cudaMalloc(p, bytes);
  if (e != cudaSuccess) return mmr_fail_cuda(e, "allocate workspace", __FILE__, __LINE__);
  *have = bytes;
  return 0;
}

int mmr_ws_reserve(mmr_ctx* ctx, size_t bytes) {
  // the stream may still be using the old buffer
  if (ctx->ws_bytes < bytes && ctx->ws) cudaStreamSynchronize(ctx->stream);
  return reserve(&ctx->ws, &ctx->ws_bytes, bytes, false);
}
int mmr_ws2_reserve(mmr_ctx* ctx, size_t bytes) {
  if (ctx->ws2_bytes < bytes && ctx->ws2) cudaStreamSynchronize(ctx->stream);
  return reserve(&ctx->ws2, &ctx->ws2_bytes, bytes, false);
}
int mmr_hpin_reserve(mmr_ctx* ctx, size_t bytes) {
  if (ctx->hpin_bytes < bytes && ctx->hpin) cudaStreamSynchronize(ctx->stream);
  return reserve(&ctx->hpin, &ctx->hpin_bytes, bytes, true);
}

extern "C" int mmr_abi_version(void) { return MMR_ABI_VERSION; }
extern "C" const char* mmr_last_error(void) { return g_err; }

extern "C" int mmr_ctx_create(int device, void* stream, mmr_ctx** out) {
  MMR_ARG(out != nullptr, "out is NULL");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    mmr_set_error("no CUDA device available (%s); this library has no CPU fallback",
                  cudaGetErrorString(e));
    return 1000 + (int)(e == cudaSuccess ? cudaErrorNoDevice : e);
  }
  MMR_ARG(device >= 0 && device < count, "device index out of range");
  DeviceGuard g(device);
  mmr_ctx* c = new mmr_ctx();
  c->device = device;
  c->stream = (cudaStream_t)stream;
  cudaDeviceProp prop;
  MMR_CUDA(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  if (prop.major < 10) {
    mmr_set_error("device %d is sm_%d%d; libmmr_b200 is built for sm_100a only", device, prop.major,
                  prop.minor);
    delete c;
    return -2;
  }
  MMR_CUDA(cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking));
  for (int i = 0; i < 4; ++i) MMR_CUDA(cudaEventCreateWithFlags(&c->ev[i], cudaEventDisableTiming));
  *out = c;
  return 0;
}

extern "C" int mmr_ctx_destroy(mmr_ctx* ctx) {
  if (!ctx) return 0;
  DeviceGuard g(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  mmr_comm_destroy(ctx);
  if (ctx->ws) cudaFree(ctx->ws);
  if (ctx->ws2) cudaFree(ctx->ws2);
  if (ctx->hpin) cudaFreeHost(ctx->hpin);
  for (int i = 0; i < 4; ++i)
    if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  if (ctx->side) cudaStreamDestroy(ctx->side);
  delete ctx;
  return 0;
}

extern "C" int mmr_ctx_sync(mmr_ctx* ctx) {
  MMR_ARG(ctx != nullptr, "ctx is NULL");
  DeviceGuard g(ctx->device);
  MMR_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
}

extern "C" int64_t mmr_ctx_launch_count(mmr_ctx* ctx) { return ctx ? ctx->launches : 0; }
