// Context, error strings and workspace management for libmmr_b200.so.
#include <cstdlib>

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
int mmr_func_smem(mmr_ctx* ctx, const void* func, size_t bytes) {
  size_t& have = ctx->func_smem[func];
  if (bytes > have) {
    MMR_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    have = bytes;
  }
  return 0;
}

int mmr_hpin_reserve(mmr_ctx* ctx, size_t bytes) {
  if (ctx->hpin_bytes < bytes && ctx->hpin) cudaStreamSynchronize(ctx->stream);
  return reserve(&ctx->hpin, &ctx->hpin_bytes, bytes, true);
}

static cudaEvent_t prof_event(mmr_ctx* ctx) {
  if (!ctx->prof_pool.empty()) {
    cudaEvent_t e = ctx->prof_pool.back();
    ctx->prof_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

int mmr_prof_begin(mmr_ctx* ctx, int kind, cudaStream_t s) {
  if (!ctx->profiling || !((ctx->prof_mask >> kind) & 1u)) return -1;
  mmr_ctx::ProfRec r;
  r.kind = kind;
  r.side = (s == ctx->side) ? 1 : ((ctx->comm && s == ctx->comm) ? 2 : ((ctx->aux && s == ctx->aux) ? 3 : 0));
  r.e0 = prof_event(ctx);
  r.e1 = prof_event(ctx);
  r.units = 0.0;
  cudaEventRecord(r.e0, s);
  ctx->prof.push_back(r);
  return (int)ctx->prof.size() - 1;
}

void mmr_prof_end(mmr_ctx* ctx, int idx, double units, cudaStream_t s) {
  if (idx < 0) return;
  ctx->prof[idx].units = units;
  cudaEventRecord(ctx->prof[idx].e1, s);
}

extern "C" int mmr_ctx_set_profiling(mmr_ctx* ctx, int on) {
  MMR_ARG(ctx != nullptr, "ctx is NULL");
  DeviceGuard g(ctx->device);
  MMR_CUDA(cudaStreamSynchronize(ctx->stream));
  for (auto& r : ctx->prof) {
    ctx->prof_pool.push_back(r.e0);
    ctx->prof_pool.push_back(r.e1);
  }
  ctx->prof.clear();
  ctx->profiling = on != 0;
  ctx->prof_mask = (on == 0 || on == 1) ? ~0u : ((unsigned)on >> 1);  // on = 1: every class; else bit k+1 <-> class k
  if (on) {
    if (!ctx->prof_base) MMR_CUDA(cudaEventCreate(&ctx->prof_base));
    MMR_CUDA(cudaEventRecord(ctx->prof_base, ctx->stream));
  }
  return 0;
}

extern "C" int64_t mmr_ctx_profile_timeline(mmr_ctx* ctx, int* kind_out, double* start_ms_out, double* dur_ms_out,
                                            int64_t cap) {
  if (!ctx || !ctx->prof_base) return 0;
  DeviceGuard g(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->side) cudaStreamSynchronize(ctx->side);
  if (ctx->comm) cudaStreamSynchronize(ctx->comm);
  if (ctx->aux) cudaStreamSynchronize(ctx->aux);
  int64_t cnt = 0;
  for (auto& r : ctx->prof) {
    if (cnt >= cap) break;
    float t0 = 0.f, dt = 0.f;
    if (cudaEventElapsedTime(&t0, ctx->prof_base, r.e0) != cudaSuccess) continue;
    if (cudaEventElapsedTime(&dt, r.e0, r.e1) != cudaSuccess) continue;
    if (kind_out) kind_out[cnt] = r.kind | (r.side << 8);
    if (start_ms_out) start_ms_out[cnt] = t0;
    if (dur_ms_out) dur_ms_out[cnt] = dt;
    ++cnt;
  }
  cudaGetLastError();
  return cnt;
}

extern "C" int mmr_ctx_profile(mmr_ctx* ctx, int kind, double* ms_out, double* units_out,
                               int64_t* launches_out) {
  MMR_ARG(ctx != nullptr, "ctx is NULL");
  MMR_ARG(kind >= 0 && kind < MMR_PROF_NKINDS, "unknown profile kind");
  DeviceGuard g(ctx->device);
  MMR_CUDA(cudaStreamSynchronize(ctx->stream));
  double ms = 0.0, units = 0.0;
  int64_t cnt = 0;
  for (auto& r : ctx->prof) {
    if (r.kind != kind) continue;
    float t = 0.f;
    MMR_CUDA(cudaEventElapsedTime(&t, r.e0, r.e1));
    ms += t;
    units += r.units;
    ++cnt;
  }
  if (ms_out) *ms_out = ms;
  if (units_out) *units_out = units;
  if (launches_out) *launches_out = cnt;
  return 0;
}

extern "C" int64_t mmr_ctx_profile_records(mmr_ctx* ctx, int kind, double* ms_out, int64_t cap) {
  if (!ctx || !ms_out) return 0;
  DeviceGuard g(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  int64_t cnt = 0;
  for (auto& r : ctx->prof) {
    if (r.kind != kind || cnt >= cap) continue;
    float t = 0.f;
    cudaEventElapsedTime(&t, r.e0, r.e1);
    ms_out[cnt++] = t;
  }
  return cnt;
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
  {
    int lo = 0, hi = 0;
    MMR_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    MMR_CUDA(cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, hi));  // panel look-ahead stream
    MMR_CUDA(cudaStreamCreateWithPriority(&c->comm, cudaStreamNonBlocking, hi));  // broadcasts of the distributed LU
    MMR_CUDA(cudaStreamCreateWithPriority(&c->aux, cudaStreamNonBlocking, lo));   // deferred, off the critical path
    const char* la = getenv("MMR_LU_LOOKAHEAD");
    c->lookahead = !(la && la[0] == '0');
  }
  for (int i = 0; i < 4; ++i) MMR_CUDA(cudaEventCreateWithFlags(&c->ev[i], cudaEventDisableTiming));
  for (int i = 0; i < 10; ++i) MMR_CUDA(cudaEventCreateWithFlags(&c->evd[i], cudaEventDisableTiming));
  *out = c;
  return 0;
}

extern "C" int mmr_ctx_destroy(mmr_ctx* ctx) {
  if (!ctx) return 0;
  DeviceGuard g(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  mmr_comm_destroy(ctx);
  for (auto& r : ctx->prof) {
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  for (auto e : ctx->prof_pool) cudaEventDestroy(e);
  if (ctx->prof_base) cudaEventDestroy(ctx->prof_base);
  if (ctx->laswp_buf) cudaFree(ctx->laswp_buf);
  if (ctx->ws) cudaFree(ctx->ws);
  if (ctx->ws2) cudaFree(ctx->ws2);
  if (ctx->hpin) cudaFreeHost(ctx->hpin);
  for (int i = 0; i < 4; ++i)
    if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  for (int i = 0; i < 10; ++i)
    if (ctx->evd[i]) cudaEventDestroy(ctx->evd[i]);
  if (ctx->side) cudaStreamDestroy(ctx->side);
  if (ctx->comm) cudaStreamDestroy(ctx->comm);
  if (ctx->aux) cudaStreamDestroy(ctx->aux);
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
