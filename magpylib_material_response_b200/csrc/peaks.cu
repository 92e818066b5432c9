// Roofline denominators for the FP64 kernels, measured on the device the context runs on
// (MEASURED_PEAKS.json carries HBM and bf16 only):
//   (1) FP64 FMA pipe: independent DFMA chains on all SMs           -> assembly (kernel 1)
//   (2) FP64 tensor pipe: independent DMMA.8x8x4 chains on all SMs  -> LU trailing update (kernel 2a)
// bench.py calls this right before its timed region, so the roofline fraction does not rest on a
// number measured on another day / another box.
#include "common.cuh"

namespace {

__global__ void peak_dfma_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__global__ void peak_dmma_kernel(double* out, int iters, double a, double b) {
  double c[8][2];
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = threadIdx.x + i;
  const double av = a + threadIdx.x * 1e-9, bv = b;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(av), "d"(bv));
  }
  double s = 0;
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace

extern "C" int mmr_fp64_peaks(mmr_ctx* ctx, double* fma_tflops_out, double* dmma_tflops_out) {
  MMR_ARG(ctx != nullptr, "ctx is NULL");
  DeviceGuard g(ctx->device);
  const int sms = ctx->sm_count;
  int rc = mmr_ws_reserve(ctx, sizeof(double) * (size_t)sms * 2 * 1024);
  if (rc) return rc;
  double* out = (double*)ctx->ws;
  cudaStream_t s = ctx->stream;
  cudaEvent_t e0, e1;
  MMR_CUDA(cudaEventCreate(&e0));
  MMR_CUDA(cudaEventCreate(&e1));
  double best_fma = 0.0, best_mma = 0.0;
  const int iters = 8000;
  // resident warps per SM: 8 .. 32, as one or two CTAs (the best configuration wins)
  for (int warps = 8; warps <= 32; warps *= 2) {
    for (int ctas = 1; ctas <= 2; ++ctas) {
      const int threads = warps * 32 / ctas;
      if (threads > 1024) continue;
      float ms = 0.f;
      peak_dfma_kernel<<<sms * ctas, threads, 0, s>>>(out, iters, 1.0000001, 1e-9);  // warm-up
      MMR_CUDA(cudaEventRecord(e0, s));
      peak_dfma_kernel<<<sms * ctas, threads, 0, s>>>(out, iters, 1.0000001, 1e-9);
      MMR_CUDA(cudaEventRecord(e1, s));
      MMR_CUDA(cudaEventSynchronize(e1));
      MMR_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      const double tf = 2.0 * 8 * iters * (double)threads * sms * ctas / (ms * 1e-3) / 1e12;
      if (tf > best_fma) best_fma = tf;
      peak_dmma_kernel<<<sms * ctas, threads, 0, s>>>(out, iters / 4, 1.0000001, 1e-9);
      MMR_CUDA(cudaEventRecord(e0, s));
      peak_dmma_kernel<<<sms * ctas, threads, 0, s>>>(out, iters / 4, 1.0000001, 1e-9);
      MMR_CUDA(cudaEventRecord(e1, s));
      MMR_CUDA(cudaEventSynchronize(e1));
      MMR_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      // one m8n8k4 = 8*8*4 FMA = 512 flop per warp
      const double tm = 512.0 * 8 * (iters / 4) * (double)(threads / 32) * sms * ctas / (ms * 1e-3) / 1e12;
      if (tm > best_mma) best_mma = tm;
      ctx->launches += 4;
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  MMR_CUDA(cudaGetLastError());
  if (fma_tflops_out) *fma_tflops_out = best_fma;
  if (dmma_tflops_out) *dmma_tflops_out = best_mma;
  return 0;
}
