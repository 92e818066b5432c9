// Measures the FP64 roofline denominators MEASURED_PEAKS.json lacks:
//   (1) FP64 FMA-pipe peak: independent DFMA chains, all SMs
//   (2) FP64 tensor (DMMA.8x8x4) peak: independent mma.sync.m8n8k4.f64 chains, all SMs
// Prints one JSON line.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peaks fp64_peaks.cu
#include <cuda_runtime.h>
#include <cstdio>

__global__ void dfma_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__global__ void dmma_kernel(double* out, int iters, double a, double b) {
  double c[8][2];
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = threadIdx.x + i;
  double av = a + threadIdx.x * 1e-9, bv = b;
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

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best_fma = 0, best_mma = 0;
  for (int warps = 4; warps <= 32; warps *= 2) {
    for (int ctas = 1; ctas <= 2; ++ctas) {
      const int threads = warps * 32 / ctas;
      if (threads > 1024 || threads < 32) continue;
      const int iters = 20000;
      for (int rep = 0; rep < 3; ++rep) {
        dfma_kernel<<<sms * ctas, threads>>>(out, iters, 1.0000001, 1e-9);
      }
      cudaEventRecord(e0);
      dfma_kernel<<<sms * ctas, threads>>>(out, iters, 1.0000001, 1e-9);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double tf = 2.0 * 8 * iters * (double)threads * sms * ctas / (ms * 1e-3) / 1e12;
      if (tf > best_fma) best_fma = tf;
      for (int rep = 0; rep < 3; ++rep) dmma_kernel<<<sms * ctas, threads>>>(out, iters / 4, 1.0000001, 1e-9);
      cudaEventRecord(e0);
      dmma_kernel<<<sms * ctas, threads>>>(out, iters / 4, 1.0000001, 1e-9);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
      // one m8n8k4 = 8*8*4 FMA = 512 flop per warp
      double tm = 512.0 * 8 * (iters / 4) * (double)(threads / 32) * sms * ctas / (ms * 1e-3) / 1e12;
      if (tm > best_mma) best_mma = tm;
      fprintf(stderr, "warps/SM %d (ctas %d): dfma %.2f TF, dmma %.2f TF\n", warps, ctas, tf, tm);
    }
  }
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"fp64_fma_tflops\": %.3f, \"fp64_dmma_tflops\": %.3f}\n", p.name, sms, best_fma, best_mma);
  return 0;
}
