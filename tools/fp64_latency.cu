// Dependent-chain latencies (SM cycles) of the instructions on the critical path of the serial panel kernels:
// DFMA, DMUL, MUFU.RCP64H + Newton, SHFL of a double, DMMA.8x8x4, LDS.64, and a CTA barrier (512 threads).
// One warp (or CTA) on one SM.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu
#include <cuda_runtime.h>
#include <cstdio>

__global__ void lat_kernel(double* out, long long* cyc, int iters, double a, double b) {
  __shared__ double sm[1024];
  const int lane = threadIdx.x & 31;
  sm[threadIdx.x] = a + threadIdx.x;
  __syncthreads();
  double x = a + lane * 1e-9, y = b;
  long long t0, t1;
  int k = 0;
  // DFMA chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < iters; ++i) x = fma(x, y, b);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[k] = t1 - t0;
  ++k;
  // DMUL chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < iters; ++i) x = x * y;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[k] = t1 - t0;
  ++k;
  // rcp.approx + 2 Newton steps chain
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < iters; ++i) {
    double r;
    asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(r, fma(-x, r, 1.0), r);
    r = fma(r, fma(-x, r, 1.0), r);
    x = r + 1.5;
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[k] = t1 - t0;
  ++k;
  // bare MUFU.RCP64H chain
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < iters; ++i) {
    double r;
    asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    x = r;
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[k] = t1 - t0;
  ++k;
  // SHFL (double = 2 x SHFL.IDX) chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < iters; ++i) x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[k] = t1 - t0;
  ++k;
  // DMMA chain (same accumulator)
  double c0 = x, c1 = y;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < iters; ++i)
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
  t1 = clock64();
  if (threadIdx.x == 0) cyc[k] = t1 - t0;
  ++k;
  x += c0 + c1;
  // LDS.64 pointer-chase-like chain (address depends on the loaded value)
  int idx = lane;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < iters; ++i) {
    const double v = sm[idx];
    idx = (idx + (int)(v * 0.0) + 1) & 1023;
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[k] = t1 - t0;
  ++k;
  x += idx;
  // CTA barrier
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < iters; ++i) __syncthreads();
  t1 = clock64();
  if (threadIdx.x == 0) cyc[k] = t1 - t0;
  ++k;
  // DFMA throughput of ONE warp: 8 independent chains
  double z0 = x, z1 = x + 1, z2 = x + 2, z3 = x + 3, z4 = x + 4, z5 = x + 5, z6 = x + 6, z7 = x + 7;
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < iters; ++i) {
    z0 = fma(z0, y, b); z1 = fma(z1, y, b); z2 = fma(z2, y, b); z3 = fma(z3, y, b);
    z4 = fma(z4, y, b); z5 = fma(z5, y, b); z6 = fma(z6, y, b); z7 = fma(z7, y, b);
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[k] = t1 - t0;
  ++k;
  out[blockIdx.x * blockDim.x + threadIdx.x] = x + z0 + z1 + z2 + z3 + z4 + z5 + z6 + z7;
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8 * 16);
  const int iters = 4096;
  const char* names[] = {"dfma", "dmul", "rcp+2newton+add", "mufu_rcp64h", "shfl_f64", "dmma884", "lds64_dependent", "bar_sync", "dfma_x8_one_warp"};
  for (int threads : {32, 512}) {
    lat_kernel<<<1, threads>>>(out, cyc, iters, 1.0000001, 0.9999999);
    lat_kernel<<<1, threads>>>(out, cyc, iters, 1.0000001, 0.9999999);
    cudaDeviceSynchronize();
    long long h[16];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("{\"threads\": %d", threads);
    for (int k = 0; k < 9; ++k) printf(", \"%s\": %.1f", names[k], (double)h[k] / iters);
    printf("}\n");
  }
  return 0;
}
