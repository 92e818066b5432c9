// Stand-alone laboratory for the LU trailing update C -= A B (FP64 DMMA): times the per-tile kernel of
// csrc/dgemm.cuh at 2 and 1 CTAs per SM (what a lone CTA makes of the DMMA pipe) and the persistent kernel of
// csrc/dgemm_tma.cuh, over k; checks that both give bit-identical results.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o tools/gemm_lab tools/gemm_lab.cu
// Run  : tools/gemm_lab [m] [n]      (JSON lines on stdout)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>

#include "../magpylib_material_response_b200/csrc/dgemm_tma.cuh"

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("{\"error\": \"%s at %s:%d\"}\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

__global__ void fill_kernel(double* p, size_t n, unsigned seed) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    unsigned long long x = (i + 1) * 0x9E3779B97F4A7C15ull + seed;
    x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
    p[i] = ((double)(x & 0xFFFFFF) / 16777216.0 - 0.5);
  }
}
__global__ void diff_kernel(const double* a, const double* b, size_t n, unsigned long long* cnt, double* maxabs) {
  unsigned long long c = 0;
  double mx = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    if (__double_as_longlong(a[i]) != __double_as_longlong(b[i])) ++c;
    mx = fmax(mx, fabs(a[i] - b[i]));
  }
  if (c) atomicAdd(cnt, c);
  if (mx > 0) atomicMax((unsigned long long*)maxabs, (unsigned long long)__double_as_longlong(mx));
}

using namespace mmr_gemm;

int main(int argc, char** argv) {
  const int m = argc > 1 ? atoi(argv[1]) : 12288;
  const int n = argc > 2 ? atoi(argv[2]) : m;
  const int reps = 5;
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const int kmax = 2048;
  const int64_t lda = m + 16, ldb = kmax + 16, ldc = m + 16;
  double *A, *B, *C0, *C1, *C2;
  CK(cudaMalloc(&A, sizeof(double) * lda * kmax));
  CK(cudaMalloc(&B, sizeof(double) * ldb * n));
  CK(cudaMalloc(&C0, sizeof(double) * ldc * n));
  CK(cudaMalloc(&C1, sizeof(double) * ldc * n));
  CK(cudaMalloc(&C2, sizeof(double) * ldc * n));
  unsigned long long* cnt;
  double* mx;
  CK(cudaMalloc(&cnt, 8));
  CK(cudaMalloc(&mx, 8));
  fill_kernel<<<1024, 256>>>(A, (size_t)lda * kmax, 1);
  fill_kernel<<<1024, 256>>>(B, (size_t)ldb * n, 2);
  fill_kernel<<<1024, 256>>>(C0, (size_t)ldc * n, 3);
  CK(cudaDeviceSynchronize());
  CK(cudaFuncSetAttribute(dgemm_nn_fast_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int tiles_m = (m + BM - 1) / BM, tiles_n = (n + BN - 1) / BN;

  auto run_tile = [&](double* C, int k, size_t smem, int pf) {
    dgemm_nn_fast_kernel<true><<<dim3(tiles_m, tiles_n), THREADS, smem>>>(m, n, k, -1.0, A, lda, B, ldb, 1.0, C, ldc, pf,
                                                                           nullptr, 0, 0, GemmAlt());
  };
  typedef void (*pk_t)(int, int, int, const double*, int64_t, const double*, int64_t, double*, int64_t, int, int, const int*, int);
  const pk_t pks[8] = {dgemm_sub_persist_kernel<false, 0>, dgemm_sub_persist_kernel<true, 0>, dgemm_sub_persist_kernel<true, 3>,
                       nullptr, nullptr, nullptr, nullptr, nullptr};
  const char* pnames[8] = {"persist_stg", "persist_bulk", "persist_bulk_noio", "tile_tma", "tile_tma_noout", "tile_tma_noin", "tile_tma_noio", ""};
  typedef void (*tk_t)(int, int, int, const double*, int64_t, const double*, int64_t, double*, int64_t, const int*, int);
  const tk_t tks[4] = {dgemm_sub_tma_kernel<0>, dgemm_sub_tma_kernel<1>, dgemm_sub_tma_kernel<2>, dgemm_sub_tma_kernel<3>};
  for (int i = 0; i < 4; ++i) CK(cudaFuncSetAttribute(tks[i], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T_SMEM_BYTES));
  for (int i = 0; i < 3; ++i) CK(cudaFuncSetAttribute(pks[i], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P_SMEM_BYTES));
  auto run_pers = [&](int which, double* C, int k, int grid) {
    if (which < 3)
      pks[which]<<<grid, P_THREADS, P_SMEM_BYTES>>>(m, n, k, A, lda, B, ldb, C, ldc, tiles_m, tiles_m * tiles_n, nullptr, 0);
    else
      tks[which - 3]<<<dim3(tiles_m, tiles_n), THREADS, T_SMEM_BYTES>>>(m, n, k, A, lda, B, ldb, C, ldc, nullptr, 0);
  };
  auto time_it = [&](auto&& fn, const char* name, int k) {
    fn();  // warm-up
    CK(cudaDeviceSynchronize());
    std::vector<float> ts;
    for (int r = 0; r < reps; ++r) {
      CK(cudaEventRecord(e0));
      fn();
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      ts.push_back(ms);
    }
    CK(cudaGetLastError());
    std::sort(ts.begin(), ts.end());
    const double fl = 2.0 * m * (double)n * k;
    printf("{\"kernel\": \"%s\", \"m\": %d, \"n\": %d, \"k\": %d, \"ms_best\": %.4f, \"ms_median\": %.4f, \"tflops_best\": %.3f, \"tflops_median\": %.3f}\n",
           name, m, n, k, ts[0], ts[reps / 2], fl / ts[0] * 1e-9, fl / ts[reps / 2] * 1e-9);
    fflush(stdout);
  };

  // correctness: one application of each kernel to the same C, k = 256 and 48 (KT = 3: tile changes inside the prologue depth)
  for (int v : {1, 3})
  for (int k : {256, 48, 16}) {
    CK(cudaMemcpy(C1, C0, sizeof(double) * ldc * n, cudaMemcpyDeviceToDevice));
    CK(cudaMemcpy(C2, C0, sizeof(double) * ldc * n, cudaMemcpyDeviceToDevice));
    run_tile(C1, k, SMEM_BYTES, 2 * sms);
    run_pers(v, C2, k, sms);
    CK(cudaDeviceSynchronize());
    CK(cudaMemset(cnt, 0, 8));
    CK(cudaMemset(mx, 0, 8));
    diff_kernel<<<1024, 256>>>(C1, C2, (size_t)ldc * n, cnt, mx);
    unsigned long long hc;
    double hm;
    CK(cudaMemcpy(&hc, cnt, 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&hm, mx, 8, cudaMemcpyDeviceToHost));
    CK(cudaMemset(cnt, 0, 8));
    diff_kernel<<<1024, 256>>>(C0, C2, (size_t)ldc * n, cnt, mx);
    unsigned long long hc0;
    CK(cudaMemcpy(&hc0, cnt, 8, cudaMemcpyDeviceToHost));
    printf("{\"check\": \"%s vs per-tile\", \"k\": %d, \"mismatching_entries\": %llu, \"max_abs_diff\": %.3e, \"entries_changed_vs_input\": %llu}\n",
           pnames[v], k, hc, hm, hc0);
    fflush(stdout);
  }
  for (int k : {256, 512, 2048}) {
    time_it([&] { run_tile(C1, k, SMEM_BYTES, 2 * sms); }, "per_tile_2cta", k);
    time_it([&] { run_tile(C1, k, SMEM_BYTES + 48 * 1024, sms); }, "per_tile_1cta", k);
    for (int v = 0; v < 7; ++v) time_it([&] { run_pers(v, C2, k, sms); }, pnames[v], k);
  }
  // the shapes of a late LU step (few tiles per SM) at k = 256
  return 0;
}
