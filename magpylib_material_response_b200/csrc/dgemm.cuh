// FP64 GEMM on the tensor-core DMMA path (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4 on sm_100a;
// tcgen05 has no FP64 kind and wgmma is sm_90a-only).  This is the trailing-update engine of
// the blocked LU (the ~2/3 N^3 flops of np.linalg.solve, demag.py:470).
//
//   C (m x n) = alpha * A (m x k) * B (k x n) + beta * C      column-major, "N","N"
//
// CTA tile 128 x 64, BK = 16, 8 warps (4 along m x 2 along n), warp tile 32 x 32 = 4 x 4 DMMA
// fragments (32 FP64 accumulators per thread), 3-stage cp.async pipeline, 2 CTAs per SM so one
// CTA's C read-modify-write epilogue overlaps the other's main loop.  Shared-memory rows are
// padded by 4 doubles so that both fragment load patterns are bank-conflict free.
#pragma once

#include "common.cuh"

namespace mmr_gemm {

constexpr int BM = 128, BN = 64, BK = 16, STAGES = 3, THREADS = 256;
constexpr int LDA_S = BM + 4;  // doubles per k-row of the A stage   ((k*132 + m) mod 16 distinct)
constexpr int LDB_S = BK + 4;  // doubles per n-row of the B stage   ((n*20 + k) mod 16 distinct)
constexpr int A_STAGE = BK * LDA_S;
constexpr int B_STAGE = BN * LDB_S;
constexpr size_t SMEM_BYTES = (size_t)STAGES * (A_STAGE + B_STAGE) * sizeof(double);

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, int src_bytes) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// Speculation guard of the LU ("poison"): `fail` points to a sticky device word that holds kb+1 of the first
// panel whose speculative (diagonal-pivot) factorization was rejected, or 0.  An operation that needs the
// result of the panel starting at column `dep` (and hence of all earlier ones) must not write anything once a
// panel with start <= dep has failed; operations that only depend on earlier panels still run, so the matrix
// stays in a consistent "everything before the failed panel applied" state the host can resume from.
__device__ __forceinline__ bool lu_poisoned(int failword, int dep) { return failword != 0 && failword - 1 <= dep; }
__device__ __forceinline__ int lu_failword(const int* fail) { return fail ? __ldcg(fail) : 0; }
// (the GEMM kernels take one more word, check_val: when non-zero, a stored value with |c| > 1 (or NaN) writes
// check_val into *fail -- the a-posteriori test of a speculative panel's multipliers, fused into the epilogue
// of the L21 = A21 U11^-1 product)

// Sign flip on the integer pipe (a DADD/DMUL negation would compete with the DMMAs for the FP64 pipe).
__device__ __forceinline__ double flip_sign(double x) {
  return __hiloint2double(__double2hiint(x) ^ (int)0x80000000, __double2loint(x));
}

// A second, independent product that shares a launch with the first one (blockIdx.z == 1): the LU panel chain
// is a sequence of small dependent kernels whose cost is launch + scheduling latency, so two products that
// depend on the same predecessor (L21 = A21 U11^-1 and U12 = L11^-1 A12) go out together.  m == 0: none.
struct GemmAlt {
  int m = 0, n = 0, k = 0;
  const double* A = nullptr;
  int64_t lda = 0;
  const double* B = nullptr;
  int64_t ldb = 0;
  double* C = nullptr;
  int64_t ldc = 0;
};

// ALIGNED16: A, B 16-byte aligned with even leading dimensions (cp.async 16 B); otherwise 8 B.
// SUBTRACT : the LU form C <- C - A*B (alpha = -1, beta = 1).  The C tile is loaded straight into
//            the accumulators (negated) BEFORE the main loop, so its DRAM latency overlaps the
//            cp.async prologue and the epilogue is store-only.  (ncu on the first version showed
//            50% of all stall samples on the epilogue's C loads.)
template <bool ALIGNED16, bool SUBTRACT>
__global__ void __launch_bounds__(THREADS, 2)
    dgemm_nn_kernel(int m, int n, int k, double alpha, const double* __restrict__ A, int64_t lda,
                    const double* __restrict__ B, int64_t ldb, double beta, double* __restrict__ C,
                    int64_t ldc, int* __restrict__ fail, int dep, int check_val) {
  extern __shared__ __align__(16) double smem[];
  double* As = smem;
  double* Bs = smem + STAGES * A_STAGE;
  const int failword = lu_failword(fail);  // consumed before the first store (epilogue)

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int wm = warp & 3;   // 4 warps along m
  const int wn = warp >> 2;  // 2 warps along n
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int KT = (k + BK - 1) / BK;

  if (SUBTRACT) {
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      const int gm = m0 + wm * 32 + 8 * f + (lane >> 2);
#pragma unroll
      for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int gn = n0 + wn * 32 + 8 * g + 2 * (lane & 3) + e;
          if (gm < m && gn < n) acc[f][g][e] = flip_sign(__ldcs(C + (int64_t)gn * ldc + gm));
        }
    }
  }

  auto load_tile = [&](int stage, int kt) {
    const int k0 = kt * BK;
    double* as = As + stage * A_STAGE;
    double* bs = Bs + stage * B_STAGE;
    if (ALIGNED16) {
      // A: 16 k-rows x 64 chunks of 2 doubles
#pragma unroll
      for (int it = 0; it < (BK * BM / 2) / THREADS; ++it) {
        const int idx = tid + it * THREADS;
        const int kr = idx / (BM / 2);
        const int mc = (idx % (BM / 2)) * 2;
        const int gm = m0 + mc, gk = k0 + kr;
        int bytes = 0;
        if (gk < k && gm < m) bytes = (m - gm >= 2) ? 16 : 8;
        const double* src = bytes ? (A + (int64_t)gk * lda + gm) : A;
        cp_async16(as + kr * LDA_S + mc, src, bytes);
      }
      // B: 64 n-rows x 8 chunks of 2 doubles
#pragma unroll
      for (int it = 0; it < (BN * BK / 2) / THREADS; ++it) {
        const int idx = tid + it * THREADS;
        const int nr = idx / (BK / 2);
        const int kc = (idx % (BK / 2)) * 2;
        const int gn = n0 + nr, gk = k0 + kc;
        int bytes = 0;
        if (gn < n && gk < k) bytes = (k - gk >= 2) ? 16 : 8;
        const double* src = bytes ? (B + (int64_t)gn * ldb + gk) : B;
        cp_async16(bs + nr * LDB_S + kc, src, bytes);
      }
    } else {
#pragma unroll
      for (int it = 0; it < (BK * BM) / THREADS; ++it) {
        const int idx = tid + it * THREADS;
        const int kr = idx / BM;
        const int mc = idx % BM;
        const int gm = m0 + mc, gk = k0 + kr;
        const bool ok = (gk < k && gm < m);
        cp_async8(as + kr * LDA_S + mc, ok ? (A + (int64_t)gk * lda + gm) : A, ok ? 8 : 0);
      }
#pragma unroll
      for (int it = 0; it < (BN * BK) / THREADS; ++it) {
        const int idx = tid + it * THREADS;
        const int nr = idx / BK;
        const int kc = idx % BK;
        const int gn = n0 + nr, gk = k0 + kc;
        const bool ok = (gn < n && gk < k);
        cp_async8(bs + nr * LDB_S + kc, ok ? (B + (int64_t)gn * ldb + gk) : B, ok ? 8 : 0);
      }
    }
  };

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < KT) load_tile(s, s);
    cp_async_commit();
  }

  const int a_row = wm * 32 + (lane >> 2);  // + 8*f
  const int b_row = wn * 32 + (lane >> 2);  // + 8*g
  const int kq = lane & 3;

  for (int kt = 0; kt < KT; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    const int nxt = kt + STAGES - 1;
    if (nxt < KT) load_tile(nxt % STAGES, nxt);
    cp_async_commit();

    const double* as = As + (kt % STAGES) * A_STAGE;
    const double* bs = Bs + (kt % STAGES) * B_STAGE;
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      double af[4], bf[4];
#pragma unroll
      for (int f = 0; f < 4; ++f) af[f] = as[(kk + kq) * LDA_S + a_row + 8 * f];
#pragma unroll
      for (int g = 0; g < 4; ++g) bf[g] = bs[(b_row + 8 * g) * LDB_S + kk + kq];
#pragma unroll
      for (int f = 0; f < 4; ++f)
#pragma unroll
        for (int g = 0; g < 4; ++g) dmma884(acc[f][g][0], acc[f][g][1], af[f], bf[g]);
    }
  }
  cp_async_wait<0>();
  if (lu_poisoned(failword, dep)) return;
  bool viol = false;  // check_val != 0: the speculation check |c| <= 1 on every stored value (LU panel's L21)

  // epilogue: thread holds rows (lane>>2), cols 2*(lane&3) + {0,1} of each 8x8 fragment
#pragma unroll
  for (int f = 0; f < 4; ++f) {
    const int gm = m0 + wm * 32 + 8 * f + (lane >> 2);
    if (gm >= m) continue;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int gn = n0 + wn * 32 + 8 * g + 2 * (lane & 3) + e;
        if (gn < n) {
          double* cp = C + (int64_t)gn * ldc + gm;
          if (SUBTRACT) {
            *cp = flip_sign(acc[f][g][e]);
          } else {
            double v = alpha * acc[f][g][e];
            if (beta != 0.0) v += beta * (*cp);
            *cp = v;
            if (!(fabs(v) <= 1.0)) viol = true;
          }
        }
      }
    }
  }
  if (check_val != 0 && viol) atomicCAS(fail, 0, check_val);
}

// ---------------------------------------------------------------------------------------
// Fast path: 16-byte aligned operands and k % BK == 0 (always true inside the LU).
// Compared with dgemm_nn_kernel: (1) the cp.async source pointers are computed once and advanced
// by constant strides (the generic kernel spends ~70% of its issue slots on address arithmetic and
// bounds checks, all of it right after the barrier while the tensor pipe idles); (2) the copies of
// the next stage are issued BETWEEN the four DMMA groups of a k-tile, so the address work overlaps
// tensor-pipe execution.
// Tried and rejected (measured on B200, m = n = 12288, k = 256): a persistent variant with an L2 prefetch of
// the next C tile (57% tensor-pipe activity), de-phasing the two CTAs of an SM with a start-up delay
// (no effect), 64 x 64 tiles with 4 CTAs per SM (32.2 TF against 33.1 TF for this kernel), and BK = 32 with two
// stages, i.e. half the CTA barriers at the same prefetch depth (31.1 TF).
// ---------------------------------------------------------------------------------------
template <bool SUBTRACT>
__global__ void __launch_bounds__(THREADS, 2)
    dgemm_nn_fast_kernel(int m, int n, int k, double alpha, const double* __restrict__ A, int64_t lda,
                         const double* __restrict__ B, int64_t ldb, double beta, double* __restrict__ C,
                         int64_t ldc, int prefetch_dist, int* __restrict__ fail, int dep, int check_val,
                         const GemmAlt alt) {
  extern __shared__ __align__(16) double smem[];
  double* As = smem;
  double* Bs = smem + STAGES * A_STAGE;
  const int failword = lu_failword(fail);  // consumed before the first store (epilogue)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 3, wn = warp >> 2;
  if (blockIdx.z != 0) {  // the second product of a paired launch (block-uniform)
    m = alt.m, n = alt.n, k = alt.k;
    A = alt.A, lda = alt.lda, B = alt.B, ldb = alt.ldb, C = alt.C, ldc = alt.ldc;
    check_val = 0;
  }
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  if (m0 >= m || n0 >= n) return;  // (the grid of a paired launch covers the larger product)
  const int KT = k / BK;

  double acc[4][4][2];
  if (SUBTRACT) {
    // all 32 loads are issued back to back into the accumulators, signs flipped afterwards
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      const int gm = min(m0 + wm * 32 + 8 * f + (lane >> 2), m - 1);
#pragma unroll
      for (int g = 0; g < 4; ++g)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int gn = min(n0 + wn * 32 + 8 * g + 2 * (lane & 3) + e, n - 1);
          acc[f][g][e] = __ldcs(C + (int64_t)gn * ldc + gm);
        }
    }
  } else {
#pragma unroll
    for (int f = 0; f < 4; ++f)
#pragma unroll
      for (int g = 0; g < 4; ++g) acc[f][g][0] = acc[f][g][1] = 0.0;
  }

  // A: 4 chunks of 16 B per thread and stage: k-row kr = tid/64 + 4*it, m offset mc = (tid%64)*2
  const int a_mc = (tid & 63) * 2;
  const int a_kr = tid >> 6;
  const int a_gm = m0 + a_mc;
  const int a_bytes = (a_gm < m) ? ((m - a_gm >= 2) ? 16 : 8) : 0;
  const double* a_src = A + (int64_t)a_kr * lda + (a_bytes ? a_gm : 0);
  const int64_t a_step4 = 4 * lda;            // next chunk of the same stage
  const int64_t a_stepk = (int64_t)BK * lda;  // next k-tile
  const int a_dst = a_kr * LDA_S + a_mc;
  // B: 2 chunks per thread and stage: n-row nr = tid/8 + 32*it, k offset kc = (tid%8)*2
  const int b_kc = (tid & 7) * 2;
  const int b_nr = tid >> 3;
  const int b_bytes0 = (n0 + b_nr < n) ? 16 : 0;
  const int b_bytes1 = (n0 + b_nr + 32 < n) ? 16 : 0;
  const double* b_src0 = B + (b_bytes0 ? (int64_t)(n0 + b_nr) * ldb : 0) + b_kc;
  const double* b_src1 = B + (b_bytes1 ? (int64_t)(n0 + b_nr + 32) * ldb : 0) + b_kc;
  const int b_dst = b_nr * LDB_S + b_kc;

  auto copy_a = [&](int stage, int kt, int c0) {  // chunks c0, c0+1
    double* as = As + stage * A_STAGE + a_dst;
    const double* src = a_src + (int64_t)kt * a_stepk;
    cp_async16(as + (c0 * 4) * LDA_S, src + c0 * a_step4, a_bytes);
    cp_async16(as + (c0 * 4 + 4) * LDA_S, src + (c0 + 1) * a_step4, a_bytes);
  };
  auto copy_b = [&](int stage, int kt) {
    double* bs = Bs + stage * B_STAGE + b_dst;
    cp_async16(bs, b_src0 + kt * BK, b_bytes0);
    cp_async16(bs + 32 * LDB_S, b_src1 + kt * BK, b_bytes1);
  };

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < KT) {
      copy_a(s, s, 0);
      copy_a(s, s, 2);
      copy_b(s, s);
    }
    cp_async_commit();
  }
  if (SUBTRACT && prefetch_dist > 0) {
    // CTAs are dispatched in linear order, so the CTA that will take over this SM slot when this one
    // retires is about `prefetch_dist` (= resident CTAs of the grid) tiles ahead: pull its C tile
    // (64 columns x 8 lines of 128 B) into L2 now, so that its accumulator loads do not pay DRAM latency
    const int64_t nxt = (int64_t)blockIdx.x + (int64_t)blockIdx.y * gridDim.x + prefetch_dist;
    if (nxt < (int64_t)gridDim.x * gridDim.y) {
      const int pm0 = (int)(nxt % gridDim.x) * BM, pn0 = (int)(nxt / gridDim.x) * BN;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int line = tid + q * THREADS;
        const int col = pn0 + (line >> 3), row = pm0 + (line & 7) * 16;
        if (col < n && row < m) asm volatile("prefetch.global.L2 [%0];" ::"l"(C + (int64_t)col * ldc + row));
      }
    }
  }
  if (SUBTRACT) {
#pragma unroll
    for (int f = 0; f < 4; ++f)
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        acc[f][g][0] = flip_sign(acc[f][g][0]);
        acc[f][g][1] = flip_sign(acc[f][g][1]);
      }
  }

  const int a_row = wm * 32 + (lane >> 2);
  const int b_row = wn * 32 + (lane >> 2);
  const int kq = lane & 3;
  int cs = 0, ls = STAGES - 1;  // consumer / producer stage

  for (int kt = 0; kt < KT; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    const int nxt = kt + STAGES - 1;
    const bool do_load = nxt < KT;
    const double* as = As + cs * A_STAGE;
    const double* bs = Bs + cs * B_STAGE;
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      double af[4], bf[4];
#pragma unroll
      for (int f = 0; f < 4; ++f) af[f] = as[(kk + kq) * LDA_S + a_row + 8 * f];
#pragma unroll
      for (int g = 0; g < 4; ++g) bf[g] = bs[(b_row + 8 * g) * LDB_S + kk + kq];
#pragma unroll
      for (int f = 0; f < 4; ++f)
#pragma unroll
        for (int g = 0; g < 4; ++g) dmma884(acc[f][g][0], acc[f][g][1], af[f], bf[g]);
      // next stage's copies, spread over the DMMA groups
      if (do_load) {
        if (kk == 0) copy_a(ls, nxt, 0);
        if (kk == 4) copy_a(ls, nxt, 2);
        if (kk == 8) copy_b(ls, nxt);
      }
    }
    cp_async_commit();
    cs = (cs + 1 == STAGES) ? 0 : cs + 1;
    ls = (ls + 1 == STAGES) ? 0 : ls + 1;
  }
  cp_async_wait<0>();
  if (lu_poisoned(failword, dep)) return;
  bool viol = false;

#pragma unroll
  for (int f = 0; f < 4; ++f) {
    const int gm = m0 + wm * 32 + 8 * f + (lane >> 2);
    if (gm >= m) continue;
#pragma unroll
    for (int g = 0; g < 4; ++g)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int gn = n0 + wn * 32 + 8 * g + 2 * (lane & 3) + e;
        if (gn < n) {
          double* cp = C + (int64_t)gn * ldc + gm;
          if (SUBTRACT) {
            *cp = flip_sign(acc[f][g][e]);
          } else {
            double v = alpha * acc[f][g][e];
            if (beta != 0.0) v += beta * (*cp);
            *cp = v;
            if (!(fabs(v) <= 1.0)) viol = true;
          }
        }
      }
  }
  if (check_val != 0 && viol) atomicCAS(fail, 0, check_val);
}

}  // namespace mmr_gemm

// host-side launcher (defined in lu.cu)
int mmr_dgemm_launch(mmr_ctx* ctx, cudaStream_t s, int64_t m, int64_t n, int64_t k, double alpha,
                     const double* A, int64_t lda, const double* B, int64_t ldb, double beta,
                     double* C, int64_t ldc, int* fail = nullptr, int dep = 0, int check_val = 0,
                     const mmr_gemm::GemmAlt* alt = nullptr);
