// U12 rows of a two-panel outer block in ONE launch:   B (ob x ncols, in place)  <-  L_blk^-1 B,
//
//     L_blk = [ L00  0  ]      T0 = L00^-1 B0,   B1' = B1 - L10 T0,   T1 = L11^-1 B1',     B <- [T0; T1]
//             [ L10 L11 ]
//
// with the explicit inverses L00^-1 (nb0 x nb0) and L11^-1 (nb1 x nb1) the panel kernels produce anyway.  The LU
// issued these as three dependent launches (DMMA GEMM, GEMM, GEMM); on the narrow look-ahead update (the next
// block's 240 / 256 columns) each of them is a handful of CTAs whose cost is launch + pipeline latency, and the
// three sit on the per-block chain that bounds the multi-GPU factorization and in front of every big
// trailing GEMM of the single-GPU one.  Here a CTA keeps its 32 columns of B in shared memory in the layout the
// DMMA B-fragments want ([column][row], row stride 260 doubles: conflict-free), streams the three left factors
// through the cp.async stages of dgemm.cuh and writes each result back into the tile before the next product
// reads it.  FP64 DMMA.8x8x4, 8 warps = 4 (m) x 2 (n), warp tile 32 x 16.
#pragma once

#include "dgemm.cuh"

namespace mmr_gemm {

constexpr int BT_COLS = 32;                 // columns of B per CTA
constexpr int BT_LD = 2 * BM + 4;           // doubles per staged column (ob <= 256 rows)
constexpr size_t BT_SMEM_BYTES = ((size_t)BT_COLS * BT_LD + (size_t)STAGES * A_STAGE) * sizeof(double);

// requires: nb0 == BM, 0 < nb1 <= BM, nb1 % BK == 0; Linv0 (ld nb0), Linv1 (ld nb1), L10 (ld ldl), B (ld ldb) 16-byte
// aligned with even leading dimensions.  grid = ceil(ncols / BT_COLS), block = THREADS, smem = BT_SMEM_BYTES.
__global__ void __launch_bounds__(THREADS, 1)
    lu_block_trsm_kernel(int nb0, int nb1, const double* __restrict__ Linv0, const double* __restrict__ Linv1,
                         const double* __restrict__ L10, int64_t ldl, double* __restrict__ B, int64_t ldb, int ncols,
                         const int* __restrict__ fail, int dep) {
  extern __shared__ __align__(16) double smem[];
  double* Bt = smem;                              // [BT_COLS][BT_LD]
  double* As = smem + BT_COLS * BT_LD;            // STAGES x A_STAGE
  if (lu_poisoned(lu_failword(fail), dep)) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp & 3, wn = warp >> 2;
  const int rr = lane >> 2, kq = lane & 3;
  const int c0 = blockIdx.x * BT_COLS;
  const int ob = nb0 + nb1;

  // the B tile, column by column (16-byte chunks)
  {
    const int half = ob / 2;
    for (int idx = tid; idx < BT_COLS * half; idx += THREADS) {
      const int col = idx / half, r2 = (idx - col * half) * 2;
      const bool ok = c0 + col < ncols;
      cp_async16(Bt + col * BT_LD + r2, ok ? (B + (int64_t)(c0 + col) * ldb + r2) : B, ok ? 16 : 0);
    }
    cp_async_commit();
  }

  double acc[4][2][2];
  // acc = Ap (M x K, column-major, ld lda) * Bt[:, koff : koff + K]^T   (rows >= M of the result are zero)
  auto product = [&](const double* __restrict__ Ap, int64_t lda, int M, int K, int koff) {
#pragma unroll
    for (int f = 0; f < 4; ++f)
#pragma unroll
      for (int g = 0; g < 2; ++g) acc[f][g][0] = acc[f][g][1] = 0.0;
    const int KT = K / BK;
    const int a_mc = (tid & 63) * 2, a_kr = tid >> 6;
    const int a_bytes = (a_mc < M) ? 16 : 0;
    const double* a_src = Ap + (int64_t)a_kr * lda + (a_bytes ? a_mc : 0);
    auto copy_a = [&](int stage, int kt) {
      double* as = As + stage * A_STAGE + a_kr * LDA_S + a_mc;
      const double* src = a_src + (int64_t)kt * BK * lda;
#pragma unroll
      for (int c = 0; c < 4; ++c) cp_async16(as + (4 * c) * LDA_S, src + (int64_t)(4 * c) * lda, a_bytes);
    };
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
      if (s < KT) copy_a(s, s);
      cp_async_commit();
    }
    int cs = 0, ls = STAGES - 1;
    for (int kt = 0; kt < KT; ++kt) {
      cp_async_wait<STAGES - 2>();
      __syncthreads();
      if (kt + STAGES - 1 < KT) copy_a(ls, kt + STAGES - 1);
      cp_async_commit();
      const double* as = As + cs * A_STAGE + wm * 32 + rr;
      const double* bs = Bt + (wn * 16 + rr) * BT_LD + koff + kt * BK + kq;
#pragma unroll
      for (int kk = 0; kk < BK; kk += 4) {
        double af[4], bf[2];
#pragma unroll
        for (int f = 0; f < 4; ++f) af[f] = as[(kk + kq) * LDA_S + 8 * f];
#pragma unroll
        for (int g = 0; g < 2; ++g) bf[g] = bs[(8 * g) * BT_LD + kk];
#pragma unroll
        for (int f = 0; f < 4; ++f)
#pragma unroll
          for (int g = 0; g < 2; ++g) dmma884(acc[f][g][0], acc[f][g][1], af[f], bf[g]);
      }
      cs = (cs + 1 == STAGES) ? 0 : cs + 1;
      ls = (ls + 1 == STAGES) ? 0 : ls + 1;
    }
    cp_async_wait<0>();
    __syncthreads();  // every warp is done with Bt[:, koff : koff + K] and with the stages
  };
  // result rows [0, M) of acc -> Bt[:, roff + row]  (SUB: Bt <- Bt - acc)
  auto put = [&](int M, int roff, bool sub) {
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      const int mrow = wm * 32 + 8 * f + rr;
      if (mrow >= M) continue;
#pragma unroll
      for (int g = 0; g < 2; ++g)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          double* p = Bt + (wn * 16 + 8 * g + 2 * kq + e) * BT_LD + roff + mrow;
          *p = sub ? (*p - acc[f][g][e]) : acc[f][g][e];
        }
    }
    __syncthreads();
  };

  product(Linv0, nb0, nb0, nb0, 0);  // (its first wait also covers the B tile: one more group in flight -> see below)
  put(nb0, 0, false);                // T0
  product(L10, ldl, nb1, nb0, 0);    // L10 T0
  put(nb1, nb0, true);               // B1' = B1 - L10 T0
  product(Linv1, nb1, nb1, nb1, nb0);
  put(nb1, nb0, false);              // T1

  {
    const int half = ob / 2;
    for (int idx = tid; idx < BT_COLS * half; idx += THREADS) {
      const int col = idx / half, r2 = (idx - col * half) * 2;
      if (c0 + col < ncols) *(double2*)(B + (int64_t)(c0 + col) * ldb + r2) = *(const double2*)(Bt + col * BT_LD + r2);
    }
  }
}

}  // namespace mmr_gemm
