// Diagonal block of a speculative LU panel in ONE cluster launch: no-pivot LU of the nb x nb (<= 128) block
// and the inverses of both triangular factors (what the panel's L21 = A21 U11^-1 product and the trailing
// U12 = L11^-1 A12 product consume).  Replaces the round-1 pair lu_diag_nopiv_tile_kernel (one CTA, 128
// CTA-wide barrier steps, ~50 us) + lu_tri_inv_kernel (8 CTAs, ~48 us): that serial pair was the largest
// item of the per-panel chain that limits the multi-GPU LU at small N ("what's weak" 5).
//
//   phase A (cluster rank 0): the block sits in shared memory; blocked right-looking LU with 32-wide steps:
//            (1) one warp factors the 32 x 32 diagonal sub-block in registers (shuffles, no barrier),
//            (2) the row block U12 = L_D^-1 A12 and the column block L21 = A21 U_D^-1 are substitutions with
//                one thread per column / row,
//            (3) the Schur complement is updated with DMMA (mma.sync.m8n8k4.f64) tiles on all 16 warps.
//            3 CTA barriers per 32 columns instead of one per column.
//   phase B (all 8 ranks): rank r < 4 computes columns [32r, 32r+32) of L^-1, rank r >= 4 the same columns of
//            U^-1, by blocked substitution on the identity (32-row diagonal solves + DMMA updates), from a
//            copy of the factored block fetched through distributed shared memory.
//
// Speculation check: a zero pivot or a multiplier with |l| > 1 (partial pivoting would have interchanged)
// sets *fail = kb + 1 (sticky); everything downstream is guarded by lu_poisoned.
#pragma once

#include <cooperative_groups.h>

#include "dgemm.cuh"

namespace mmr_panel {

namespace cg = cooperative_groups;

constexpr int PB = 128;        // block order
constexpr int SUB = 32;        // sub-block of the blocked algorithms
constexpr int LDM = PB + 4;    // shared-memory leading dimension: DMMA fragment loads are conflict free
constexpr int PTHREADS = 512;
constexpr int PCLUSTER = 8;
constexpr size_t PANEL_SMEM = ((size_t)PB * LDM + (size_t)SUB * LDM + PB) * sizeof(double);

__device__ __forceinline__ double rcp_newton(double a) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
  x = fma(x, fma(-a, x, 1.0), x);
  x = fma(x, fma(-a, x, 1.0), x);
  return x;
}

// C[8 x 8] -= A[8 x kdim] * B[kdim x 8], all column-major in shared memory: C(r, c) = C[c*ldc + r],
// A(r, k) = A[k*lda + r], B(k, c) = B[c*ldb + k].  One warp, kdim a multiple of 4.
__device__ __forceinline__ void tile_sub(double* C, int ldc, const double* A, int lda, const double* B, int ldb,
                                         int kdim, int lane) {
  const int row = lane >> 2, kq = lane & 3;
  double* c = C + (2 * kq) * ldc + row;
  double c0 = c[0], c1 = c[ldc];
  const double* a = A + kq * lda + row;
  const double* b = B + row * ldb + kq;
#pragma unroll 4
  for (int kk = 0; kk < kdim; kk += 4) {
    const double av = -a[kk * lda];
    const double bv = b[kk];
    mmr_gemm::dmma884(c0, c1, av, bv);
  }
  c[0] = c0;
  c[ldc] = c1;
}

__global__ void __cluster_dims__(PCLUSTER, 1, 1) __launch_bounds__(PTHREADS, 1)
    lu_diag_inv_cluster_kernel(const double* __restrict__ A, int64_t ld, int kb, int nb, double* __restrict__ S,
                               int64_t lds, double* __restrict__ Linv, double* __restrict__ Uinv,
                               int* __restrict__ fail) {
  extern __shared__ __align__(16) double psm[];
  double* M = psm;                // [PB][LDM]  M[c*LDM + r]
  double* X = psm + PB * LDM;     // [SUB][LDM] X[cc*LDM + r]   (phase B)
  double* rdiag = X + SUB * LDM;  // [PB] reciprocals of U's diagonal
  __shared__ int s_bad;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // an earlier panel was rejected: nothing to do (the word cannot change while this kernel runs -- only a
  // panel that precedes it in stream order could set a value <= kb)
  if (mmr_gemm::lu_poisoned(mmr_gemm::lu_failword(fail), kb - 1)) return;

  if (rank == 0) {
    // ---- phase A: blocked no-pivot LU in shared memory -----------------------------------------------
    for (int e = tid; e < PB * PB; e += PTHREADS) {
      const int r = e % PB, c = e / PB;
      M[c * LDM + r] = (r < nb && c < nb) ? A[(int64_t)(kb + c) * ld + kb + r] : ((r == c) ? 1.0 : 0.0);
    }
    if (tid == 0) s_bad = 0;
    __syncthreads();
    bool bad = false;
    for (int r0 = 0; r0 < PB; r0 += SUB) {
      // (1) diagonal sub-block: lane i holds row r0 + i
      if (warp == 0) {
        double a[SUB];
#pragma unroll
        for (int c = 0; c < SUB; ++c) a[c] = M[(r0 + c) * LDM + r0 + lane];
#pragma unroll
        for (int j = 0; j < SUB; ++j) {
          const double pj = __shfl_sync(0xffffffffu, a[j], j);
          if (pj == 0.0) bad = true;
          const double rinv = rcp_newton(pj);
          if (lane == j) rdiag[r0 + j] = rinv;
          double l = 0.0;
          if (lane > j) {
            l = a[j] * rinv;
            if (!(fabs(l) <= 1.0)) bad = true;
            a[j] = l;
          }
#pragma unroll
          for (int c = j + 1; c < SUB; ++c) {
            const double u = __shfl_sync(0xffffffffu, a[c], j);
            a[c] = fma(-l, u, a[c]);
          }
        }
#pragma unroll
        for (int c = 0; c < SUB; ++c) M[(r0 + c) * LDM + r0 + lane] = a[c];
      }
      __syncthreads();
      const int rest = PB - r0 - SUB;  // rows / columns beyond this sub-block
      if (rest > 0) {
        // (2) U12 = L_D^-1 A12 (threads [0, rest): one column each) and L21 = A21 U_D^-1 (threads
        //     [256, 256 + rest): one row each)
        if (tid < rest) {
          double* col = M + (r0 + SUB + tid) * LDM + r0;
          double x[SUB];
#pragma unroll
          for (int i = 0; i < SUB; ++i) x[i] = col[i];
#pragma unroll
          for (int q = 0; q < SUB - 1; ++q) {
            const double* lq = M + (r0 + q) * LDM + r0;  // column q of L_D (broadcast reads)
#pragma unroll
            for (int i = q + 1; i < SUB; ++i) x[i] = fma(-lq[i], x[q], x[i]);
          }
#pragma unroll
          for (int i = 0; i < SUB; ++i) col[i] = x[i];
        } else if (tid >= 256 && tid < 256 + rest) {
          double* rowp = M + r0 * LDM + r0 + SUB + (tid - 256);  // element q of the row at rowp[q*LDM]
          double y[SUB];
#pragma unroll
          for (int q = 0; q < SUB; ++q) y[q] = rowp[q * LDM];
#pragma unroll
          for (int q = 0; q < SUB; ++q) {
            const double* uq = M + (r0 + q) * LDM + r0;  // column q of U_D: uq[p], p <= q
#pragma unroll
            for (int p = 0; p < q; ++p) y[q] = fma(-y[p], uq[p], y[q]);
            y[q] *= rdiag[r0 + q];
            if (!(fabs(y[q]) <= 1.0)) bad = true;
          }
#pragma unroll
          for (int q = 0; q < SUB; ++q) rowp[q * LDM] = y[q];
        }
        __syncthreads();
        // (3) Schur complement: M[r0+32:, r0+32:] -= L21 U12, 8 x 8 DMMA tiles dealt to the 16 warps
        const int nt = rest / 8;
        for (int t = warp; t < nt * nt; t += PTHREADS / 32) {
          const int tr = (t % nt) * 8, tc = (t / nt) * 8;
          tile_sub(M + (r0 + SUB + tc) * LDM + r0 + SUB + tr, LDM, M + r0 * LDM + r0 + SUB + tr, LDM,
                   M + (r0 + SUB + tc) * LDM + r0, LDM, SUB, lane);
        }
        __syncthreads();
      }
    }
    if (bad) s_bad = 1;
    __syncthreads();
    if (tid == 0 && s_bad) atomicCAS(fail, 0, kb + 1);
    for (int e = tid; e < nb * nb; e += PTHREADS) {
      const int r = e % nb, c = e / nb;
      S[(int64_t)c * lds + r] = M[c * LDM + r];
    }
  }
  cluster.sync();  // rank 0's shared memory (M, rdiag, s_bad) is complete and visible cluster-wide

  // ---- phase B: inverses of the two factors, 32 columns per CTA --------------------------------------
  {
    const int* bad0 = cluster.map_shared_rank(&s_bad, 0);
    const bool rejected = (*bad0 != 0);
    if (!rejected) {
      if (rank != 0) {
        const double* M0 = cluster.map_shared_rank(M, 0);
        const double* rd0 = cluster.map_shared_rank(rdiag, 0);
        for (int e = tid; e < PB * PB; e += PTHREADS) {
          const int r = e % PB, c = e / PB;
          M[c * LDM + r] = M0[c * LDM + r];
        }
        if (tid < PB) rdiag[tid] = rd0[tid];
      }
      const bool upper = rank >= 4;
      const int cb = (rank & 3) * SUB;
      for (int e = tid; e < SUB * PB; e += PTHREADS) {
        const int r = e % PB, cc = e / PB;
        X[cc * LDM + r] = (r == cb + cc) ? 1.0 : 0.0;
      }
      __syncthreads();
      if (!upper) {
        // L X = E: rows above cb stay zero; walk the 32-row blocks downwards
        for (int rb = cb; rb < PB; rb += SUB) {
          if (tid < SUB) {
            double* col = X + tid * LDM + rb;
            double x[SUB];
#pragma unroll
            for (int i = 0; i < SUB; ++i) x[i] = col[i];
#pragma unroll
            for (int q = 0; q < SUB - 1; ++q) {
              const double* lq = M + (rb + q) * LDM + rb;
#pragma unroll
              for (int i = q + 1; i < SUB; ++i) x[i] = fma(-lq[i], x[q], x[i]);
            }
#pragma unroll
            for (int i = 0; i < SUB; ++i) col[i] = x[i];
          }
          __syncthreads();
          const int below = PB - rb - SUB;
          if (below > 0) {
            const int ntr = below / 8;
            for (int t = warp; t < ntr * (SUB / 8); t += PTHREADS / 32) {
              const int tr = (t % ntr) * 8, tc = (t / ntr) * 8;
              tile_sub(X + tc * LDM + rb + SUB + tr, LDM, M + rb * LDM + rb + SUB + tr, LDM, X + tc * LDM + rb, LDM,
                       SUB, lane);
            }
            __syncthreads();
          }
        }
        for (int e = tid; e < SUB * PB; e += PTHREADS) {
          const int r = e % PB, cc = e / PB;
          if (r < nb && cb + cc < nb) Linv[(int64_t)(cb + cc) * nb + r] = X[cc * LDM + r];
        }
      } else {
        // U X = E: rows below cb + 31 stay zero; walk the 32-row blocks upwards
        for (int rb = cb; rb >= 0; rb -= SUB) {
          if (tid < SUB) {
            double* col = X + tid * LDM + rb;
            double x[SUB];
#pragma unroll
            for (int i = 0; i < SUB; ++i) x[i] = col[i];
#pragma unroll
            for (int q = SUB - 1; q >= 0; --q) {
              const double* uq = M + (rb + q) * LDM + rb;  // column q of U_D: uq[i], i <= q
              x[q] *= rdiag[rb + q];
#pragma unroll
              for (int i = 0; i < q; ++i) x[i] = fma(-uq[i], x[q], x[i]);
            }
#pragma unroll
            for (int i = 0; i < SUB; ++i) col[i] = x[i];
          }
          __syncthreads();
          if (rb > 0) {
            const int ntr = rb / 8;
            for (int t = warp; t < ntr * (SUB / 8); t += PTHREADS / 32) {
              const int tr = (t % ntr) * 8, tc = (t / ntr) * 8;
              tile_sub(X + tc * LDM + tr, LDM, M + rb * LDM + tr, LDM, X + tc * LDM + rb, LDM, SUB, lane);
            }
            __syncthreads();
          }
        }
        for (int e = tid; e < SUB * PB; e += PTHREADS) {
          const int r = e % PB, cc = e / PB;
          if (r < nb && cb + cc < nb) Uinv[(int64_t)(cb + cc) * nb + r] = X[cc * LDM + r];
        }
      }
    }
  }
  cluster.sync();  // nobody leaves while its shared memory can still be read by a peer
}

}  // namespace mmr_panel
