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
//            copy of the factored block read back through L2.
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

// One warp: a strip of `nt` (<= MAXT) 8 x 8 tiles in one tile row, C[8 x 8 nt] -= A[8 x kdim] * B[kdim x 8 nt];
// the A fragment is loaded once per k step and the nt accumulator pairs are independent DMMA chains.
template <int MAXT>
__device__ __forceinline__ void strip_sub(double* C, int ldc, const double* A, int lda, const double* B, int ldb,
                                          int nt, int kdim, int lane) {
  const int row = lane >> 2, kq = lane & 3;
  double c0[MAXT], c1[MAXT];
  double* c = C + (2 * kq) * ldc + row;
#pragma unroll
  for (int t = 0; t < MAXT; ++t) {
    c0[t] = (t < nt) ? c[(8 * t) * ldc] : 0.0;
    c1[t] = (t < nt) ? c[(8 * t + 1) * ldc] : 0.0;
  }
  const double* a = A + kq * lda + row;
  const double* b = B + row * ldb + kq;
  for (int kk = 0; kk < kdim; kk += 4) {
    const double av = -a[kk * lda];
#pragma unroll
    for (int t = 0; t < MAXT; ++t) {
      if (t < nt) {
        const double bv = b[(8 * t) * ldb + kk];
        mmr_gemm::dmma884(c0[t], c1[t], av, bv);
      }
    }
  }
#pragma unroll
  for (int t = 0; t < MAXT; ++t) {
    if (t < nt) {
      c[(8 * t) * ldc] = c0[t];
      c[(8 * t + 1) * ldc] = c1[t];
    }
  }
}

// One elimination step of the 32 x 32 diagonal sub-block (see the kernel): J is a template parameter so that
// every register index is a compile-time constant (nvcc does not fully unroll a 32-trip loop with this body,
// and a dynamically indexed row lands in local memory).
template <int J>
__device__ __forceinline__ void diag_step(double (&a)[SUB], double (&prow)[2][SUB + 2], double* __restrict__ rd,
                                          int lane, bool& bad) {
  const double* pr = prow[J & 1];
  double* pn = prow[(J + 1) & 1];
  const double rinv = pr[SUB];
  double l = 0.0;
  if (lane > J) {
    l = a[J] * rinv;
    if (!(fabs(l) <= 1.0)) bad = true;
    a[J] = l;
  }
  const bool next = (lane == J + 1);
  double rn = 0.0;
#pragma unroll
  for (int c = J + 1; c < SUB; ++c) {
    a[c] = fma(-l, pr[c], a[c]);
    if (c == J + 1 && next) {
      if (a[c] == 0.0) bad = true;
      rn = rcp_newton(a[c]);
    }
    if (next && c > J + 1) pn[c] = a[c];
  }
  if (J + 1 < SUB && next) {
    pn[SUB] = rn;
    rd[J + 1] = rn;
  }
  __syncwarp();
}
template <int J>
struct DiagSteps {
  static __device__ __forceinline__ void run(double (&a)[SUB], double (&prow)[2][SUB + 2], double* __restrict__ rd,
                                             int lane, bool& bad) {
    diag_step<J>(a, prow, rd, lane, bad);
    DiagSteps<J + 1>::run(a, prow, rd, lane, bad);
  }
};
template <>
struct DiagSteps<SUB> {
  static __device__ __forceinline__ void run(double (&)[SUB], double (&)[2][SUB + 2], double* __restrict__, int, bool&) {}
};

__global__ void __cluster_dims__(PCLUSTER, 1, 1) __launch_bounds__(PTHREADS, 1)
    lu_diag_inv_cluster_kernel(const double* __restrict__ A, int64_t ld, int kb, int nb, double* __restrict__ S,
                               int64_t lds, double* __restrict__ Linv, double* __restrict__ Uinv,
                               int* __restrict__ fail, long long* __restrict__ clk) {
  extern __shared__ __align__(16) double psm[];
  double* M = psm;                // [PB][LDM]  M[c*LDM + r]
  double* X = psm + PB * LDM;     // [SUB][LDM] X[cc*LDM + r]   (phase B)
  double* rdiag = X + SUB * LDM;  // [PB] reciprocals of U's diagonal
  __shared__ int s_bad;
  __shared__ __align__(16) double prow[2][SUB + 2];  // pivot row of the current elimination step + 1 / pivot
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // an earlier panel was rejected: nothing to do (the word cannot change while this kernel runs -- only a
  // panel that precedes it in stream order could set a value <= kb)
  if (mmr_gemm::lu_poisoned(mmr_gemm::lu_failword(fail), kb - 1)) return;
  int nclk = 0;
#define PANEL_STAMP()                                                  \
  do {                                                                 \
    if (clk && tid == 0 && rank == 0 && nclk < 30) clk[nclk] = clock64(); \
    ++nclk;                                                            \
  } while (0)
  PANEL_STAMP();

  if (rank == 0) {
    // ---- phase A: blocked no-pivot LU in shared memory -----------------------------------------------
    {
      const int r = tid & (PB - 1), cq = tid >> 7;  // 4 columns in flight per pass, 8 passes per batch
      const double* src = A + (int64_t)kb * ld + kb + r;
#pragma unroll 8
      for (int c = cq; c < PB; c += PTHREADS / PB)
        M[c * LDM + r] = (r < nb && c < nb) ? __ldg(src + (int64_t)c * ld) : ((r == c) ? 1.0 : 0.0);
    }
    if (tid == 0) s_bad = 0;
    __syncthreads();
    PANEL_STAMP();  // 1: block loaded
    bool bad = false;
    for (int r0 = 0; r0 < PB; r0 += SUB) {
      // (1) diagonal sub-block: lane i holds row r0 + i in registers.  The pivot row of a step (and the
      //     reciprocal of its pivot) reaches the other lanes through a double-buffered shared-memory row: its
      //     owner stores every entry right after the update that finalises it, and computes the reciprocal of
      //     the next pivot as soon as that one value is known, so the MUFU + Newton latency overlaps the rest
      //     of the step.  (A first version broadcast the row with shuffles: the in-order SHFL -> DFMA pairs
      //     cost 540 cycles per step.)
      if (warp == 0) {
        double a[SUB];
#pragma unroll
        for (int c = 0; c < SUB; ++c) a[c] = M[(r0 + c) * LDM + r0 + lane];
        if (lane == 0) {
          if (a[0] == 0.0) bad = true;
          const double r = rcp_newton(a[0]);
          prow[0][SUB] = r;
          rdiag[r0] = r;
#pragma unroll
          for (int c = 1; c < SUB; ++c) prow[0][c] = a[c];
        }
        __syncwarp();
        DiagSteps<0>::run(a, prow, rdiag + r0, lane, bad);
#pragma unroll
        for (int c = 0; c < SUB; ++c) M[(r0 + c) * LDM + r0 + lane] = a[c];
      }
      __syncthreads();
      PANEL_STAMP();  // 2 + 3b: diagonal sub-block done
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
        PANEL_STAMP();  // 3 + 3b: substitutions done
        // (3) Schur complement: M[r0+32:, r0+32:] -= L21 U12, 8 x 8 DMMA tiles dealt to the 16 warps
        const int nt = rest / 8;
        // strips of up to 6 tiles: nt tile rows x ceil(nt / 6) strips are dealt to the 16 warps
        const int nsx = (nt + 5) / 6;
        for (int t = warp; t < nt * nsx; t += PTHREADS / 32) {
          const int tr = (t % nt) * 8, ts = (t / nt) * 6;
          strip_sub<6>(M + (r0 + SUB + 8 * ts) * LDM + r0 + SUB + tr, LDM, M + r0 * LDM + r0 + SUB + tr, LDM,
                       M + (r0 + SUB + 8 * ts) * LDM + r0, LDM, min(6, nt - ts), SUB, lane);
        }
        __syncthreads();
        PANEL_STAMP();  // 4 + 3b: Schur update done
      }
    }
    if (bad) s_bad = 1;
    __syncthreads();
    if (tid == 0 && s_bad) atomicCAS(fail, 0, kb + 1);
    {
      const int r = tid & (PB - 1);
      if (r < nb)
        for (int c = tid >> 7; c < nb; c += PTHREADS / PB) S[(int64_t)c * lds + r] = M[c * LDM + r];
    }
  } else {
    nclk = 12;
  }
  PANEL_STAMP();  // 12: factors written
  cluster.sync();  // rank 0's writes (S, rdiag column, guard word) are visible cluster-wide
  PANEL_STAMP();  // 13

  // ---- phase B: inverses of the two factors, 32 columns per CTA --------------------------------------
  {
    const bool rejected = mmr_gemm::lu_poisoned(mmr_gemm::lu_failword(fail), kb);
    if (!rejected) {
      const bool upper = rank >= 4;
      const int cb = (rank & 3) * SUB;
      if (rank != 0) {
        // the factored block comes back through L2 (distributed shared memory moves ~20 B/clk per SM: seven
        // peers pulling 128 KB each from rank 0 cost more than the whole factorization); only the part this
        // CTA needs: L columns >= cb (rows >= cb) or U columns <= cb + 31 (rows <= cb + 31)
        const int c_lo = upper ? 0 : cb, c_hi = upper ? cb + SUB : PB;
        const int ncol = c_hi - c_lo;
        for (int e = tid; e < ncol * PB; e += PTHREADS) {
          const int r = e % PB, c = c_lo + e / PB;
          M[c * LDM + r] = (r < nb && c < nb) ? __ldcg(S + (int64_t)c * lds + r) : ((r == c) ? 1.0 : 0.0);
        }
        __syncthreads();
        if (upper && tid < PB) rdiag[tid] = rcp_newton(M[tid * LDM + tid]);  // the same values rank 0 used
      }
      for (int e = tid; e < SUB * PB; e += PTHREADS) {
        const int r = e % PB, cc = e / PB;
        X[cc * LDM + r] = (r == cb + cc) ? 1.0 : 0.0;
      }
      __syncthreads();
      PANEL_STAMP();  // 14: factors fetched
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
            for (int t = warp; t < ntr; t += PTHREADS / 32)
              strip_sub<4>(X + rb + SUB + 8 * t, LDM, M + rb * LDM + rb + SUB + 8 * t, LDM, X + rb, LDM, SUB / 8, SUB, lane);
            __syncthreads();
          }
        }
        for (int e = tid; e < SUB * PB; e += PTHREADS) {
          const int r = e % PB, cc = e / PB;
          if (r < nb && cb + cc < nb) Linv[(int64_t)(cb + cc) * nb + r] = X[cc * LDM + r];
        }
        PANEL_STAMP();  // 15: L^-1 columns [0, 32) done (rank 0: the longest substitution)
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
            for (int t = warp; t < ntr; t += PTHREADS / 32)
              strip_sub<4>(X + 8 * t, LDM, M + rb * LDM + 8 * t, LDM, X + rb, LDM, SUB / 8, SUB, lane);
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
  cluster.sync();  // the cluster retires together
  nclk = 16;
  PANEL_STAMP();  // 16: end
#undef PANEL_STAMP
}

}  // namespace mmr_panel
