// Kernel 2a: dense LU with partial pivoting + triangular solves -- replaces
// np.linalg.solve(Q, rhs) (LAPACK dgesv) of demag.py:469-470.
//
// Storage trick: Q is row-major (rows = target cells, the multi-GPU sharding unit).  Its memory
// is, verbatim, the column-major matrix A = Q^T.  We run a textbook right-looking blocked LU with
// partial (row) pivoting on that column-major view, A = P L U, and solve Q x = b through
// Q = A^T = U^T L^T P^T.  Every pivot search then stays inside one column panel of A = one
// row block of Q = one GPU, which is what the 1-D block-cyclic distribution needs.
//
// Per panel of NB columns:
//   strips of W columns:  cooperative "strip" kernel (panel rows spread over all SMs in shared
//                         memory, one grid barrier per column for the pivot search)
//                         -> swaps + W x W triangular solve on the rest of the panel
//                         -> rank-W DMMA update of the rest of the panel
//   row swaps on the columns outside the panel, TRSM (U12 = L11^-1 A12), DMMA trailing update.
#include <cooperative_groups.h>

#include <cstdlib>

#include "comm.h"
#include "dgemm.cuh"
#include "dgemm_tma.cuh"
#include "block_trsm.cuh"
#include "panel128.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int NB = 128;  // panel width
constexpr int W = 16;    // strip width inside the panel
constexpr int STRIP_THREADS = 256;
constexpr int MIN_ROWS_PER_CTA = 64;

struct Cand {
  double absval;
  int row;
  int pad;
};

// ---------------------------------------------------------------------------------------
// Strip factorization: columns [c0, c0+w) (w <= W), rows [c0, N) of column-major A.
// Rows are split into contiguous chunks, one per CTA, held in shared memory for the whole
// strip.  CTA 0 owns the top W rows.  One grid barrier per column.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(STRIP_THREADS)
    lu_strip_kernel(double* __restrict__ A, int64_t ld, int N, int c0, int w, int rows_per,
                    int* __restrict__ ipiv, Cand* __restrict__ cand, double* __restrict__ candrow,
                    double* __restrict__ diagrow, int* __restrict__ info) {
  extern __shared__ __align__(16) double S[];  // [W][rows_per]
  __shared__ double s_red_val[STRIP_THREADS / 32];
  __shared__ int s_red_row[STRIP_THREADS / 32];
  __shared__ double s_prow[W];
  __shared__ int s_piv_row;
  __shared__ int s_piv_cta;

  cg::grid_group grid = cg::this_grid();
  const int G = gridDim.x;
  const int cta = blockIdx.x;
  const int tid = threadIdx.x;
  const int m = N - c0;
  const int r_begin = cta * rows_per;                // local (panel-relative) first row
  const int r_cnt = max(0, min(rows_per, m - r_begin));

  // load chunk: S[c*rows_per + r] = A[(c0 + r_begin + r) + (c0 + c)*ld]
  for (int c = 0; c < w; ++c) {
    const double* src = A + (int64_t)(c0 + c) * ld + c0 + r_begin;
    for (int r = tid; r < r_cnt; r += STRIP_THREADS) S[c * rows_per + r] = src[r];
  }
  __syncthreads();

  for (int jj = 0; jj < w; ++jj) {
    const int buf = jj & 1;
    // 1. local argmax over rows with panel-relative index >= jj
    double best = -1.0;
    int best_row = 0x7fffffff;
    {
      const int lo = max(0, jj - r_begin);
      for (int r = lo + tid; r < r_cnt; r += STRIP_THREADS) {
        const double v = fabs(S[jj * rows_per + r]);
        if (v > best) {  // strict: first maximum wins inside a thread (rows ascending)
          best = v;
          best_row = r_begin + r;
        }
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, best, off);
        const int orow = __shfl_down_sync(0xffffffffu, best_row, off);
        if (ov > best || (ov == best && orow < best_row)) {
          best = ov;
          best_row = orow;
        }
      }
      if ((tid & 31) == 0) {
        s_red_val[tid >> 5] = best;
        s_red_row[tid >> 5] = best_row;
      }
      __syncthreads();
      if (tid == 0) {
        for (int q = 1; q < STRIP_THREADS / 32; ++q) {
          if (s_red_val[q] > best || (s_red_val[q] == best && s_red_row[q] < best_row)) {
            best = s_red_val[q];
            best_row = s_red_row[q];
          }
        }
        Cand cd;
        cd.absval = best;
        cd.row = best_row;
        cd.pad = 0;
        cand[buf * G + cta] = cd;
        s_piv_row = best_row;  // local candidate, reused below to publish the row
      }
      __syncthreads();
      // publish my candidate row (w values) and, from CTA 0, the current diagonal row
      if (best >= 0.0 || true) {
        const int lr = s_piv_row - r_begin;
        if (tid < w && s_piv_row != 0x7fffffff && lr >= 0 && lr < r_cnt)
          candrow[((int64_t)buf * G + cta) * W + tid] = S[tid * rows_per + lr];
      }
      if (cta == 0 && tid < w) diagrow[buf * W + tid] = S[tid * rows_per + jj];
    }
    __threadfence();
    grid.sync();

    // 2. every CTA picks the winner redundantly
    if (tid < 32) {
      double bv = -1.0;
      int br = 0x7fffffff, bc = 0;
      for (int q = tid; q < G; q += 32) {
        const Cand cd = cand[buf * G + q];
        if (cd.absval > bv || (cd.absval == bv && cd.row < br)) {
          bv = cd.absval;
          br = cd.row;
          bc = q;
        }
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, bv, off);
        const int orow = __shfl_down_sync(0xffffffffu, br, off);
        const int oc = __shfl_down_sync(0xffffffffu, bc, off);
        if (ov > bv || (ov == bv && orow < br)) {
          bv = ov;
          br = orow;
          bc = oc;
        }
      }
      if (tid == 0) {
        s_piv_row = br;
        s_piv_cta = bc;
      }
    }
    __syncthreads();
    const int prow = s_piv_row;  // panel-relative pivot row
    const int pcta = s_piv_cta;
    if (tid < w) s_prow[tid] = candrow[((int64_t)buf * G + pcta) * W + tid];
    __syncthreads();
    // 3. swap: pivot row <- old diagonal row (owner of prow); diagonal row <- pivot row (CTA 0)
    if (prow != jj) {
      if (cta == pcta && tid < w) {
        const double dv = diagrow[buf * W + tid];
        S[tid * rows_per + (prow - r_begin)] = dv;
      }
      __syncthreads();  // if pcta == 0 the two writes touch different rows; order is irrelevant
      if (cta == 0 && tid < w) S[tid * rows_per + jj] = s_prow[tid];
    }
    if (cta == 0 && tid == 0) {
      ipiv[c0 + jj] = c0 + prow;
      if (s_prow[jj] == 0.0) atomicCAS(info, 0, c0 + jj + 1);
    }
    __syncthreads();
    // 4. scale column jj and rank-1 update of the strip columns to its right
    const double piv = s_prow[jj];
    if (piv != 0.0) {
      const double rinv = 1.0 / piv;
      const int lo = max(0, jj + 1 - r_begin);
      for (int r = lo + tid; r < r_cnt; r += STRIP_THREADS) {
        const double l = S[jj * rows_per + r] * rinv;
        S[jj * rows_per + r] = l;
        for (int c = jj + 1; c < w; ++c) S[c * rows_per + r] -= l * s_prow[c];
      }
    }
    __syncthreads();
  }

  // write back
  for (int c = 0; c < w; ++c) {
    double* dst = A + (int64_t)(c0 + c) * ld + c0 + r_begin;
    for (int r = tid; r < r_cnt; r += STRIP_THREADS) dst[r] = S[c * rows_per + r];
  }
}

// ---------------------------------------------------------------------------------------
// Cluster variant of the strip factorization: ONE thread-block cluster (<= 16 CTAs) holds the
// whole strip in distributed shared memory.  The per-column pivot search uses the hardware
// cluster barrier and DSMEM reads of the other CTAs' candidates instead of a grid-wide software
// barrier through L2 (measured 4.8 us per column with the cooperative-grid kernel above).
// Used whenever rows x W doubles fit 16 x 200 KB (N <= ~25600 at W = 16; always in the late
// panels of larger matrices).
// ---------------------------------------------------------------------------------------
constexpr int CSTRIP_THREADS = 512;

struct StripRec {
  double absval;
  int row;
  int pad;
  double vals[W];
};

__global__ void __launch_bounds__(CSTRIP_THREADS)
    lu_strip_cluster_kernel(double* __restrict__ A, int64_t ld, int N, int c0, int w, int rows_per,
                            int* __restrict__ ipiv, int* __restrict__ info) {
  extern __shared__ __align__(16) double S[];  // [W][rows_per]
  __shared__ StripRec rec[2];
  __shared__ double diag[2][W];
  __shared__ double s_red_val[CSTRIP_THREADS / 32];
  __shared__ int s_red_row[CSTRIP_THREADS / 32];
  __shared__ double s_prow[W];
  __shared__ int s_piv_row;
  __shared__ int s_piv_cta;

  cg::cluster_group cluster = cg::this_cluster();
  const int G = (int)cluster.num_blocks();
  const int cta = (int)cluster.block_rank();
  const int tid = threadIdx.x;
  const int m = N - c0;
  const int r_begin = cta * rows_per;
  const int r_cnt = max(0, min(rows_per, m - r_begin));

  for (int c = 0; c < w; ++c) {
    const double* src = A + (int64_t)(c0 + c) * ld + c0 + r_begin;
    for (int r = tid; r < r_cnt; r += CSTRIP_THREADS) S[c * rows_per + r] = src[r];
  }
  __syncthreads();

  for (int jj = 0; jj < w; ++jj) {
    const int buf = jj & 1;
    double best = -1.0;
    int best_row = 0x7fffffff;
    const int lo = max(0, jj - r_begin);
    for (int r = lo + tid; r < r_cnt; r += CSTRIP_THREADS) {
      const double v = fabs(S[jj * rows_per + r]);
      if (v > best) {
        best = v;
        best_row = r_begin + r;
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double ov = __shfl_down_sync(0xffffffffu, best, off);
      const int orow = __shfl_down_sync(0xffffffffu, best_row, off);
      if (ov > best || (ov == best && orow < best_row)) {
        best = ov;
        best_row = orow;
      }
    }
    if ((tid & 31) == 0) {
      s_red_val[tid >> 5] = best;
      s_red_row[tid >> 5] = best_row;
    }
    __syncthreads();
    if (tid < 32) {
      best = (tid < CSTRIP_THREADS / 32) ? s_red_val[tid] : -1.0;
      best_row = (tid < CSTRIP_THREADS / 32) ? s_red_row[tid] : 0x7fffffff;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, best, off);
        const int orow = __shfl_down_sync(0xffffffffu, best_row, off);
        if (ov > best || (ov == best && orow < best_row)) {
          best = ov;
          best_row = orow;
        }
      }
      best = __shfl_sync(0xffffffffu, best, 0);
      best_row = __shfl_sync(0xffffffffu, best_row, 0);
      if (tid == 0) {
        rec[buf].absval = best;
        rec[buf].row = best_row;
      }
      const int lr = best_row - r_begin;
      if (tid < w && best_row != 0x7fffffff) rec[buf].vals[tid] = S[tid * rows_per + lr];
      if (cta == 0 && tid < w) diag[buf][tid] = S[tid * rows_per + jj];
    }
    cluster.sync();  // release/acquire: every CTA's record is visible cluster-wide

    if (tid < 32) {
      double bv = -1.0;
      int br = 0x7fffffff, bc = 0;
      if (tid < G) {
        const StripRec* rr = cluster.map_shared_rank(&rec[buf], tid);
        bv = rr->absval;
        br = rr->row;
        bc = tid;
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, bv, off);
        const int orow = __shfl_down_sync(0xffffffffu, br, off);
        const int oc = __shfl_down_sync(0xffffffffu, bc, off);
        if (ov > bv || (ov == bv && orow < br)) {
          bv = ov;
          br = orow;
          bc = oc;
        }
      }
      br = __shfl_sync(0xffffffffu, br, 0);
      bc = __shfl_sync(0xffffffffu, bc, 0);
      if (tid == 0) {
        s_piv_row = br;
        s_piv_cta = bc;
      }
      if (tid < w) {
        const StripRec* rr = cluster.map_shared_rank(&rec[buf], bc);
        s_prow[tid] = rr->vals[tid];
      }
    }
    __syncthreads();
    const int prow = s_piv_row;
    const int pcta = s_piv_cta;
    if (prow != jj) {
      if (cta == pcta && tid < w) {
        const double* dg = cluster.map_shared_rank(&diag[buf][0], 0);
        S[tid * rows_per + (prow - r_begin)] = dg[tid];
      }
      __syncthreads();
      if (cta == 0 && tid < w) S[tid * rows_per + jj] = s_prow[tid];
    }
    if (cta == 0 && tid == 0) {
      ipiv[c0 + jj] = c0 + prow;
      if (s_prow[jj] == 0.0) atomicCAS(info, 0, c0 + jj + 1);
    }
    __syncthreads();
    const double piv = s_prow[jj];
    if (piv != 0.0) {
      const double rinv = 1.0 / piv;
      const int lo2 = max(0, jj + 1 - r_begin);
      for (int r = lo2 + tid; r < r_cnt; r += CSTRIP_THREADS) {
        const double l = S[jj * rows_per + r] * rinv;
        S[jj * rows_per + r] = l;
        for (int c = jj + 1; c < w; ++c) S[c * rows_per + r] -= l * s_prow[c];
      }
    }
    __syncthreads();
  }
  for (int c = 0; c < w; ++c) {
    double* dst = A + (int64_t)(c0 + c) * ld + c0 + r_begin;
    for (int r = tid; r < r_cnt; r += CSTRIP_THREADS) dst[r] = S[c * rows_per + r];
  }
  cluster.sync();  // nobody may exit while its shared memory can still be read remotely
}

// ---------------------------------------------------------------------------------------
// Register-resident cluster strip kernel: every thread keeps RPT rows x W columns of the strip
// in registers, so the rank-1 updates are pure register DFMAs.  Each CTA PUSHES its pivot
// candidate (|value|, row, the row's W values) into every CTA's shared-memory inbox through
// DSMEM before the cluster barrier; after the barrier all pivot data is local.  The barrier is
// split (arrive.release ... wait.acquire).  One barrier per column, no global-memory traffic
// inside the column loop.
// ---------------------------------------------------------------------------------------
constexpr int RSTRIP_THREADS = 512;
constexpr int MAX_CLUSTER = 16;

__device__ __forceinline__ void cluster_arrive_release() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void cluster_wait_acquire() {
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

template <int RPT>
__global__ void __launch_bounds__(RSTRIP_THREADS, 1)
    lu_strip_reg_kernel(double* __restrict__ A, int64_t ld, int N, int c0, int w, int rows_per,
                        int* __restrict__ ipiv, int* __restrict__ info) {
  __shared__ StripRec inbox[2][MAX_CLUSTER];
  __shared__ double diagbox[2][W];
  __shared__ StripRec mine;
  __shared__ double mydiag[W];
  __shared__ double s_red_val[RSTRIP_THREADS / 32];
  __shared__ int s_red_row[RSTRIP_THREADS / 32];
  __shared__ double s_prow[W];
  __shared__ int s_piv_row;

  cg::cluster_group cluster = cg::this_cluster();
  const int G = (int)cluster.num_blocks();
  const int cta = (int)cluster.block_rank();
  const int tid = threadIdx.x;
  const int m = N - c0;
  const int r_begin = cta * rows_per;
  const int r_cnt = max(0, min(rows_per, m - r_begin));

  double a[RPT][W];
  int rowidx[RPT];  // panel-relative row, or -1
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    const int lr = tid + RSTRIP_THREADS * i;
    rowidx[i] = (lr < r_cnt) ? (r_begin + lr) : -1;
#pragma unroll
    for (int c = 0; c < W; ++c)
      a[i][c] = (rowidx[i] >= 0 && c < w) ? A[(int64_t)(c0 + c) * ld + c0 + rowidx[i]] : 0.0;
  }
  // all CTAs of the cluster must be running before anyone writes into a remote inbox
  cluster_arrive_release();
  cluster_wait_acquire();

  // Column loop, NOT unrolled (a fully unrolled body is ~180 KB of code that is executed once per
  // launch: the first version was instruction-cache-miss bound).  The register columns are
  // rotated by one after every step so the current column is always a[.][0]; positions
  // 1 .. w-1-jj hold the columns still to be updated, the rest already-final L columns.
#pragma unroll 1
  for (int jj = 0; jj < w; ++jj) {
    const int buf = jj & 1;
    // 1. local candidate
    double best = -1.0;
    int best_row = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const double v = fabs(a[i][0]);
      if (rowidx[i] >= jj && v > best) {
        best = v;
        best_row = rowidx[i];
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double ov = __shfl_down_sync(0xffffffffu, best, off);
      const int orow = __shfl_down_sync(0xffffffffu, best_row, off);
      if (ov > best || (ov == best && orow < best_row)) {
        best = ov;
        best_row = orow;
      }
    }
    if ((tid & 31) == 0) {
      s_red_val[tid >> 5] = best;
      s_red_row[tid >> 5] = best_row;
    }
    __syncthreads();
    if (tid < 32) {
      best = (tid < RSTRIP_THREADS / 32) ? s_red_val[tid] : -1.0;
      best_row = (tid < RSTRIP_THREADS / 32) ? s_red_row[tid] : 0x7fffffff;
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, best, off);
        const int orow = __shfl_down_sync(0xffffffffu, best_row, off);
        if (ov > best || (ov == best && orow < best_row)) {
          best = ov;
          best_row = orow;
        }
      }
      if (tid == 0) {
        mine.absval = best;
        mine.row = best_row;
      }
    }
    __syncthreads();
    // owner threads deposit the candidate row and (CTA 0) the current diagonal row
    {
      const int cand_row = mine.row;
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        if (rowidx[i] == cand_row && cand_row != 0x7fffffff) {
#pragma unroll
          for (int c = 0; c < W; ++c) mine.vals[c] = a[i][c];
        }
        if (rowidx[i] == jj) {
#pragma unroll
          for (int c = 0; c < W; ++c) mydiag[c] = a[i][c];
        }
      }
    }
    __syncthreads();
    // 2. push my record into every CTA's inbox (DSMEM stores), then the cluster barrier
    {
      constexpr int REC_WORDS = sizeof(StripRec) / 8;  // 18
      const int wd = tid & 31;
      for (int dst = tid / 32; dst < G; dst += RSTRIP_THREADS / 32) {  // one warp per destination CTA
        StripRec* remote = cluster.map_shared_rank(&inbox[buf][cta], dst);
        if (wd < REC_WORDS) reinterpret_cast<double*>(remote)[wd] = reinterpret_cast<const double*>(&mine)[wd];
        if (cta == 0 && wd < W) {
          double* rd = cluster.map_shared_rank(&diagbox[buf][0], dst);
          rd[wd] = mydiag[wd];
        }
      }
    }
    cluster_arrive_release();
    cluster_wait_acquire();
    // 3. winner (local data only)
    if (tid < 32) {
      double bv = -1.0;
      int br = 0x7fffffff, bc = 0;
      if (tid < G) {
        bv = inbox[buf][tid].absval;
        br = inbox[buf][tid].row;
        bc = tid;
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, bv, off);
        const int orow = __shfl_down_sync(0xffffffffu, br, off);
        const int oc = __shfl_down_sync(0xffffffffu, bc, off);
        if (ov > bv || (ov == bv && orow < br)) {
          bv = ov;
          br = orow;
          bc = oc;
        }
      }
      br = __shfl_sync(0xffffffffu, br, 0);
      bc = __shfl_sync(0xffffffffu, bc, 0);
      if (tid == 0) s_piv_row = br;
      if (tid < W) s_prow[tid] = inbox[buf][bc].vals[tid];
    }
    __syncthreads();
    const int prow = s_piv_row;
    // 4. swap rows jj <-> prow (inside the registers of their owner threads)
    if (prow != jj) {
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        if (rowidx[i] == prow) {
#pragma unroll
          for (int c = 0; c < W; ++c) a[i][c] = diagbox[buf][c];
        } else if (rowidx[i] == jj) {
#pragma unroll
          for (int c = 0; c < W; ++c) a[i][c] = s_prow[c];
        }
      }
    }
    if (cta == 0 && tid == 0) {
      ipiv[c0 + jj] = c0 + prow;
      if (s_prow[0] == 0.0) atomicCAS(info, 0, c0 + jj + 1);
    }
    // 5. scale + rank-1 update fused with the rotation: a[.][c-1] <- a[.][c] - l * prow[c]
    const double piv = s_prow[0];
    const double rinv = (piv != 0.0) ? 1.0 / piv : 0.0;
    const int nright = w - jj;  // positions 1 .. nright-1 are still to be updated
    double l[RPT], keep[RPT];
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const bool act = rowidx[i] > jj && piv != 0.0;
      l[i] = act ? a[i][0] * rinv : 0.0;
      keep[i] = act ? l[i] : a[i][0];
    }
#pragma unroll
    for (int c = 1; c < W; ++c) {
      const double pc = (c < nright) ? s_prow[c] : 0.0;
#pragma unroll
      for (int i = 0; i < RPT; ++i) a[i][c - 1] = a[i][c] - l[i] * pc;
    }
#pragma unroll
    for (int i = 0; i < RPT; ++i) a[i][W - 1] = keep[i];
    // s_prow / mine / mydiag are rewritten only after the next column's first __syncthreads
  }
  // undo the remaining rotation (w < W only) so a[.][c] is column c again
  for (int extra = w; extra < W; ++extra) {
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const double t = a[i][0];
#pragma unroll
      for (int c = 0; c < W - 1; ++c) a[i][c] = a[i][c + 1];
      a[i][W - 1] = t;
    }
  }
  double* Aw = A;
  asm volatile("" : "+l"(Aw));  // fresh address arithmetic: do not keep 16 x RPT pointers live
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
    if (rowidx[i] >= 0) {
      double* dst = Aw + (int64_t)c0 * ld + c0 + rowidx[i];
#pragma unroll
      for (int c = 0; c < W; ++c) {
        if (c < w) *dst = a[i][c];
        dst += ld;
      }
    }
  }
  // nobody may exit while a peer can still write into its inbox
  cluster_arrive_release();
  cluster_wait_acquire();
}

// ---------------------------------------------------------------------------------------
// After a strip [c0, c0+w) is factored: apply its row swaps to the other panel columns
// [pc0, c0) and [c0+w, pc1), and for the right part also solve with the strip's unit-lower
// W x W block (rows c0..c0+w of those columns become rows of U).  One thread per column.
// ---------------------------------------------------------------------------------------
__global__ void lu_panel_swap_trsm_kernel(double* __restrict__ A, int64_t ld, int c0, int w, int pc0,
                                          int pc1, const int* __restrict__ ipiv) {
  __shared__ double L[W][W + 1];
  __shared__ int piv[W];
  const int tid = threadIdx.x;
  if (tid < w) piv[tid] = ipiv[c0 + tid];
  for (int e = tid; e < w * w; e += blockDim.x) {
    const int r = e % w, c = e / w;
    L[r][c] = A[(int64_t)(c0 + c) * ld + c0 + r];
  }
  __syncthreads();
  const int ncol_left = c0 - pc0;
  const int ncol = ncol_left + (pc1 - (c0 + w));
  for (int t = blockIdx.x * blockDim.x + tid; t < ncol; t += gridDim.x * blockDim.x) {
    const bool right = t >= ncol_left;
    const int col = right ? (c0 + w + (t - ncol_left)) : (pc0 + t);
    double* a = A + (int64_t)col * ld;
    for (int i = 0; i < w; ++i) {
      const int p = piv[i];
      if (p != c0 + i) {
        const double tmp = a[c0 + i];
        a[c0 + i] = a[p];
        a[p] = tmp;
      }
    }
    if (right) {
      double x[W];
#pragma unroll
      for (int i = 0; i < W; ++i) x[i] = (i < w) ? a[c0 + i] : 0.0;
#pragma unroll
      for (int i = 0; i < W; ++i) {
#pragma unroll
        for (int q = 0; q < i; ++q) x[i] -= L[i][q] * x[q];
      }
#pragma unroll
      for (int i = 0; i < W; ++i)
        if (i < w) a[c0 + i] = x[i];
    }
  }
}

// ---------------------------------------------------------------------------------------
// TRSM: B (nb x ncols, column-major, ldb) <- L^-1 B with L = unit lower nb x nb block of the
// factored panel (column-major, ldl).  One CTA per TRSM_COLS columns of B, B block in shared
// memory, rows resolved in order with the multipliers read through L1/L2.
// ---------------------------------------------------------------------------------------
constexpr int TRSM_COLS = 32;

__global__ void __launch_bounds__(256)
    lu_trsm_kernel(const double* __restrict__ L, int64_t ldl, int nb, double* __restrict__ B,
                   int64_t ldb, int ncols) {
  extern __shared__ __align__(16) double sm[];
  double* Bs = sm;                       // [TRSM_COLS][nb+1]
  double* Ls = sm + TRSM_COLS * (nb + 1);  // [nb][nb+1] column-major with padding: Ls[c*(nb+1)+r]
  const int tid = threadIdx.x;
  const int cb = blockIdx.x * TRSM_COLS;
  const int nc = min(TRSM_COLS, ncols - cb);
  for (int e = tid; e < nc * nb; e += blockDim.x) {
    const int r = e % nb, c = e / nb;
    Bs[c * (nb + 1) + r] = B[(int64_t)(cb + c) * ldb + r];
  }
  for (int e = tid; e < nb * nb; e += blockDim.x) {
    const int r = e % nb, c = e / nb;
    Ls[c * (nb + 1) + r] = (r > c) ? L[(int64_t)c * ldl + r] : 0.0;
  }
  __syncthreads();
  // Blocked forward substitution, TB rows at a time: (i) one thread per column solves the TB x TB
  // unit-lower diagonal block from registers, (ii) all threads apply the rank-TB update to the rows
  // below.  2 barriers per TB rows instead of 1 per row.
  constexpr int TB = 16;
  const int c = tid % TRSM_COLS;
  const int rg = tid / TRSM_COLS;
  constexpr int RG = 256 / TRSM_COLS;
  for (int t0 = 0; t0 < nb; t0 += TB) {
    const int tb = min(TB, nb - t0);
    if (tid < nc) {
      double x[TB];
#pragma unroll
      for (int i = 0; i < TB; ++i) x[i] = (i < tb) ? Bs[tid * (nb + 1) + t0 + i] : 0.0;
#pragma unroll
      for (int i = 1; i < TB; ++i) {
#pragma unroll
        for (int q = 0; q < i; ++q)
          if (i < tb) x[i] -= Ls[(t0 + q) * (nb + 1) + t0 + i] * x[q];
      }
#pragma unroll
      for (int i = 0; i < TB; ++i)
        if (i < tb) Bs[tid * (nb + 1) + t0 + i] = x[i];
    }
    __syncthreads();
    if (c < nc && t0 + tb < nb) {
      double x[TB];
#pragma unroll
      for (int i = 0; i < TB; ++i) x[i] = (i < tb) ? Bs[c * (nb + 1) + t0 + i] : 0.0;
      for (int r = t0 + tb + rg; r < nb; r += RG) {
        double acc = 0.0;
#pragma unroll
        for (int q = 0; q < TB; ++q) acc += Ls[(t0 + q) * (nb + 1) + r] * x[q];
        Bs[c * (nb + 1) + r] -= acc;
      }
    }
    __syncthreads();
  }
  for (int e = tid; e < nc * nb; e += blockDim.x) {
    const int r = e % nb, cc = e / nb;
    B[(int64_t)(cb + cc) * ldb + r] = Bs[cc * (nb + 1) + r];
  }
}

// ---------------------------------------------------------------------------------------
// Triangular solves for Q x = b with Q^T = A = P L U:  U^T w = b,  L^T v = w,  x = P v.
// Each "row" of U^T / L^T is a contiguous column of A, so the off-diagonal part is a
// coalesced dot-product GEMV (one warp per column), the diagonal block a small serial solve.
// ---------------------------------------------------------------------------------------
constexpr int SB = 128;  // diagonal block of the triangular solves

// b[c] -= sum_{r in [r0,r1)} A[r + c*ld] * x[r]   for c in [c_begin, c_end); one warp per column
__global__ void __launch_bounds__(256)
    lu_gemv_t_kernel(const double* __restrict__ A, int64_t ld, int r0, int r1, int c_begin,
                     int c_end, const double* x, double* b) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int c = c_begin + warp;
  if (c >= c_end) return;
  const double* a = A + (int64_t)c * ld;
  double acc = 0.0;
  for (int r = r0 + lane; r < r1; r += 32) acc += a[r] * x[r];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
  if (lane == 0) b[c] -= acc;
}

// Solve the SB x SB diagonal block in place (one CTA; block staged in dynamic shared memory).
//   UPPER_T: w[c] = (b[c] - sum_{k0<=r<c} A[r,c] w[r]) / A[c,c]      ascending c
//   else   : v[c] =  b[c] - sum_{c<r<k1} A[r,c] v[r]                 descending c (unit diag)
// The needed triangle is fetched by 512 threads with 8 independent loads in flight each (the
// first version's 128 threads x 128 dependent iterations cost ~30 us of pure DRAM latency per
// block), the next block on the walk is prefetched into L2, and the substitution runs in 32 x 32
// sub-blocks: one warp solves a sub-block with register-resident columns and shuffles, then 128
// threads eliminate it from the rest of the block.
constexpr int DIAG_THREADS = 512;
template <bool UPPER_T>
__global__ void __launch_bounds__(DIAG_THREADS)
    lu_diag_solve_kernel(const double* __restrict__ A, int64_t ld, int k0, int k1, int N,
                         double* __restrict__ b) {
  extern __shared__ __align__(16) double T[];  // T[r*(SB+1) + c] = A[k0+r, k0+c]
  __shared__ double x[SB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nbk = k1 - k0;
#pragma unroll 8
  for (int i = 0; i < SB * SB / DIAG_THREADS; ++i) {
    const int e = tid + i * DIAG_THREADS;
    const int r = e % SB, c = e / SB;
    const bool need = UPPER_T ? (r <= c) : (r > c);
    if (need && r < nbk && c < nbk) T[r * (SB + 1) + c] = A[(int64_t)(k0 + c) * ld + k0 + r];
  }
  if (tid < nbk) x[tid] = b[k0 + tid];
  {  // L2 prefetch of the next diagonal block on the walk (8 lines of 128 B per column)
    const int nk0 = UPPER_T ? k0 + SB : k0 - SB;
    if (nk0 >= 0 && nk0 < N) {
      for (int line = tid; line < SB * 8; line += DIAG_THREADS) {
        const int c = nk0 + (line >> 3), r = nk0 + (line & 7) * 16;
        if (c < N && r < N) asm volatile("prefetch.global.L2 [%0];" ::"l"(A + (int64_t)c * ld + r));
      }
    }
  }
  __syncthreads();
  const int nsub = (nbk + 31) / 32;
  for (int q = 0; q < nsub; ++q) {
    const int sb = UPPER_T ? 32 * q : 32 * (nsub - 1 - q);
    const int len = min(32, nbk - sb);
    if (warp == 0) {
      const int c = sb + lane;
      double xv = (lane < len) ? x[c] : 0.0;
      double tc[32];  // column c of the sub-block: T[sb + r][c]
#pragma unroll
      for (int r = 0; r < 32; ++r) {
        const bool need = UPPER_T ? (r < lane) : (r > lane);
        tc[r] = (need && r < len && lane < len) ? T[(sb + r) * (SB + 1) + c] : 0.0;
      }
      if (UPPER_T) {
        const double rd = (lane < len) ? 1.0 / T[c * (SB + 1) + c] : 1.0;
#pragma unroll
        for (int r = 0; r < 32; ++r) {
          const double xr = __shfl_sync(0xffffffffu, xv, r) * __shfl_sync(0xffffffffu, rd, r);
          if (lane == r) xv = xr;
          xv -= tc[r] * xr;  // tc[r] == 0 for lanes <= r
        }
      } else {
#pragma unroll
        for (int r = 31; r >= 0; --r) {
          const double xr = __shfl_sync(0xffffffffu, xv, r);
          xv -= tc[r] * xr;  // tc[r] == 0 for lanes >= r
        }
      }
      if (lane < len) x[c] = xv;
    }
    __syncthreads();
    if (tid < nbk && (UPPER_T ? (tid >= sb + 32) : (tid < sb))) {
      double acc = x[tid];
      for (int r = sb; r < sb + len; ++r) acc -= T[r * (SB + 1) + tid] * x[r];
      x[tid] = acc;
    }
    __syncthreads();
  }
  if (tid < nbk) b[k0 + tid] = x[tid];
}

__global__ void lu_gather_kernel(const double* __restrict__ v, const int* __restrict__ perm, int N,
                                 double* __restrict__ x) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) x[i] = v[perm[i]];
}

}  // namespace

int mmr_dgemm_launch(mmr_ctx* ctx, cudaStream_t s, int64_t m, int64_t n, int64_t k, double alpha,
                     const double* A, int64_t lda, const double* B, int64_t ldb, double beta,
                     double* C, int64_t ldc, int* fail, int dep, int check_val, const mmr_gemm::GemmAlt* alt) {
  using namespace mmr_gemm;
  if (alt && (alt->m <= 0 || alt->n <= 0)) alt = nullptr;
  if (m <= 0 || n <= 0) {
    if (!alt) return 0;
    return mmr_dgemm_launch(ctx, s, alt->m, alt->n, alt->k, alpha, alt->A, alt->lda, alt->B, alt->ldb, beta, alt->C, alt->ldc,
                            fail, dep, 0, nullptr);
  }
  const bool aligned = (((uintptr_t)A | (uintptr_t)B) % 16 == 0) && (lda % 2 == 0) && (ldb % 2 == 0);
  const bool subtract = (alpha == -1.0 && beta == 1.0);
  typedef void (*kern_t)(int, int, int, double, const double*, int64_t, const double*, int64_t, double,
                         double*, int64_t, int*, int, int);
  static const kern_t kerns[4] = {dgemm_nn_kernel<false, false>, dgemm_nn_kernel<false, true>,
                                  dgemm_nn_kernel<true, false>, dgemm_nn_kernel<true, true>};
  const int which = (aligned ? 2 : 0) + (subtract ? 1 : 0);
  int arc = mmr_func_smem(ctx, (const void*)kerns[which], SMEM_BYTES);
  if (arc) return arc;
  dim3 grid((unsigned)ceil_div64(m, BM), (unsigned)ceil_div64(n, BN));
  const int64_t ntiles = (int64_t)grid.x * grid.y;
  static int fast = -1;
  if (fast < 0) {
    const char* e = getenv("MMR_DGEMM_FAST");
    fast = (e && e[0] == '0') ? 0 : 1;
  }
  if ((arc = mmr_func_smem(ctx, (const void*)dgemm_nn_fast_kernel<true>, SMEM_BYTES))) return arc;
  if ((arc = mmr_func_smem(ctx, (const void*)dgemm_nn_fast_kernel<false>, SMEM_BYTES))) return arc;
  static int prefetch_on = -1;
  if (prefetch_on < 0) {
    const char* e = getenv("MMR_DGEMM_PREFETCH");
    prefetch_on = (e && e[0] == '0') ? 0 : 1;
  }
  const bool fast_ok = fast && aligned && k > 0 && k % BK == 0;
  if (alt) {
    // paired launch: both products on the fast path, same alpha / beta; otherwise one after the other
    const bool alt_ok = (((uintptr_t)alt->A | (uintptr_t)alt->B) % 16 == 0) && (alt->lda % 2 == 0) && (alt->ldb % 2 == 0) &&
                        alt->k > 0 && alt->k % BK == 0;
    if (!(fast_ok && alt_ok && !subtract)) {
      int r = mmr_dgemm_launch(ctx, s, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, fail, dep, check_val, nullptr);
      if (r) return r;
      return mmr_dgemm_launch(ctx, s, alt->m, alt->n, alt->k, alpha, alt->A, alt->lda, alt->B, alt->ldb, beta, alt->C,
                              alt->ldc, fail, dep, 0, nullptr);
    }
    grid.x = std::max(grid.x, (unsigned)ceil_div64(alt->m, BM));
    grid.y = std::max(grid.y, (unsigned)ceil_div64(alt->n, BN));
    grid.z = 2;
  }
  const GemmAlt alt_v = alt ? *alt : GemmAlt();
  // C -= A B with all three operands 16-byte aligned, even leading dimensions and even m: the kernels of
  // dgemm_tma.cuh move C by bulk copies (TMA) through shared memory instead of strided LDG / STG.
  //   MMR_DGEMM_TMA     = minimum number of tiles for the per-tile TMA kernel (default: one full wave; 0 = never)
  //   MMR_DGEMM_PERSIST = minimum tiles per group for the persistent kernel (default 0 = never: a grid that holds
  //                       every SM for the whole update keeps the look-ahead panel chain of the LU waiting)
  static int tma_min_tiles = -1, persist_min_rounds = -1;
  if (tma_min_tiles < 0) {
    const char* e = getenv("MMR_DGEMM_TMA");
    tma_min_tiles = e ? atoi(e) : 2 * ctx->sm_count;
    e = getenv("MMR_DGEMM_PERSIST");
    persist_min_rounds = e ? atoi(e) : 0;
  }
  const bool tma_ok = fast_ok && subtract && !alt && check_val == 0 && ntiles < (1LL << 30) && (uintptr_t)C % 16 == 0 &&
                      ldc % 2 == 0 && m % 2 == 0;
  if (tma_ok && persist_min_rounds > 0 && ntiles >= (int64_t)persist_min_rounds * 2 * ctx->sm_count) {
    auto kern = dgemm_sub_persist_kernel<true, 0>;
    if ((arc = mmr_func_smem(ctx, (const void*)kern, P_SMEM_BYTES))) return arc;
    kern<<<ctx->sm_count, P_THREADS, P_SMEM_BYTES, s>>>((int)m, (int)n, (int)k, A, lda, B, ldb, C, ldc, (int)grid.x, (int)ntiles,
                                                        fail, dep);
    MMR_CHECK_LAUNCH(ctx);
    return 0;
  }
  // MMR_DGEMM_TMA2 (default 1): the tensor-map kernel with the barrier-free main loop (dgemm_sub_tma2_kernel, 34.8 TF
  // at k = 256); 0: A / B by cp.async (dgemm_sub_tma_kernel, 33.9 TF); 2 / 3 / 4: variants kept for A/B timing.
  static int tma2_on = -1;
  if (tma2_on < 0) {
    const char* e = getenv("MMR_DGEMM_TMA2");
    tma2_on = e ? atoi(e) : 1;
  }
  if (tma_ok && tma2_on && tma_min_tiles > 0 && ntiles >= tma_min_tiles) {
    alignas(64) CUtensorMap mapA, mapB;  // passed by value (__grid_constant__ parameters)
    if (mmr_tmap_2d(&mapA, A, (uint64_t)m, (uint64_t)k, lda, 16, 16) && mmr_tmap_2d(&mapB, B, (uint64_t)k, (uint64_t)n, ldb, 16, BN)) {
      auto kern = dgemm_sub_tma2_kernel<0>;
      if (tma2_on == 2) kern = dgemm_sub_tma2_kernel<16>;  // CTA barrier per k-tile
      if (tma2_on == 3) kern = dgemm_sub_tma2_kernel<32>;  // direct C - P epilogue
      if (tma2_on == 4) kern = dgemm_sub_tma2_kernel<64>;  // C tile by coalesced generic loads
      if ((arc = mmr_func_smem(ctx, (const void*)kern, T2_SMEM_BYTES))) return arc;
      kern<<<grid, THREADS, T2_SMEM_BYTES, s>>>(mapA, mapB, (int)m, (int)n, (int)k, C, ldc, fail, dep);
      MMR_CHECK_LAUNCH(ctx);
      return 0;
    }
  }
  if (tma_ok && tma_min_tiles > 0 && ntiles >= tma_min_tiles) {
    auto kern = dgemm_sub_tma_kernel<0>;
    if ((arc = mmr_func_smem(ctx, (const void*)kern, T_SMEM_BYTES))) return arc;
    kern<<<grid, THREADS, T_SMEM_BYTES, s>>>((int)m, (int)n, (int)k, A, lda, B, ldb, C, ldc, fail, dep);
    MMR_CHECK_LAUNCH(ctx);
    return 0;
  }
  if (fast_ok) {
    // L2 prefetch distance for the C tiles: the number of CTAs resident at once (2 per SM)
    const int pf = (prefetch_on && !alt && ntiles > 2 * (int64_t)ctx->sm_count) ? 2 * ctx->sm_count : 0;
    if (subtract)
      dgemm_nn_fast_kernel<true><<<grid, THREADS, SMEM_BYTES, s>>>((int)m, (int)n, (int)k, alpha, A, lda, B, ldb, beta, C, ldc, pf, fail, dep, check_val, alt_v);
    else
      dgemm_nn_fast_kernel<false><<<grid, THREADS, SMEM_BYTES, s>>>((int)m, (int)n, (int)k, alpha, A, lda, B, ldb, beta, C, ldc, pf, fail, dep, check_val, alt_v);
    MMR_CHECK_LAUNCH(ctx);
    return 0;
  }
  kerns[which]<<<grid, THREADS, SMEM_BYTES, s>>>((int)m, (int)n, (int)k, alpha, A, lda, B, ldb, beta, C, ldc, fail, dep, check_val);
  MMR_CHECK_LAUNCH(ctx);
  return 0;
}

// B (nb0 + nb1 rows, ncols columns, in place) <- L_blk^-1 B in one launch (block_trsm.cuh).  Returns -2 when the
// shapes / alignments are not the ones the kernel is written for (the caller then issues the three products).
int mmr_block_trsm_launch(mmr_ctx* ctx, cudaStream_t s, int nb0, int nb1, const double* Linv0, const double* Linv1,
                          const double* L10, int64_t ldl, double* B, int64_t ldb, int ncols, int* fail, int dep) {
  using namespace mmr_gemm;
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("MMR_LU_BLOCK_TRSM");  // 0: three launches (round-2 start)
    on = (e && e[0] == '0') ? 0 : 1;
  }
  if (!on || nb0 != BM || nb1 <= 0 || nb1 > BM || nb1 % BK != 0 || ncols <= 0 || ldl % 2 != 0 || ldb % 2 != 0 ||
      (((uintptr_t)Linv0 | (uintptr_t)Linv1 | (uintptr_t)L10 | (uintptr_t)B) % 16) != 0)
    return -2;
  int arc = mmr_func_smem(ctx, (const void*)lu_block_trsm_kernel, BT_SMEM_BYTES);
  if (arc) return arc;
  const int ps = mmr_prof_begin(ctx, MMR_PROF_TRSM, s);
  lu_block_trsm_kernel<<<(unsigned)ceil_div64(ncols, BT_COLS), THREADS, BT_SMEM_BYTES, s>>>(nb0, nb1, Linv0, Linv1, L10, ldl, B, ldb,
                                                                                          ncols, fail, dep);
  MMR_CHECK_LAUNCH(ctx);
  mmr_prof_end(ctx, ps, (2.0 * nb0 * nb0 + 2.0 * nb1 * nb0 + 2.0 * nb1 * nb1) * (double)ncols, s);
  return 0;
}

extern "C" int mmr_block_trsm(mmr_ctx* ctx, int nb0, int nb1, const double* d_Linv0, const double* d_Linv1, const double* d_L10,
                              int64_t ldl, double* d_B, int64_t ldb, int ncols) {
  MMR_ARG(ctx != nullptr, "ctx is NULL");
  DeviceGuard g(ctx->device);
  const int r = mmr_block_trsm_launch(ctx, ctx->stream, nb0, nb1, d_Linv0, d_Linv1, d_L10, ldl, d_B, ldb, ncols, nullptr, 0);
  if (r == -2) {
    mmr_set_error("mmr_block_trsm: unsupported shape or alignment");
    return -1;
  }
  return r;
}

extern "C" int mmr_dgemm_nn(mmr_ctx* ctx, int64_t m, int64_t n, int64_t k, double alpha,
                            const double* d_A, int64_t lda, const double* d_B, int64_t ldb,
                            double beta, double* d_C, int64_t ldc) {
  MMR_ARG(ctx != nullptr, "ctx is NULL");
  MMR_ARG(m >= 0 && n >= 0 && k >= 0, "negative size");
  MMR_ARG(m < (1LL << 31) && n < (1LL << 31) && k < (1LL << 31), "size exceeds int32");
  MMR_ARG(lda >= m && ldb >= k && ldc >= m, "leading dimension too small");
  DeviceGuard g(ctx->device);
  return mmr_dgemm_launch(ctx, ctx->stream, m, n, k, alpha, d_A, lda, d_B, ldb, beta, d_C, ldc);
}

// Factor the panel columns [kb, kb+nb) rows [kb, N) of column-major A (local pointer/ld), pivots
// into d_ipiv[kb..kb+nb) (global row indices).  Used by the single-GPU and distributed drivers.
int mmr_lu_panel(mmr_ctx* ctx, cudaStream_t strm, double* A, int64_t ld, int N, int kb, int nb, int* d_ipiv,
                 int* d_info, void* d_scratch) {
  // capabilities of this context's device, probed once per context
  const size_t strip_smem_cap = 200 * 1024;
  const int max_coop_ctas = ctx->sm_count;  // 1 CTA per SM at up to 200 KB of shared memory
  if (ctx->panel_cluster_max < 0) {
    int rc0 = mmr_func_smem(ctx, (const void*)lu_strip_kernel, strip_smem_cap);
    if (rc0) return rc0;
    int found = 0;  // largest launchable cluster size for the cluster strip kernel
    if (mmr_func_smem(ctx, (const void*)lu_strip_cluster_kernel, strip_smem_cap) == 0 &&
        cudaFuncSetAttribute(lu_strip_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) ==
            cudaSuccess) {
      for (int cs = 16; cs >= 1; cs >>= 1) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cs);
        cfg.blockDim = dim3(CSTRIP_THREADS);
        cfg.dynamicSmemBytes = strip_smem_cap;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cs;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        int nclusters = 0;
        if (cudaOccupancyMaxActiveClusters(&nclusters, lu_strip_cluster_kernel, &cfg) == cudaSuccess &&
            nclusters >= 1) {
          found = cs;
          break;
        }
      }
    }
    cudaGetLastError();  // clear any probe error
    ctx->panel_reg_ok =
        cudaFuncSetAttribute(lu_strip_reg_kernel<1>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
        cudaFuncSetAttribute(lu_strip_reg_kernel<2>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
        cudaFuncSetAttribute(lu_strip_reg_kernel<3>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
    cudaGetLastError();
    ctx->panel_cluster_max = found;
  }
  const int cluster_max = ctx->panel_cluster_max;
  const bool reg_ok = ctx->panel_reg_ok;
  Cand* cand = (Cand*)d_scratch;
  double* candrow = (double*)((char*)d_scratch + 2 * 256 * sizeof(Cand));
  double* diagrow = candrow + 2 * 256 * W;
  for (int c0 = kb; c0 < kb + nb; c0 += W) {
    const int w = min(W, kb + nb - c0);
    const int m = N - c0;
    const int pidx = mmr_prof_begin(ctx, MMR_PROF_PANEL, strm);
    // cluster path: smallest power-of-two cluster that keeps >= 128 rows per CTA and fits smem
    int cs = 0;
    if (cluster_max >= 1) {
      cs = 1;
      while (cs < cluster_max && (m / cs > 256 || (size_t)W * ((m + cs - 1) / cs + 2) * sizeof(double) > strip_smem_cap))
        cs <<= 1;
      int rp = max((int)ceil_div64(m, cs), W);
      rp = (rp + 1) & ~1;
      if ((size_t)W * rp * sizeof(double) > strip_smem_cap) cs = 0;
    }
    // register-resident variant: rows per CTA <= RSTRIP_THREADS * 3
    int csr = 0;
    if (cluster_max >= 1 && reg_ok) {
      csr = 1;
      while (csr < cluster_max && m / csr > 512) csr <<= 1;
      if ((int)ceil_div64(m, csr) > RSTRIP_THREADS * 3) csr = 0;
    }
    if (csr >= 1) {
      const int rows_per = max((int)ceil_div64(m, csr), W);
      const int rpt = (int)ceil_div64(rows_per, RSTRIP_THREADS);
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(csr);
      cfg.blockDim = dim3(RSTRIP_THREADS);
      cfg.dynamicSmemBytes = 0;
      cfg.stream = strm;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = csr;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      if (rpt <= 1) MMR_CUDA(cudaLaunchKernelEx(&cfg, lu_strip_reg_kernel<1>, A, ld, N, c0, w, rows_per, d_ipiv, d_info));
      else if (rpt == 2) MMR_CUDA(cudaLaunchKernelEx(&cfg, lu_strip_reg_kernel<2>, A, ld, N, c0, w, rows_per, d_ipiv, d_info));
      else MMR_CUDA(cudaLaunchKernelEx(&cfg, lu_strip_reg_kernel<3>, A, ld, N, c0, w, rows_per, d_ipiv, d_info));
    } else if (cs >= 1) {
      int rows_per = max((int)ceil_div64(m, cs), W);
      rows_per = (rows_per + 1) & ~1;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(cs);
      cfg.blockDim = dim3(CSTRIP_THREADS);
      cfg.dynamicSmemBytes = (size_t)W * rows_per * sizeof(double);
      cfg.stream = strm;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = cs;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      MMR_CUDA(cudaLaunchKernelEx(&cfg, lu_strip_cluster_kernel, A, ld, N, c0, w, rows_per, d_ipiv, d_info));
    } else {
      int G = min(max_coop_ctas, max(1, m / MIN_ROWS_PER_CTA));
      int rows_per = (int)ceil_div64(m, G);
      rows_per = max(rows_per, W);
      rows_per = (rows_per + 1) & ~1;
      G = (int)ceil_div64(m, rows_per);
      const size_t smem = (size_t)W * rows_per * sizeof(double);
      if (smem > strip_smem_cap) {
        mmr_set_error("LU strip of %d rows does not fit the distributed shared memory (%zu B/CTA)", m, smem);
        return -5;
      }
      int w_arg = w, N_arg = N, c0_arg = c0;
      void* args[] = {&A, &ld, &N_arg, &c0_arg, &w_arg, &rows_per, &d_ipiv, &cand, &candrow, &diagrow, &d_info};
      MMR_CUDA(cudaLaunchCooperativeKernel((void*)lu_strip_kernel, dim3(G), dim3(STRIP_THREADS), args, smem,
                                           strm));
    }
    mmr_prof_end(ctx, pidx, (double)w, strm);
    ctx->launches++;
    const int ncol_other = nb - w;
    if (ncol_other > 0) {
      lu_panel_swap_trsm_kernel<<<1, 128, 0, strm>>>(A, ld, c0, w, kb, kb + nb, d_ipiv);
      MMR_CHECK_LAUNCH(ctx);
    }
    const int nright = kb + nb - (c0 + w);
    const int mbelow = N - (c0 + w);
    if (nright > 0 && mbelow > 0) {
      const int pg = mmr_prof_begin(ctx, MMR_PROF_PANEL_GEMM, strm);
      int rc = mmr_dgemm_launch(ctx, strm, mbelow, nright, w, -1.0,
                                A + (int64_t)c0 * ld + (c0 + w), ld,
                                A + (int64_t)(c0 + w) * ld + c0, ld, 1.0,
                                A + (int64_t)(c0 + w) * ld + (c0 + w), ld);
      mmr_prof_end(ctx, pg, 2.0 * mbelow * nright * w, strm);
      if (rc) return rc;
    }
  }
  return 0;
}

size_t mmr_lu_panel_scratch_bytes() { return 2 * 256 * sizeof(Cand) + (2 * 256 * W + 2 * W) * sizeof(double); }

// =======================================================================================
// Speculative panel: "the pivot is on the diagonal" (true for the demag matrices Q = I - chi*T
// of every BASELINE config), verified a posteriori.
//
//   1. lu_diag_nopiv_tile_kernel  LU without interchanges of the nb x nb diagonal block (one CTA)
//   2. lu_tri_inv_kernel      L11^-1 and U11^-1                                   (8 CTAs)
//   3. DMMA GEMM              L21 = A21 * U11^-1                                   (all SMs)
//   4. lu_spec_check / commit max |l_ij| <= 1 ?  then copy the packed panel into A, ipiv = identity
//
// Partial pivoting picks row j at step j exactly when every multiplier of that step has
// magnitude <= 1 (ties go to the lowest index = the diagonal, as in LAPACK's idamax), and as long
// as that holds the partially eliminated matrix is the one of the unpivoted factorization.  So
// "all |l_ij| <= 1" <=> the pivoted factorization performs no interchange in this panel, and then
// both produce the same L and U (up to the rounding of a different operation order).  Nothing is
// written to A before the check, so a failed check simply falls back to the pivoting strip
// kernels above.  This turns the latency-bound chain of 128 cluster-synchronised column steps
// (~0.5 ms per panel) into one small CTA kernel plus tensor-core work (~0.1 ms).
// =======================================================================================
namespace {

// 1/a without the division slow path: MUFU.RCP64H seed (>= 20 bits) + two Newton steps (-> 40 -> 80
// bits), straight-line so that it overlaps the independent DFMAs around it; it sits on the critical
// path of every elimination step.  (LAPACK's dgetf2 also scales by a reciprocal.)
__device__ __forceinline__ double fast_rcp(double a) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
  x = fma(x, fma(-a, x, 1.0), x);
  x = fma(x, fma(-a, x, 1.0), x);
  return x;
}

// LU without interchanges of the nb x nb (<= 128) diagonal block in ONE CTA.  Warp w owns the columns
// {w + 16 b}, lane t the rows {t + 32 a} -- 4 x 8 values per thread, the whole block lives in registers.
// Step j needs no shared-memory traffic for the pivot
// ROW at all (each warp broadcasts its own eight entries of row j from lane j % 32 with shuffles); only
// the multiplier column, computed by the one warp that holds column j, goes through shared memory
// (4 conflict-free stores, double buffered -> one CTA barrier per step).  All register indices are
// static: the step loop is split into 8 blocks of 16 steps in which the register row (a = j / 32) and the
// register column (b = j / 16) of the pivot are compile-time constants, and finished register rows /
// columns drop out of the unrolled update.  The reciprocal of the next pivot is computed by the lane that
// owns it right after its update.  Per step the critical path is the recurrence pivot -> reciprocal
// (MUFU.RCP64H + 2 Newton steps) -> multipliers -> update of the next pivot: ~0.4 us, 52 us per block.
// (A first layout -- one row segment per thread, the pivot row published through shared memory by its
// single owner thread -- measured the same 0.4 us per step with 30x more shared-memory stores.)
__global__ void __launch_bounds__(512, 1)
    lu_diag_nopiv_tile_kernel(const double* __restrict__ A, int64_t ld, int kb, int nb, double* __restrict__ S,
                              int64_t lds, int* __restrict__ fail) {
  __shared__ double lcol[2][128];
  __shared__ int s_bad;
  if (mmr_gemm::lu_poisoned(mmr_gemm::lu_failword(fail), kb - 1)) return;  // an earlier panel was rejected
  const int tid = threadIdx.x, tr = tid & 31, tc = tid >> 5;
  double v[4][8];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const int r = tr + 32 * a, c = tc + 16 * b;
      v[a][b] = (r < nb && c < nb) ? A[(int64_t)(kb + c) * ld + kb + r] : ((r == c) ? 1.0 : 0.0);
    }
  if (tid == 0) s_bad = 0;
  double rinv_next = (tid == 0) ? fast_rcp(v[0][0]) : 0.0;  // lives in the lane that holds the next pivot
  __syncthreads();
  bool bad = false;
#pragma unroll
  for (int aj = 0; aj < 4; ++aj) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int bj = 2 * aj + h;  // compile-time after unrolling
#pragma unroll 1
      for (int q = 0; q < 16; ++q) {
        const int j = 32 * aj + 16 * h + q;  // pivot row: lane jl, register row aj; pivot column: warp q, register column bj
        const int jl = 16 * h + q;
        if (tc == q) {  // the warp that holds column j: multipliers
          const double rinv = __shfl_sync(0xffffffffu, rinv_next, jl);
          const double piv = __shfl_sync(0xffffffffu, v[aj][bj], jl);
          if (piv == 0.0) bad = true;
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            const int r = tr + 32 * a;
            double l = 0.0;
            if (a >= aj && r > j) {
              l = v[a][bj] * rinv;
              if (!(fabs(l) <= 1.0)) bad = true;
              v[a][bj] = l;
            }
            lcol[j & 1][r] = l;
          }
        }
        __syncthreads();
        double l[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) l[a] = (a >= aj) ? lcol[j & 1][tr + 32 * a] : 0.0;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
          if (b < bj) continue;             // finished register columns (static)
          if (b == bj && tc <= q) continue;  // column tc + 16 bj <= j (warp-uniform)
          const double u = __shfl_sync(0xffffffffu, v[aj][b], jl);
#pragma unroll
          for (int a = 0; a < 4; ++a)
            if (a >= aj) v[a][b] -= l[a] * u;
        }
        // the reciprocal of the next pivot, in the lane that owns it (hidden behind the barrier wait)
        const int jn = j + 1;
        if (tc == (jn & 15) && tr == (jn & 31)) {
          const int an = min(jn >> 5, 3), bn = min(jn >> 4, 7);  // (aj, bj) or, at q == 15, the next block's
          double d = v[aj][bj];
          if (q == 15) {
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
              for (int b = 0; b < 8; ++b)
                if (a == an && b == bn) d = v[a][b];
          }
          rinv_next = fast_rcp(d);
        }
      }
    }
  }
  if (bad) s_bad = 1;
  __syncthreads();
  if (tid == 0 && s_bad) atomicCAS(fail, 0, kb + 1);  // sticky: the first rejected panel
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const int r = tr + 32 * a, c = tc + 16 * b;
      if (r < nb && c < nb) S[(int64_t)c * lds + r] = v[a][b];
    }
}

// Inverses of the two triangular factors packed in S (nb x nb, ld lds): blockIdx.y = 0: X = L^-1
// (unit lower), 1: X = U^-1 through Y = (U^T)^-1 (lower, explicit diagonal), X = Y^T.  One CTA per
// 32 columns of the lower-triangular inverse; blocked forward substitution on the identity, rows
// above the first column of the block are zero and skipped.  Output: dense nb x nb, ld = nb.
constexpr int TINV_COLS = 32;
__global__ void __launch_bounds__(256)
    lu_tri_inv_kernel(const double* __restrict__ S, int64_t lds, int nb, double* __restrict__ Linv,
                      double* __restrict__ Uinv, const int* __restrict__ fail, int dep) {
  extern __shared__ __align__(16) double sm[];
  if (mmr_gemm::lu_poisoned(mmr_gemm::lu_failword(fail), dep)) return;
  double* Bs = sm;                          // [TINV_COLS][nb+1]
  double* Ls = sm + TINV_COLS * (nb + 1);   // Ls[c*(nb+1) + r], r >= c
  const bool upper = blockIdx.y == 1;
  const int tid = threadIdx.x;
  const int cb = blockIdx.x * TINV_COLS;
  const int nc = min(TINV_COLS, nb - cb);
  // lower-triangular T in Ls[c*(nb+1) + r] (r >= c).  L: T(r, c) = S[c*lds + r] (unit diagonal);
  // U: T = U^T, T(r, c) = U(c, r) = S[r*lds + c].  `f` is the fast (contiguous) index of S.
  constexpr int INFLIGHT = 16;  // independent loads per thread and round trip
  for (int base = 0; base < nb * nb; base += INFLIGHT * 256) {
    double tmp[INFLIGHT];
#pragma unroll
    for (int k = 0; k < INFLIGHT; ++k) {
      const int e = base + tid + k * 256;
      const int f = e % nb, g = e / nb;  // S[g*lds + f]
      const bool need = (e < nb * nb) && (upper ? (g >= f) : (f > g));
      tmp[k] = need ? __ldg(S + (int64_t)g * lds + f) : 1.0;  // L: the unit diagonal
    }
#pragma unroll
    for (int k = 0; k < INFLIGHT; ++k) {
      const int e = base + tid + k * 256;
      const int f = e % nb, g = e / nb;
      if (e < nb * nb) {
        if (upper) {
          if (g >= f) Ls[f * (nb + 1) + g] = tmp[k];
        } else if (f >= g) {
          Ls[g * (nb + 1) + f] = tmp[k];
        }
      }
    }
  }
  for (int e = tid; e < nc * nb; e += blockDim.x) {
    const int r = e % nb, c = e / nb;
    Bs[c * (nb + 1) + r] = (r == cb + c) ? 1.0 : 0.0;
  }
  __syncthreads();
  __shared__ double s_rd[NB];  // reciprocals of the diagonal: 16 dependent divisions per block step otherwise
  if (tid < nb) s_rd[tid] = upper ? 1.0 / Ls[tid * (nb + 1) + tid] : 1.0;
  __syncthreads();
  constexpr int TB = 16;
  const int c = tid % TINV_COLS;
  const int rg = tid / TINV_COLS;
  constexpr int RG = 256 / TINV_COLS;
  for (int t0 = (cb / TB) * TB; t0 < nb; t0 += TB) {
    const int tb = min(TB, nb - t0);
    if (tid < nc) {
      double x[TB];
#pragma unroll
      for (int i = 0; i < TB; ++i) x[i] = (i < tb) ? Bs[tid * (nb + 1) + t0 + i] : 0.0;
#pragma unroll
      for (int i = 0; i < TB; ++i) {
        if (i < tb) {
#pragma unroll
          for (int qq = 0; qq < i; ++qq) x[i] -= Ls[(t0 + qq) * (nb + 1) + t0 + i] * x[qq];
          x[i] *= s_rd[t0 + i];
        }
      }
#pragma unroll
      for (int i = 0; i < TB; ++i)
        if (i < tb) Bs[tid * (nb + 1) + t0 + i] = x[i];
    }
    __syncthreads();
    if (c < nc && t0 + tb < nb) {
      double x[TB];
#pragma unroll
      for (int i = 0; i < TB; ++i) x[i] = (i < tb) ? Bs[c * (nb + 1) + t0 + i] : 0.0;
      // four rows per pass: four independent accumulation chains instead of one
      for (int r0 = t0 + tb + rg; r0 < nb; r0 += 4 * RG) {
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int qq = 0; qq < TB; ++qq) {
          if (qq < tb) {
            const double* lrow = Ls + (t0 + qq) * (nb + 1);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int r = r0 + k * RG;
              if (r < nb) acc[k] += lrow[r] * x[qq];
            }
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int r = r0 + k * RG;
          if (r < nb) Bs[c * (nb + 1) + r] -= acc[k];
        }
      }
    }
    __syncthreads();
  }
  // Bs[c][r] = Y(r, cb + c).  L: X(r, cb+c) = Y(r, cb+c);  U: X(cb+c, r) = Y(r, cb+c)
  for (int e = tid; e < nc * nb; e += blockDim.x) {
    const int r = e % nb, cc = e / nb;
    const double val = Bs[cc * (nb + 1) + r];
    if (upper) Uinv[(int64_t)r * nb + cb + cc] = val;
    else Linv[(int64_t)(cb + cc) * nb + r] = val;
  }
}

// unless this or an earlier panel was rejected: A[kb + r, kb + c] = S[r, c] for r < m, c < nb; ipiv[kb + c] = kb + c
__global__ void __launch_bounds__(256)
    lu_spec_commit_kernel(const double* __restrict__ S, int64_t lds, int m, int nb, double* __restrict__ A,
                          int64_t ld, int kb, int* __restrict__ ipiv, const int* __restrict__ fail) {
  if (mmr_gemm::lu_poisoned(mmr_gemm::lu_failword(fail), kb)) return;
  const int64_t total = (int64_t)m * nb;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e % m, c = e / m;
    A[(kb + c) * ld + kb + r] = S[c * lds + r];
  }
  if (ipiv && blockIdx.x == 0 && threadIdx.x < nb) ipiv[kb + threadIdx.x] = kb + threadIdx.x;
}

// Dst[r + c*ldd] = Src[r + c*lds] for r < rows, c < cols, unless a panel starting at or before `dep` was rejected
__global__ void __launch_bounds__(256)
    lu_copy_guarded_kernel(const double* __restrict__ Src, int64_t lds, int rows, int cols, double* __restrict__ Dst,
                           int64_t ldd, const int* __restrict__ fail, int dep) {
  if (mmr_gemm::lu_poisoned(mmr_gemm::lu_failword(fail), dep)) return;
  const int64_t total = (int64_t)rows * cols;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e % rows, c = e / rows;
    Dst[c * ldd + r] = Src[c * lds + r];
  }
}

// identity pivots for the columns [kb, kb + n) of a speculative block, into up to two arrays
__global__ void lu_fill_ipiv_kernel(int* __restrict__ p0, int* __restrict__ p1, int kb, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    if (p0) p0[i] = kb + i;
    if (p1) p1[i] = kb + i;
  }
}

}  // namespace

constexpr int BLOCK_TRSM_MAX_COLS = 512;  // wider ranges keep the three full-grid products

size_t mmr_lu_spec_scratch_bytes() { return sizeof(double) * NB * NB + 256; }  // U11^-1

// Speculative factorization of the panel columns [kb, kb+nb), rows [kb, N) of A (see above), entirely
// stream-ordered: no host synchronization.  S: m x nb packed panel buffer (ld = m = N - kb).  d_fail: the
// sticky device word of this factorization.  If the panel passes, A, ipiv, Linv and S hold its result; if it
// (or an earlier one) is rejected, *d_fail = (start of the first rejected panel) + 1, nothing of this panel is
// written, and every later operation that depends on it skips its stores (lu_poisoned) -- the host finds out
// when it reads *d_fail at the end and resumes from that panel with the pivoting kernels.
int mmr_lu_panel_spec(mmr_ctx* ctx, cudaStream_t strm, double* A, int64_t ld, int N, int kb, int nb, int* d_ipiv,
                      double* S, int64_t lds, double* Linv, void* d_spec, int* d_fail, bool commit,
                      const mmr_gemm::GemmAlt* with_l21 = nullptr) {
  const size_t inv_smem = ((size_t)TINV_COLS * (NB + 1) + (size_t)NB * (NB + 1)) * sizeof(double);
  {
    const int rc0 = mmr_func_smem(ctx, (const void*)lu_tri_inv_kernel, inv_smem);
    if (rc0) return rc0;
  }
  const int m = N - kb;
  double* Uinv = (double*)d_spec;
  static int panel_v1 = -1;
  if (panel_v1 < 0) {
    const char* e = getenv("MMR_LU_PANEL_V1");  // 1: the round-1 kernel pair (one-CTA LU + 8-CTA inverses)
    panel_v1 = (e && e[0] == '1') ? 1 : 0;
  }
  const int pidx = mmr_prof_begin(ctx, MMR_PROF_PANEL, strm);
  if (panel_v1) {
    lu_diag_nopiv_tile_kernel<<<1, 512, 0, strm>>>(A, ld, kb, nb, S, lds, d_fail);
    MMR_CHECK_LAUNCH(ctx);
    const size_t smem = ((size_t)TINV_COLS * (nb + 1) + (size_t)nb * (nb + 1)) * sizeof(double);
    lu_tri_inv_kernel<<<dim3((unsigned)ceil_div64(nb, TINV_COLS), 2), 256, smem, strm>>>(S, lds, nb, Linv, Uinv, d_fail, kb);
    MMR_CHECK_LAUNCH(ctx);
  } else {
    // one cluster launch: blocked no-pivot LU of the diagonal block + both triangular inverses (panel128.cuh)
    const int rc1 = mmr_func_smem(ctx, (const void*)mmr_panel::lu_diag_inv_cluster_kernel, mmr_panel::PANEL_SMEM);
    if (rc1) return rc1;
    mmr_panel::lu_diag_inv_cluster_kernel<<<mmr_panel::PCLUSTER, mmr_panel::PTHREADS, mmr_panel::PANEL_SMEM, strm>>>(
        A, ld, kb, nb, S, lds, Linv, Uinv, d_fail, nullptr);
    MMR_CHECK_LAUNCH(ctx);
  }
  mmr_prof_end(ctx, pidx, (double)nb, strm);
  if (m > nb) {
    const int pg = mmr_prof_begin(ctx, MMR_PROF_PANEL_GEMM, strm);
    // L21 = A21 U11^-1 with the speculation check (every multiplier |l| <= 1) fused into the epilogue
    // (with_l21: a second product that depends on this panel's diagonal block only rides in the same launch)
    int rc = mmr_dgemm_launch(ctx, strm, m - nb, nb, nb, 1.0, A + (int64_t)kb * ld + kb + nb, ld, Uinv, nb, 0.0,
                              S + nb, lds, d_fail, kb, kb + 1, with_l21);
    if (rc) return rc;
    mmr_prof_end(ctx, pg, 2.0 * (m - nb) * (double)nb * nb, strm);
  } else if (with_l21) {
    int rc = mmr_dgemm_launch(ctx, strm, 0, 0, 0, 1.0, nullptr, 0, nullptr, 0, 0.0, nullptr, 0, d_fail, kb, 0, with_l21);
    if (rc) return rc;
  }
  if (commit) {
    const unsigned g = (unsigned)std::min<int64_t>((int64_t)4 * ctx->sm_count, ceil_div64((int64_t)m * nb, 256 * 8));
    lu_spec_commit_kernel<<<g, 256, 0, strm>>>(S, lds, m, nb, A, ld, kb, d_ipiv, d_fail);
    MMR_CHECK_LAUNCH(ctx);
  }
  return 0;
}

// The diagonal-block step of the speculative panel on its own (tests / micro-benchmarks): no-pivot LU of the
// nb x nb (<= 128) leading block of column-major d_A into d_S (ld lds, packed L and U), L^-1 and U^-1 (dense
// nb x nb, ld nb).  *rejected_out = 1 if the speculation check fails (zero pivot or a multiplier above 1).
extern "C" int mmr_lu_diag_block(mmr_ctx* ctx, const double* d_A, int64_t ld, int nb, double* d_S, int64_t lds,
                                 double* d_Linv, double* d_Uinv, int* rejected_out) {
  MMR_ARG(ctx != nullptr, "ctx is NULL");
  MMR_ARG(nb >= 1 && nb <= NB && ld >= nb && lds >= nb, "bad block size / leading dimension");
  MMR_ARG(d_A && d_S && d_Linv && d_Uinv, "NULL pointer");
  DeviceGuard g(ctx->device);
  int rc = mmr_ws_reserve(ctx, 512);
  if (rc) return rc;
  rc = mmr_hpin_reserve(ctx, 512);
  if (rc) return rc;
  int* d_fail = (int*)ctx->ws;
  MMR_CUDA(cudaMemsetAsync(d_fail, 0, 512, ctx->stream));
  const char* v1 = getenv("MMR_LU_PANEL_V1");
  if (v1 && v1[0] == '1') {  // the round-1 kernel pair, for A/B timing
    const size_t smem = ((size_t)TINV_COLS * (nb + 1) + (size_t)nb * (nb + 1)) * sizeof(double);
    rc = mmr_func_smem(ctx, (const void*)lu_tri_inv_kernel, ((size_t)TINV_COLS * (NB + 1) + (size_t)NB * (NB + 1)) * sizeof(double));
    if (rc) return rc;
    lu_diag_nopiv_tile_kernel<<<1, 512, 0, ctx->stream>>>(d_A, ld, 0, nb, d_S, lds, d_fail);
    MMR_CHECK_LAUNCH(ctx);
    lu_tri_inv_kernel<<<dim3((unsigned)ceil_div64(nb, TINV_COLS), 2), 256, smem, ctx->stream>>>(d_S, lds, nb, d_Linv, d_Uinv, d_fail, 0);
    MMR_CHECK_LAUNCH(ctx);
  } else {
    rc = mmr_func_smem(ctx, (const void*)mmr_panel::lu_diag_inv_cluster_kernel, mmr_panel::PANEL_SMEM);
    if (rc) return rc;
    mmr_panel::lu_diag_inv_cluster_kernel<<<mmr_panel::PCLUSTER, mmr_panel::PTHREADS, mmr_panel::PANEL_SMEM, ctx->stream>>>(
        d_A, ld, 0, nb, d_S, lds, d_Linv, d_Uinv, d_fail, (long long*)((char*)ctx->ws + 64));
    MMR_CHECK_LAUNCH(ctx);
    MMR_CUDA(cudaMemcpyAsync((char*)ctx->hpin + 8, (char*)ctx->ws + 64, 30 * sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
  }
  MMR_CUDA(cudaMemcpyAsync(ctx->hpin, d_fail, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  MMR_CUDA(cudaStreamSynchronize(ctx->stream));
  if (rejected_out) *rejected_out = (*(int*)ctx->hpin != 0) ? 1 : 0;
  return 0;
}

// SM clock stamps of the last mmr_lu_diag_block call (phase boundaries of the cluster kernel; micro-benchmarks)
extern "C" int mmr_lu_diag_block_clocks(mmr_ctx* ctx, long long* out30) {
  MMR_ARG(ctx != nullptr && out30 != nullptr && ctx->hpin != nullptr, "no diagonal-block call recorded");
  memcpy(out30, (char*)ctx->hpin + 8, 30 * sizeof(long long));
  return 0;
}

// ---------------------------------------------------------------------------------------
// Row interchanges (LAPACK laswp) for general matrices.  The nb sequential swaps of a panel touch at
// most 2*nb rows; one warp turns the pivot sequence into a permutation of those rows
// ("the data that ends at row rows[t] comes from row src[t]"), then one warp per column gathers
// and scatters.  (The first version walked the nb dependent swaps serially per column: 51 ms per
// factorization of a random 24000^2 matrix.)  Panels without any interchange -- the usual case for
// the demag matrix Q = I + chi*N -- cost one early-exit launch.
// ---------------------------------------------------------------------------------------
namespace {
constexpr int LASWP_MAX = 2 * NB;
struct LaswpPerm {
  int count;
  int moved;
  int rows[LASWP_MAX];
  int src[LASWP_MAX];
};

__global__ void __launch_bounds__(32)
    lu_laswp_prepare_kernel(const int* __restrict__ ipiv, int k0, int nb, LaswpPerm* __restrict__ out,
                            const int* __restrict__ fail) {
  __shared__ int piv[NB], rows[LASWP_MAX], cur[LASWP_MAX];
  if (mmr_gemm::lu_poisoned(mmr_gemm::lu_failword(fail), k0)) return;  // the panel [k0, k0+nb) was never committed
  const int lane = threadIdx.x;
  int any = 0;
  for (int i = lane; i < nb; i += 32) {
    const int p = ipiv[k0 + i];
    piv[i] = p;
    rows[i] = k0 + i;
    cur[i] = k0 + i;
    any |= (p != k0 + i);
  }
  if (!__any_sync(0xffffffffu, any)) {  // no interchange in this panel: the usual demag case
    if (lane == 0) {
      out->count = 0;
      out->moved = 0;
    }
    return;
  }
  int cnt = nb;
  __syncwarp();
  for (int i = 0; i < nb; ++i) {
    const int p = piv[i];
    if (p == k0 + i) continue;  // warp-uniform
    int found = -1;
    for (int t = lane; t < cnt; t += 32)
      if (rows[t] == p) found = t;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) found = max(found, __shfl_xor_sync(0xffffffffu, found, off));
    if (found < 0) {
      found = cnt;
      if (lane == 0) {
        rows[cnt] = p;
        cur[cnt] = p;
      }
      ++cnt;
    }
    __syncwarp();
    if (lane == 0) {
      const int t = cur[i];
      cur[i] = cur[found];
      cur[found] = t;
    }
    __syncwarp();
  }
  // keep only the rows whose content changes (warp-level stream compaction)
  int kept = 0;
  for (int base = 0; base < cnt; base += 32) {
    const int t = base + lane;
    const bool mv = (t < cnt) && (rows[t] != cur[t]);
    const unsigned bal = __ballot_sync(0xffffffffu, mv);
    if (mv) {
      const int slot = kept + __popc(bal & ((1u << lane) - 1u));
      out->rows[slot] = rows[t];
      out->src[slot] = cur[t];
    }
    kept += __popc(bal);
  }
  if (lane == 0) {
    out->count = kept;
    out->moved = kept > 0;
  }
}

// Applies up to two prepared permutations (p0 first, then p1) to the columns [col_begin, col_end):
// one warp per column gathers the moved rows into registers and scatters them.
__global__ void __launch_bounds__(256)
    lu_laswp_apply_kernel(double* __restrict__ A, int64_t ld, int col_begin, int col_end,
                          const LaswpPerm* __restrict__ p0, const LaswpPerm* __restrict__ p1,
                          const int* __restrict__ fail, int dep) {
  if (mmr_gemm::lu_poisoned(mmr_gemm::lu_failword(fail), dep)) return;
  const int m0 = p0->moved, m1 = p1 ? p1->moved : 0;
  if ((m0 | m1) == 0) return;
  __shared__ int s_rows[LASWP_MAX], s_src[LASWP_MAX];
  const int lane = threadIdx.x & 31;
  const int col = col_begin + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  double* a = A + (int64_t)min(col, col_end - 1) * ld;
  for (int pass = 0; pass < 2; ++pass) {
    const LaswpPerm* perm = pass ? p1 : p0;
    if (!(pass ? m1 : m0)) continue;  // block-uniform
    const int cnt = perm->count;
    __syncthreads();
    for (int t = threadIdx.x; t < cnt; t += blockDim.x) {
      s_rows[t] = perm->rows[t];
      s_src[t] = perm->src[t];
    }
    __syncthreads();
    if (col < col_end) {
      double v[LASWP_MAX / 32];
#pragma unroll
      for (int q = 0; q < LASWP_MAX / 32; ++q) {
        const int t = lane + 32 * q;
        v[q] = (t < cnt) ? a[s_src[t]] : 0.0;
      }
      __syncwarp();  // every gather of this column is performed before any scatter
#pragma unroll
      for (int q = 0; q < LASWP_MAX / 32; ++q) {
        const int t = lane + 32 * q;
        if (t < cnt) a[s_rows[t]] = v[q];
      }
      __syncwarp();  // ... and the scatter before the second permutation's gather
    }
  }
}
}  // namespace

// Permutation record slots in the context: [0], [1] = scratch of mmr_lu_laswp (main / other
// stream), [2..5] = per-panel records of the single-GPU driver (outer-block parity x inner panel).
static int laswp_slots(mmr_ctx* ctx, LaswpPerm** out) {
  if (!ctx->laswp_buf) MMR_CUDA(cudaMalloc(&ctx->laswp_buf, 6 * sizeof(LaswpPerm)));
  *out = (LaswpPerm*)ctx->laswp_buf;
  return 0;
}

// Turns the interchanges ipiv[k0..k1) into the permutation record `slot` (0..5).
int mmr_lu_laswp_prepare(mmr_ctx* ctx, cudaStream_t strm, const int* d_ipiv, int k0, int k1, int slot,
                         const int* d_fail = nullptr) {
  LaswpPerm* perms;
  int rc = laswp_slots(ctx, &perms);
  if (rc) return rc;
  lu_laswp_prepare_kernel<<<1, 32, 0, strm>>>(d_ipiv, k0, k1 - k0, perms + slot, d_fail);
  MMR_CHECK_LAUNCH(ctx);
  return 0;
}

// Applies record slot0 and then (slot1 >= 0) slot1 to the local columns [col_begin, col_end).
int mmr_lu_laswp_apply(mmr_ctx* ctx, cudaStream_t strm, double* A, int64_t ld, int col_begin, int col_end, int slot0,
                       int slot1, const int* d_fail = nullptr, int dep = 0) {
  if (col_end <= col_begin) return 0;
  LaswpPerm* perms;
  int rc = laswp_slots(ctx, &perms);
  if (rc) return rc;
  const int warps = 8;
  lu_laswp_apply_kernel<<<(unsigned)ceil_div64(col_end - col_begin, warps), warps * 32, 0, strm>>>(
      A, ld, col_begin, col_end, perms + slot0, slot1 >= 0 ? perms + slot1 : nullptr, d_fail, dep);
  MMR_CHECK_LAUNCH(ctx);
  return 0;
}

// Applies the interchanges ipiv[k0..k1) to local columns [col_begin, col_end) (prepare + apply;
// one scratch record per stream: calls on one stream are ordered).
int mmr_lu_laswp(mmr_ctx* ctx, cudaStream_t strm, double* A, int64_t ld, int k0, int k1, int col_begin, int col_end,
                 const int* d_ipiv, const int* d_fail = nullptr) {
  if (col_end <= col_begin || k1 <= k0) return 0;
  const int slot = (strm == ctx->stream) ? 0 : 1;
  int rc = mmr_lu_laswp_prepare(ctx, strm, d_ipiv, k0, k1, slot, d_fail);
  if (rc) return rc;
  return mmr_lu_laswp_apply(ctx, strm, A, ld, col_begin, col_end, slot, -1, d_fail, k0);
}

int mmr_lu_trsm(mmr_ctx* ctx, cudaStream_t strm, const double* L, int64_t ldl, int nb, double* B, int64_t ldb, int ncols) {
  if (ncols <= 0 || nb <= 0) return 0;
  const size_t smem = ((size_t)TRSM_COLS * (nb + 1) + (size_t)nb * (nb + 1)) * sizeof(double);
  {
    const int rc0 = mmr_func_smem(ctx, (const void*)lu_trsm_kernel, smem);
    if (rc0) return rc0;
  }
  lu_trsm_kernel<<<(unsigned)ceil_div64(ncols, TRSM_COLS), 256, smem, strm>>>(L, ldl, nb, B, ldb, ncols);
  MMR_CHECK_LAUNCH(ctx);
  return 0;
}

namespace {
__global__ void lu_set_identity_kernel(double* __restrict__ X, int nb) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < nb * nb) X[e] = (e % nb == e / nb) ? 1.0 : 0.0;
}
}  // namespace

// X = L11^-1 as the solution of L11 X = I with the blocked TRSM kernel (nb/32 CTAs, L11 in shared
// memory).  (The first version used a one-CTA column-per-thread substitution: 304 us per panel.)
int mmr_lu_trtri(mmr_ctx* ctx, cudaStream_t strm, const double* L, int64_t ldl, int nb, double* X) {
  lu_set_identity_kernel<<<(unsigned)ceil_div64((int64_t)nb * nb, 256), 256, 0, strm>>>(X, nb);
  MMR_CHECK_LAUNCH(ctx);
  return mmr_lu_trsm(ctx, strm, L, ldl, nb, X, nb, nb);
}

// =======================================================================================
// "Inverse format" of a fully speculative factorization and the block sweeps that use it.
//
// When no panel was rejected, the factorization ends by overwriting every diagonal panel block L11\U11 (<= 128
// wide) with L11^-1 (strictly lower part, unit diagonal implied) and U11^-1 (upper part) -- both were computed
// for the panel anyway -- and sets LU_FORMAT_INVDIAG in ipiv[0].  The triangular solves then need no
// substitution at all: a block of the solution is a pair of small matrix-vector products, and one launch per
// outer block does (i) the update of every later (earlier) column with the block that just became final and
// (ii), in CTA 0, the finishing of the next block.  One launch and -- on several GPUs -- one small broadcast
// per block instead of a dependent gemv -> diagonal-solve -> gemv -> diagonal-solve chain.
// Factorizations that needed interchanges keep the plain format and the substitution kernels above (an
// explicit inverse of a pivoted U11 would cost backward stability).
// =======================================================================================
constexpr int LU_FORMAT_INVDIAG = 1 << 30;
constexpr int LU_IPIV_MASK = LU_FORMAT_INVDIAG - 1;

namespace {

// panel slot q of the inverse stash: [L^-1 (nb x nb, ld nb) | U^-1 (nb x nb, ld nb)], 2 * NB * NB doubles
__global__ void __launch_bounds__(256)
    lu_store_inverses_kernel(double* __restrict__ Al, int64_t ld, const double* __restrict__ stash, int npanels,
                             int N, int obw, int R, int me, const int* __restrict__ fail) {
  if (*fail != 0) return;  // a panel was rejected: the factors stay in plain format
  const int q = blockIdx.x;
  if (q >= npanels) return;
  // panel q of this rank: local block j = q / 2 (global block me + j R), inner panel q % 2
  const int j = q >> 1, inner = q & 1;
  const int t = me + j * R;
  const int kb0 = t * obw;
  if (kb0 >= N) return;
  const int ob = min(obw, N - kb0), nb0 = min(NB, ob);
  const int nb = inner ? ob - nb0 : nb0;
  if (nb <= 0) return;
  const int off = inner ? nb0 : 0;  // offset of the panel inside its block
  const double* Li = stash + (size_t)q * 2 * NB * NB;
  const double* Ui = Li + (size_t)NB * NB;
  double* D = Al + (int64_t)(j * obw + off) * ld + kb0 + off;  // local column of the panel, global row
  for (int e = threadIdx.x; e < nb * nb; e += blockDim.x) {
    const int r = e % nb, c = e / nb;
    D[(int64_t)c * ld + r] = (r > c) ? Li[c * nb + r] : Ui[c * nb + r];
  }
}

__global__ void lu_set_format_kernel(int* __restrict__ ipiv, const int* __restrict__ fail) {
  if (*fail == 0) ipiv[0] |= LU_FORMAT_INVDIAG;
}

// One step of the block sweeps on inverse-format factors (see above).  Global column c of block t = c / obw lives
// at local column (t / R) * obw + c % obw of the owner t % R; rows are global.
//   s_done : block whose piece of b is final (-1: none)      s_next : block CTA 0 finishes now (-1: none; owned)
//   [far_lo, far_hi) : my local columns that receive the update by block s_done (CTAs 1..)
//   peers  : NULL, or the table of every rank's exchange buffer (mmr_comm_p2p_*): finished blocks travel as
//            direct stores into all peers' buffers + one flag word per block (value = `epoch` of this sweep)
//            instead of an ncclBroadcast; a consumer spins on its own copy of the flag.  *p2p_err is set if a
//            flag does not arrive within ~2 s (a peer died): the sweep then finishes with garbage, never hangs.
constexpr int SWEEP_THREADS = 1024, SWEEP_WARPS = SWEEP_THREADS / 32, SWEEP_NC = 2 * NB / SWEEP_WARPS;
template <bool FWD>
__global__ void __launch_bounds__(SWEEP_THREADS)
    lu_sweep_kernel(const double* __restrict__ Al, int64_t ld, int N, int obw, int R, int me, int s_done, int s_next,
                    int far_lo, int far_hi, double* __restrict__ b, void* const* __restrict__ peers, int epoch,
                    int* __restrict__ p2p_err) {
  __shared__ double xs[2 * NB], ys[2 * NB], zs[2 * NB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int u0 = s_done >= 0 ? s_done * obw : 0;
  const int nu = s_done >= 0 ? min(obw, N - u0) : 0;
  const double* xsrc = b;  // where the finished block s_done is read from
  if (peers) {
    double* mine = (double*)peers[me];
    xsrc = mine + (FWD ? 0 : MMR_P2P_CAP);
    if (s_done >= 0 && s_done % R != me) {
      if (tid == 0) {
        const volatile int* flag = (const volatile int*)(mine + 2 * MMR_P2P_CAP) + (FWD ? 0 : MMR_P2P_FLAGS) + s_done;
        const long long t0 = clock64();
        while (*flag != epoch) {
          if (clock64() - t0 > (1LL << 32)) {
            atomicExch(p2p_err, 1);
            break;
          }
        }
        __threadfence_system();
      }
      __syncthreads();
    }
  }
  if (blockIdx.x != 0) {
    // far columns: a warp takes four columns per pass -- 32 independent loads per lane in flight before the
    // first reduction (one column per short-lived warp kept the memory pipeline at ~10 % of its bandwidth)
    const int nwarps = (gridDim.x - 1) * SWEEP_WARPS;
    for (int base = far_lo + 4 * ((blockIdx.x - 1) * SWEEP_WARPS + warp); base < far_hi; base += 4 * nwarps) {
      double acc[4] = {0.0, 0.0, 0.0, 0.0};
      // unconditional loads from clamped addresses, sixteen issued before the first use (see `dots` below)
      const double* cp[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) cp[c] = Al + (int64_t)min(base + c, far_hi - 1) * ld + u0;
#pragma unroll
      for (int t = 0; t < 8; t += 4) {
        double a16[4][4], xv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int i = min(lane + 32 * (t + q), nu - 1);
          xv[q] = (lane + 32 * (t + q) < nu) ? xsrc[u0 + i] : 0.0;
#pragma unroll
          for (int c = 0; c < 4; ++c) a16[q][c] = __ldg(cp[c] + i);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[c] = fma(a16[q][c], xv[q], acc[c]);
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], off);
      if (lane < 4 && base + lane < far_hi) {
        const int lc = base + lane;
        const int g = ((lc / obw) * R + me) * obw + lc % obw;
        const double a = lane == 0 ? acc[0] : (lane == 1 ? acc[1] : (lane == 2 ? acc[2] : acc[3]));
        if (g < N) b[g] -= a;
      }
    }
    return;
  }
  if (s_next < 0) return;
  const int d0 = s_next * obw;
  const int nd = min(obw, N - d0), na = min(NB, nd);
  const double* blk = Al + (int64_t)((s_next / R) * obw) * ld;  // local column of global column d0
  for (int i = tid; i < nu; i += blockDim.x) xs[i] = xsrc[u0 + i];
  __syncthreads();
  // Warp w owns the columns j = w + SWEEP_WARPS c, c < SWEEP_NC, of the block.  A product phase is latency bound (a handful of
  // DRAM round trips), so every lane first issues the loads of ALL eight columns and only then reduces: one
  // memory latency per phase instead of one per column.
  //   sums[c] = sum_{r in [lo(j), hi(j))} A[row0 + r, d0 + j] * v[r]
  auto dots = [&](int row0, int nrow, auto lo_of, auto hi_of, const double* v, double (&sums)[SWEEP_NC]) {
    double acc[SWEEP_NC];
#pragma unroll
    for (int c = 0; c < SWEEP_NC; ++c) acc[c] = 0.0;
    if (nrow > 0) {
      const int64_t cstride = (int64_t)SWEEP_WARPS * ld;
      // The loads are UNCONDITIONAL (rows / columns out of range read a clamped, valid address and are masked
      // afterwards) and eight of them are issued before the first use: with predicated load -> fma pairs the
      // in-order issue serialised them, one L2 round trip (~400 cycles) per load, 25 000 cycles per phase.
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int r = lane + 32 * t;
        const double vr = v[r];  // (v has 256 readable entries)
        const double* rowp = blk + row0 + min(r, nrow - 1);
#pragma unroll
        for (int h = 0; h < SWEEP_NC; h += 8) {
          double a8[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) a8[c] = __ldg(rowp + (int64_t)min(warp + SWEEP_WARPS * (h + c), nd - 1) * ld);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const int j = warp + SWEEP_WARPS * (h + c);
            const bool ok = j < nd && r >= lo_of(j) && r < hi_of(j);
            acc[h + c] = fma(ok ? a8[c] : 0.0, vr, acc[h + c]);
          }
        }
      }
      (void)cstride;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1)
#pragma unroll
        for (int c = 0; c < SWEEP_NC; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], off);
    }
#pragma unroll
    for (int c = 0; c < SWEEP_NC; ++c) sums[c] = acc[c];
  };
  double sm[SWEEP_NC];
  // (1) the update by block s_done
  dots(u0, nu, [&](int) { return 0; }, [&](int) { return nu; }, xs, sm);
#pragma unroll
  for (int c = 0; c < SWEEP_NC; ++c) {
    const int j = warp + SWEEP_WARPS * c;
    if (lane == 0 && j < nd) ys[j] = b[d0 + j] - sm[c];
  }
  __syncthreads();
  // (2) the block itself: two panels a = [0, na), q = [na, nd) with inverted diagonal blocks
  if (FWD) {
    // w_a = U_aa^-T y_a   and, for the columns of q, the part of U_qq^-T y_q ... no: y_q is not final yet
    dots(d0, nd, [&](int) { return 0; }, [&](int j) { return j < na ? j + 1 : 0; }, ys, sm);
#pragma unroll
    for (int c = 0; c < SWEEP_NC; ++c) {
      const int j = warp + SWEEP_WARPS * c;
      if (lane == 0 && j < na) zs[j] = sm[c];
    }
    __syncthreads();
    dots(d0, nd, [&](int) { return 0; }, [&](int j) { return j >= na ? na : 0; }, zs, sm);  // y_q -= U_aq^T w_a
#pragma unroll
    for (int c = 0; c < SWEEP_NC; ++c) {
      const int j = warp + SWEEP_WARPS * c;
      if (lane == 0 && j >= na && j < nd) ys[j] -= sm[c];
    }
    __syncthreads();
    dots(d0, nd, [&](int) { return na; }, [&](int j) { return j >= na ? j + 1 : 0; }, ys, sm);  // w_q = U_qq^-T y_q
#pragma unroll
    for (int c = 0; c < SWEEP_NC; ++c) {
      const int j = warp + SWEEP_WARPS * c;
      if (lane == 0 && j >= na && j < nd) zs[j] = sm[c];
    }
  } else {
    dots(d0, nd, [&](int j) { return j + 1; }, [&](int j) { return j >= na ? nd : 0; }, ys, sm);  // v_q = L_qq^-T y_q
#pragma unroll
    for (int c = 0; c < SWEEP_NC; ++c) {
      const int j = warp + SWEEP_WARPS * c;
      if (lane == 0 && j >= na && j < nd) zs[j] = ys[j] + sm[c];
    }
    __syncthreads();
    dots(d0, nd, [&](int) { return na; }, [&](int j) { return j < na ? nd : 0; }, zs, sm);  // y_a -= L_qa^T v_q
#pragma unroll
    for (int c = 0; c < SWEEP_NC; ++c) {
      const int j = warp + SWEEP_WARPS * c;
      if (lane == 0 && j < na) ys[j] -= sm[c];
    }
    __syncthreads();
    dots(d0, nd, [&](int j) { return j + 1; }, [&](int j) { return j < na ? na : 0; }, ys, sm);  // v_a = L_aa^-T y_a
#pragma unroll
    for (int c = 0; c < SWEEP_NC; ++c) {
      const int j = warp + SWEEP_WARPS * c;
      if (lane == 0 && j < na) zs[j] = ys[j] + sm[c];
    }
  }
  __syncthreads();
  for (int j = tid; j < nd; j += blockDim.x) b[d0 + j] = zs[j];
  if (peers) {
    // publish the finished block: data into every rank's buffer (mine included), then the flags
    for (int e = tid; e < R * nd; e += blockDim.x) {
      const int r = e / nd, j = e - r * nd;
      ((double*)peers[r] + (FWD ? 0 : MMR_P2P_CAP))[d0 + j] = zs[j];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < R) {
      volatile int* flag = (volatile int*)((double*)peers[tid] + 2 * MMR_P2P_CAP) + (FWD ? 0 : MMR_P2P_FLAGS) + s_next;
      *flag = epoch;
    }
  }
}

// every block flag of a sweep has reached this rank (value == epoch); gives up after ~2 s
__global__ void lu_p2p_wait_all_kernel(const int* flags, int S, int epoch, int* __restrict__ err) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= S) return;
  const volatile int* f = flags + t;
  const long long t0 = clock64();
  while (*f != epoch) {
    if (clock64() - t0 > (1LL << 32)) {
      atomicExch(err, 1);
      break;
    }
  }
  __threadfence_system();
}

}  // namespace

// Triangular solves U^T w = b, L^T v = w on inverse-format factors (LU_FORMAT_INVDIAG), in place on d_b.
// obw: width of the two-panel blocks the factorization used (256 on one GPU, 3 * tgt_block distributed); R, me:
// block-cyclic ownership (1, 0 on one GPU); bcast: broadcast every finished block (R > 1).
static int lu_sweeps_inv(mmr_ctx* ctx, cudaStream_t st, int N, int obw, int R, int me, const double* Al, int64_t ld,
                         int ncols_local, double* d_b) {
  const int S = (int)ceil_div64(N, obw);
  // peer-memory exchange (one NVSwitch box, mmr_comm_p2p_*) instead of one ncclBroadcast per block
  const bool p2p = R > 1 && ctx->p2p_ready && N <= MMR_P2P_CAP && S <= MMR_P2P_FLAGS;
  void* const* peers = p2p ? (void* const*)ctx->p2p_dev_table : nullptr;
  int epoch = 0;
  if (p2p) {
    const int rcw = mmr_ws_reserve(ctx, 256);
    if (rcw) return rcw;
  }
  int* p2p_err = (int*)ctx->ws + 2;  // (ws[0..1]: info / guard words of the factorization)
  if (p2p) {
    MMR_CUDA(cudaMemsetAsync(p2p_err, 0, sizeof(int), st));
    epoch = ++ctx->p2p_epoch;
  }
  auto step = [&](bool fwd, int s_done, int s_next_global) -> int {
    // my columns that receive the update by s_done, and whether I finish s_next
    const bool own_next = s_next_global >= 0 && s_next_global % R == me;
    int far_lo = 0, far_hi = 0;
    if (s_done >= 0) {
      if (fwd) {
        int jlo = (s_done >= me) ? (s_done - me) / R + 1 : 0;  // my blocks with global index <= s_done
        if (own_next) ++jlo;                                  // CTA 0 takes that one
        far_lo = jlo * obw, far_hi = ncols_local;
      } else {
        int nleft = (s_done > me) ? (s_done - 1 - me) / R + 1 : 0;  // my blocks with global index < s_done
        if (own_next) --nleft;
        far_lo = 0, far_hi = nleft * obw;
      }
    }
    const int nfar = std::max(0, far_hi - far_lo);
    if (!own_next && nfar == 0) return 0;
    // CTA 0 + up to one far CTA per remaining SM (4 columns per warp and pass)
    const unsigned grid = 1 + (unsigned)std::min<int64_t>(ctx->sm_count - 1, ceil_div64(nfar, 4 * SWEEP_WARPS));
    if (fwd)
      lu_sweep_kernel<true><<<grid, SWEEP_THREADS, 0, st>>>(Al, ld, N, obw, R, me, s_done, own_next ? s_next_global : -1, far_lo,
                                                  far_hi, d_b, peers, epoch, p2p_err);
    else
      lu_sweep_kernel<false><<<grid, SWEEP_THREADS, 0, st>>>(Al, ld, N, obw, R, me, s_done, own_next ? s_next_global : -1, far_lo,
                                                   far_hi, d_b, peers, epoch, p2p_err);
    MMR_CHECK_LAUNCH(ctx);
    return 0;
  };
  auto bcast = [&](int s) -> int {
    if (R <= 1 || p2p) return 0;
    const int kb = s * obw, ob = std::min(obw, N - kb);
    const int pc = mmr_prof_begin(ctx, MMR_PROF_COMM, st);
    const int r = mmr_comm_bcast_bytes(ctx, d_b + kb, sizeof(double) * ob, s % R, st);
    mmr_prof_end(ctx, pc, 8.0 * ob, st);
    return r;
  };
  int rc = 0;
  if ((rc = step(true, -1, 0)) || (rc = bcast(0))) return rc;
  for (int s = 0; s + 1 < S; ++s)
    if ((rc = step(true, s, s + 1)) || (rc = bcast(s + 1))) return rc;
  if (p2p) epoch = ++ctx->p2p_epoch;
  if ((rc = step(false, -1, S - 1)) || (rc = bcast(S - 1))) return rc;
  for (int s = S - 1; s >= 1; --s)
    if ((rc = step(false, s, s - 1)) || (rc = bcast(s - 1))) return rc;
  if (p2p) {
    // the solution is complete in my exchange buffer once the flags of ALL blocks are here (they come from
    // different peers: the flag of block 0 says nothing about the data of block 1); then buffer -> d_b
    lu_p2p_wait_all_kernel<<<(unsigned)ceil_div64(S, 256), 256, 0, st>>>(
        (const int*)((double*)ctx->p2p_local + 2 * MMR_P2P_CAP) + MMR_P2P_FLAGS, S, epoch, p2p_err);
    MMR_CHECK_LAUNCH(ctx);
    MMR_CUDA(cudaMemcpyAsync(d_b, (double*)ctx->p2p_local + MMR_P2P_CAP, sizeof(double) * N, cudaMemcpyDeviceToDevice, st));
    // (the callers keep the pivots and the permutation in the first 2 N ints of the pinned buffer and reserve
    // 256 bytes more)
    int* h_err = (int*)((char*)ctx->hpin + 2 * sizeof(int) * (size_t)N + 128);
    MMR_CUDA(cudaMemcpyAsync(h_err, p2p_err, sizeof(int), cudaMemcpyDeviceToHost, st));
    MMR_CUDA(cudaStreamSynchronize(st));
    if (*h_err) {
      mmr_set_error("peer-memory sweep: a block flag did not arrive within the time-out (a peer rank failed?)");
      return -7;
    }
  }
  return 0;
}

// width of the outer blocks of the single-GPU factorization (two 128-column panels unless MMR_LU_OUTER=1)
static int lu_outer_block() {
  static int outer_mult = -1;
  if (outer_mult < 0) {
    const char* e = getenv("MMR_LU_OUTER");  // 1: rank-128 updates (one panel per outer block), 2: rank-256
    outer_mult = (e && e[0] == '1') ? 1 : 2;
  }
  return outer_mult * NB;
}

// Single-GPU driver.  Two-level blocking: panels of NB = 128 columns are factored in pairs (an
// "outer block" of OB = 256 columns), and the trailing matrix receives ONE rank-256 DMMA update per
// outer block instead of two rank-128 updates: half the C read-modify-write traffic and half the
// per-tile prologue/epilogue (measured on B200, m = n = 12288: 29.7 TF at k = 128, 32.3 TF at
// k = 256).  The 256 x 256 unit-lower solve for U12 is done in two 128-row steps with the two
// inverted diagonal blocks.
extern "C" int mmr_lu_factor(mmr_ctx* ctx, int64_t N64, double* d_Q, int64_t ld, int32_t* d_ipiv,
                             int* info_out) {
  MMR_ARG(ctx != nullptr, "ctx is NULL");
  MMR_ARG(N64 >= 0 && N64 < (1LL << 31), "N out of range");
  MMR_ARG(ld >= N64, "ld must be >= N");
  if (info_out) *info_out = 0;
  if (N64 == 0) return 0;
  MMR_ARG(d_Q && d_ipiv, "NULL matrix/pivot pointer");
  DeviceGuard g(ctx->device);
  const int N = (int)N64;
  const int OB = lu_outer_block();
  static int spec_default = -1;
  if (spec_default < 0) {
    const char* e = getenv("MMR_LU_SPEC");  // 0: always the pivoting strip kernels
    spec_default = (e && e[0] == '0') ? 0 : 1;
  }
  bool spec_on = spec_default != 0;  // switched off for the rest of this factorization after a rejected panel
  const size_t scratch_bytes = (mmr_lu_panel_scratch_bytes() + 255) / 256 * 256;
  const size_t spec_bytes = (mmr_lu_spec_scratch_bytes() + 255) / 256 * 256;
  const size_t sbuf_bytes = spec_on ? ((sizeof(double) * (size_t)N * NB + 255) / 256 * 256) : 0;
  // inverse stash: L11^-1 | U11^-1 of every panel, written into the diagonal blocks at the end if no panel was
  // rejected (inverse format, see lu_sweep_kernel)
  const int npanel_slots = 2 * (int)ceil_div64(N, OB);
  const size_t stash_bytes = spec_on ? ((sizeof(double) * (size_t)npanel_slots * 2 * NB * NB + 255) / 256 * 256) : 0;
  // deferred-TRSM path (see run()): two W = L21 L_blk^-1 buffers (N x OB) and one N x NB temporary
  static int defer_default = -1;
  if (defer_default < 0) {
    const char* e = getenv("MMR_LU_DEFER_TRSM");  // 0: U12 = L_blk^-1 A12 before every trailing update (round 1)
    defer_default = (e && e[0] == '0') ? 0 : 1;
  }
  const bool defer_trsm = spec_on && ctx->lookahead && defer_default && OB == 2 * NB && N > 4 * OB;
  const size_t wbuf_bytes = defer_trsm ? ((sizeof(double) * (size_t)N * OB + 255) / 256 * 256) : 0;
  const size_t tbuf_bytes = defer_trsm ? ((sizeof(double) * (size_t)N * NB + 255) / 256 * 256) : 0;
  int rc = mmr_ws_reserve(ctx, 256 + scratch_bytes + 8 * sizeof(double) * NB * NB + spec_bytes + sbuf_bytes + stash_bytes +
                                   2 * wbuf_bytes + tbuf_bytes);
  if (rc) return rc;
  rc = mmr_hpin_reserve(ctx, 256);
  if (rc) return rc;
  int* d_info = (int*)ctx->ws;  // [0] zero-pivot report, [1] speculation guard (lu_poisoned)
  int* d_fail = d_info + 1;
  void* scratch = (char*)ctx->ws + 256;
  double* Lbuf = (double*)((char*)ctx->ws + 256 + scratch_bytes);
  void* d_spec = (char*)Lbuf + 8 * sizeof(double) * NB * NB;
  double* Sbuf = (double*)((char*)d_spec + spec_bytes);
  double* stash = (double*)((char*)Sbuf + sbuf_bytes);
  double* Wbuf[2] = {(double*)((char*)stash + stash_bytes), (double*)((char*)stash + stash_bytes + wbuf_bytes)};
  double* Tbuf = (double*)((char*)stash + stash_bytes + 2 * wbuf_bytes);
  auto slot_of = [&](int kb) { return stash + (size_t)(2 * (kb / OB) + ((kb % OB) ? 1 : 0)) * 2 * NB * NB; };
  // inverse of the unit-lower diagonal block of inner panel `inner` of the outer block at kb
  // (four generations of the pair: the deferred U12 work of block kb - 2 OB still reads its inverses while block
  // kb + OB is being factored -- with two generations the factorization had to wait for it)
  auto Linv = [&](int kb, int inner) { return Lbuf + (size_t)((((kb / OB) & 3) << 1) + inner) * NB * NB; };
  MMR_CUDA(cudaMemsetAsync(d_info, 0, 2 * sizeof(int), ctx->stream));
  double* A = d_Q;
  cudaStream_t main_s = ctx->stream;
  cudaStream_t side_s = ctx->side;
  // start of the last panel of the outer block at kb: what a consumer of the whole block depends on
  auto last_panel = [&](int kb) { return (min(OB, N - kb) > NB) ? kb + NB : kb; };
  // interchange records: slot(kb, inner), prepared once per panel on the stream that factored it
  auto pslot = [&](int kb, int inner) { return 2 + ((((kb / OB) & 1) << 1) + inner); };
  // (speculative mode: a committed panel has identity pivots and a rejected one poisons its consumers, so
  // the interchange kernels would be no-ops either way and are not launched at all)
  auto laswp = [&](cudaStream_t s, int cb, int ce, int slot0, int slot1, int dep) -> int {
    if (ce <= cb || spec_on) return 0;
    const int ps = mmr_prof_begin(ctx, MMR_PROF_LASWP, s);
    const int r = mmr_lu_laswp_apply(ctx, s, A, ld, cb, ce, slot0, slot1, d_fail, dep);
    mmr_prof_end(ctx, ps, 32.0 * NB * (double)(ce - cb), s);
    return r;
  };
  // B <- Linv * B in place on rows [r0, r0+nb) of columns [cb, ce): one m-tile, so each CTA has its
  // whole B tile in shared memory before it stores and no other CTA touches those columns.
  auto trsm = [&](cudaStream_t s, const double* Li, int r0, int nb, int cb, int ce, int dep) -> int {
    const int ps = mmr_prof_begin(ctx, MMR_PROF_TRSM, s);
    const int r = mmr_dgemm_launch(ctx, s, nb, ce - cb, nb, 1.0, Li, nb, A + (int64_t)cb * ld + r0, ld, 0.0,
                                   A + (int64_t)cb * ld + r0, ld, d_fail, dep);
    mmr_prof_end(ctx, ps, 2.0 * nb * nb * (double)(ce - cb), s);
    return r;
  };
  // A[r0:r1, cb:ce] -= A[r0:r1, k0:k0+k] * A[k0:k0+k, cb:ce]
  auto gemm = [&](cudaStream_t s, int kind, int r0, int r1, int k0, int k, int cb, int ce, int dep) -> int {
    if (r1 <= r0 || ce <= cb) return 0;
    const int ps = mmr_prof_begin(ctx, kind, s);
    const int r = mmr_dgemm_launch(ctx, s, r1 - r0, ce - cb, k, -1.0, A + (int64_t)k0 * ld + r0, ld,
                                   A + (int64_t)cb * ld + k0, ld, 1.0, A + (int64_t)cb * ld + r0, ld, d_fail, dep);
    mmr_prof_end(ctx, ps, 2.0 * (r1 - r0) * (double)(ce - cb) * k, s);
    return r;
  };
  // panel [kb, kb+nb) and the inverse of its unit-lower diagonal block: speculative (diagonal pivots,
  // verified on the device, stream-ordered) or the pivoting strip kernels
  auto panel = [&](cudaStream_t s, int kb, int nb, double* Linv_out) -> int {
    if (spec_on) {
      // U11^-1 goes straight into the panel's stash slot, L11^-1 is copied there behind the panel
      double* slot = slot_of(kb);
      const int r = mmr_lu_panel_spec(ctx, s, A, ld, N, kb, nb, d_ipiv, Sbuf, N - kb, Linv_out, slot + (size_t)NB * NB, d_fail, true);
      if (r) return r;
      MMR_CUDA(cudaMemcpyAsync(slot, Linv_out, sizeof(double) * nb * nb, cudaMemcpyDeviceToDevice, s));
      return 0;
    }
    const int r = mmr_lu_panel(ctx, s, A, ld, N, kb, nb, d_ipiv, d_info, scratch);
    if (r) return r;
    return mmr_lu_trtri(ctx, s, A + (int64_t)kb * ld + kb, ld, nb, Linv_out);
  };
  // factor the outer block [kb, kb+ob): panel 0, its update of panel 1's columns, panel 1
  // (first_inner = 1: panel 0 and its update of panel 1's columns are already done -- resuming after a
  // rejected panel 1)
  auto factor_outer = [&](cudaStream_t s, int kb, int ob, int first_inner) -> int {
    const int nb0 = min(NB, ob);
    int r = 0;
    if (first_inner == 0) {
      if ((r = panel(s, kb, nb0, Linv(kb, 0)))) return r;
      if (!spec_on && (r = mmr_lu_laswp_prepare(ctx, s, d_ipiv, kb, kb + nb0, pslot(kb, 0), d_fail))) return r;
    }
    if (first_inner != 0 && (r = mmr_lu_laswp_prepare(ctx, s, d_ipiv, kb, kb + nb0, pslot(kb, 0), d_fail))) return r;
    if (ob <= NB) return 0;
    const int k1 = kb + NB, nb1 = ob - NB;
    if (first_inner == 0) {
      if ((r = laswp(s, k1, k1 + nb1, pslot(kb, 0), -1, kb))) return r;
      if ((r = trsm(s, Linv(kb, 0), kb, NB, k1, k1 + nb1, kb))) return r;
      if ((r = gemm(s, MMR_PROF_PANEL_GEMM, k1, N, kb, NB, k1, k1 + nb1, kb))) return r;
    }
    if ((r = panel(s, k1, nb1, Linv(kb, 1)))) return r;
    if (!spec_on && (r = mmr_lu_laswp_prepare(ctx, s, d_ipiv, k1, k1 + nb1, pslot(kb, 1), d_fail))) return r;
    return laswp(s, kb, k1, pslot(kb, 1), -1, k1);  // panel 1's interchanges on panel 0's columns
  };
  // trailing update of the column range [cb, ce) by the outer block [kb, kb+ob)
  auto update_cols_on = [&](cudaStream_t us, int kind_big, int kb, int ob, int cb, int ce) -> int {
    if (ce <= cb) return 0;
    const int nb0 = min(NB, ob), nb1 = ob - nb0;
    const int dep = last_panel(kb);
    int r = laswp(us, cb, ce, pslot(kb, 0), nb1 > 0 ? pslot(kb, 1) : -1, dep);
    if (r) return r;
    // narrow column ranges (the look-ahead update): the three U12 steps in one launch
    bool fused = false;
    if (nb1 > 0 && ce - cb <= BLOCK_TRSM_MAX_COLS) {
      r = mmr_block_trsm_launch(ctx, us, nb0, nb1, Linv(kb, 0), Linv(kb, 1), A + (int64_t)kb * ld + kb + nb0, ld,
                                A + (int64_t)cb * ld + kb, ld, ce - cb, d_fail, dep);
      if (r == 0) fused = true;
      else if (r != -2) return r;
    }
    if (!fused) {
      if ((r = trsm(us, Linv(kb, 0), kb, nb0, cb, ce, dep))) return r;
      if (nb1 > 0) {
        if ((r = gemm(us, MMR_PROF_TRSM, kb + nb0, kb + ob, kb, nb0, cb, ce, dep))) return r;
        if ((r = trsm(us, Linv(kb, 1), kb + nb0, nb1, cb, ce, dep))) return r;
      }
    }
    return gemm(us, kind_big, kb + ob, N, kb, ob, cb, ce, dep);
  };
  auto update_cols = [&](int kb, int ob, int cb, int ce) -> int { return update_cols_on(main_s, MMR_PROF_GEMM, kb, ob, cb, ce); };
  // Look-ahead (depth 1): as soon as the columns of outer block k+1 have received the update of
  // block k, block k+1 is factored on the high-priority side stream while the main stream applies
  // the rest of update k.  The two touch disjoint column ranges.
  const bool lookahead = ctx->lookahead;
  // One pass over the outer blocks from kb0 on.  Entry state: everything left of kb0 is factored and applied,
  // block kb0 has received all earlier updates (and, with first_inner = 1, its panel 0 is done).
  // Deferred TRSM (speculative mode, full two-panel blocks).  The trailing update needs U12 = L_blk^-1 A12, three
  // small dependent launches (L00^-1, the L10 coupling, L11^-1) that sat on the main stream in front of every
  // big GEMM: 17 ms of a 309 ms factorization at N = 24000, at 6.5 TFLOP/s.  By associativity
  //     A22 -= L21 (L_blk^-1 A12) = (L21 L_blk^-1) A12 = W A12,      W L_blk = L21,
  // so the big GEMM can run on the UNTOUCHED rows A12 with W in place of L21; W (three m x 128 x 128 products)
  // is built on the communication stream while the main stream does the narrow look-ahead update, and the
  // U12 rows themselves -- needed by nobody before the final solve -- are finalised afterwards on a
  // low-priority stream, concurrently with the next block's GEMM.
  cudaStream_t cm_s = ctx->comm, aux_s = ctx->aux;
  cudaEvent_t e_it = ctx->evd[0], e_W[2] = {ctx->evd[1], ctx->evd[2]}, e_bulk[2] = {ctx->evd[3], ctx->evd[4]},
              e_trsm[4] = {ctx->evd[5], ctx->evd[6], ctx->evd[8], ctx->evd[9]}, e_join = ctx->evd[7];
  bool par_used[2] = {false, false};
  bool trsm_used[4] = {false, false, false, false};  // e_trsm[g]: the deferred U12 work that read inverse generation g
  auto build_W = [&](int kb, int par) -> int {  // on cm_s; block [kb, kb + 2 NB) is factored and committed
    const int m = N - kb - OB, mW = m, dep = kb + NB;
    double* W = Wbuf[par];
    const double* L21a = A + (int64_t)kb * ld + kb + OB;
    const double* L21b = A + (int64_t)(kb + NB) * ld + kb + OB;
    const double* L10 = A + (int64_t)kb * ld + kb + NB;
    int ps = mmr_prof_begin(ctx, MMR_PROF_PANEL_GEMM, cm_s);
    int r = mmr_dgemm_launch(ctx, cm_s, m, NB, NB, 1.0, L21b, ld, Linv(kb, 1), NB, 0.0, W + (size_t)NB * mW, mW, d_fail, dep);
    if (r) return r;
    {
      const unsigned g = (unsigned)std::min<int64_t>((int64_t)4 * ctx->sm_count, ceil_div64((int64_t)m * NB, 256 * 8));
      lu_copy_guarded_kernel<<<g, 256, 0, cm_s>>>(L21a, ld, m, NB, Tbuf, mW, d_fail, dep);
      MMR_CHECK_LAUNCH(ctx);
    }
    if ((r = mmr_dgemm_launch(ctx, cm_s, m, NB, NB, -1.0, W + (size_t)NB * mW, mW, L10, ld, 1.0, Tbuf, mW, d_fail, dep))) return r;
    if ((r = mmr_dgemm_launch(ctx, cm_s, m, NB, NB, 1.0, Tbuf, mW, Linv(kb, 0), NB, 0.0, W, mW, d_fail, dep))) return r;
    mmr_prof_end(ctx, ps, 6.0 * m * (double)NB * NB, cm_s);
    return 0;
  };
  // U12 rows of the columns [cb, ce): the first three steps of update_cols, on stream s
  auto finish_U12 = [&](cudaStream_t s, int kb, int cb, int ce) -> int {
    const int dep = kb + NB;
    int r = trsm(s, Linv(kb, 0), kb, NB, cb, ce, dep);
    if (r) return r;
    if ((r = gemm(s, MMR_PROF_TRSM, kb + NB, kb + OB, kb, NB, cb, ce, dep))) return r;
    return trsm(s, Linv(kb, 1), kb + NB, NB, cb, ce, dep);
  };
  // One pass over the outer blocks from kb0 on.  Entry state: everything left of kb0 is factored and applied,
  // block kb0 has received all earlier updates (and, with first_inner = 1, its panel 0 is done).
  static int w_early_on = -1;
  if (w_early_on < 0) {
    const char* e = getenv("MMR_LU_W_EARLY");  // 0: W waits for the previous big GEMM (round-2 start)
    w_early_on = (e && e[0] == '0') ? 0 : 1;
  }
  static int narrow_side_on = -1;
  if (narrow_side_on < 0) {
    const char* e = getenv("MMR_LU_NARROW_SIDE");  // 0: the narrow update stays on the main stream (round-2 start)
    narrow_side_on = (e && e[0] == '0') ? 0 : 1;
  }
  auto run = [&](int kb0, int first_inner) -> int {
    int r = factor_outer(main_s, kb0, min(OB, N - kb0), first_inner);
    if (r) return r;
    for (int kb = kb0; kb < N; kb += OB) {
      const int ob = min(OB, N - kb);
      const int next = kb + ob;
      // interchanges of this outer block on the columns already factored
      if ((r = laswp(main_s, 0, kb, pslot(kb, 0), ob > NB ? pslot(kb, 1) : -1, last_panel(kb)))) return r;
      if (next >= N) break;
      const int ob2 = min(OB, N - next);
      const int cb2 = next + ob2;  // first bulk column
      const int par = (kb / OB) & 1;
      const bool use_W = defer_trsm && spec_on && ob == OB && N - cb2 >= 4 * NB;
      if (use_W) {
        // W of this block on the communication stream, beside the narrow update below
        MMR_CUDA(cudaEventRecord(e_it, main_s));
        // W needs this block's factorization (ev[1], recorded on the side stream in the previous iteration) and its
        // buffer, NOT the previous big GEMM: waiting for e_it here (everything queued on the main stream) put the three
        // W products between two big GEMMs -- 0.15-0.2 ms of a nearly idle GPU per block, 17.7 ms of a 281 ms
        // factorization (tools/lu_single_timeline.py).  First block of a pass: factored on the main stream, so e_it.
        if (lookahead && kb != kb0 && w_early_on) MMR_CUDA(cudaStreamWaitEvent(cm_s, ctx->ev[1], 0));
        else MMR_CUDA(cudaStreamWaitEvent(cm_s, e_it, 0));
        if (par_used[par]) MMR_CUDA(cudaStreamWaitEvent(cm_s, e_bulk[par], 0));  // the GEMM that read this W buffer
        if ((r = build_W(kb, par))) return r;
        MMR_CUDA(cudaEventRecord(e_W[par], cm_s));
      }
      // The narrow update of the next block's columns: with W it does not feed the big GEMM of this iteration (other
      // columns), only the factorization of the next block -- so it goes to the high-priority side stream in front of
      // that factorization and the main stream runs one big GEMM behind the other (it used to sit between them:
      // ~0.2 ms of a nearly idle GPU per block).  e_it: everything queued on the main stream so far, i.e. the previous
      // big GEMM (which updated these columns) and the wait for this block's factorization.
      const bool narrow_on_side = use_W && lookahead && narrow_side_on;
      if (narrow_on_side) {
        MMR_CUDA(cudaStreamWaitEvent(side_s, e_it, 0));
        if ((r = update_cols_on(side_s, MMR_PROF_PANEL_GEMM, kb, ob, next, cb2))) return r;
      } else {
        if ((r = update_cols(kb, ob, next, cb2))) return r;
      }
      if (lookahead) {
        MMR_CUDA(cudaEventRecord(ctx->ev[0], main_s));
        if (use_W) {
          MMR_CUDA(cudaStreamWaitEvent(main_s, e_W[par], 0));
          const int m = N - next;
          const int ps = mmr_prof_begin(ctx, MMR_PROF_GEMM, main_s);
          r = mmr_dgemm_launch(ctx, main_s, m, N - cb2, OB, -1.0, Wbuf[par], m, A + (int64_t)cb2 * ld + kb, ld, 1.0,
                               A + (int64_t)cb2 * ld + next, ld, d_fail, kb + NB);
          if (r) return r;
          mmr_prof_end(ctx, ps, 2.0 * m * (double)(N - cb2) * OB, main_s);
          MMR_CUDA(cudaEventRecord(e_bulk[par], main_s));
          MMR_CUDA(cudaStreamWaitEvent(aux_s, e_bulk[par], 0));
          if ((r = finish_U12(aux_s, kb, cb2, N))) return r;
          MMR_CUDA(cudaEventRecord(e_trsm[(kb / OB) & 3], aux_s));
          trsm_used[(kb / OB) & 3] = true;
          par_used[par] = true;
        } else {
          if ((r = update_cols(kb, ob, cb2, N))) return r;
        }
        MMR_CUDA(cudaStreamWaitEvent(side_s, ctx->ev[0], 0));
        // the L11^-1 buffers of block `next` are those of block next - 4 OB: its deferred U12 rows must be done
        if (trsm_used[(next / OB) & 3]) MMR_CUDA(cudaStreamWaitEvent(side_s, e_trsm[(next / OB) & 3], 0));
        if ((r = factor_outer(side_s, next, ob2, 0))) return r;
        MMR_CUDA(cudaEventRecord(ctx->ev[1], side_s));
        MMR_CUDA(cudaStreamWaitEvent(main_s, ctx->ev[1], 0));
      } else {
        if ((r = update_cols(kb, ob, cb2, N))) return r;
        if ((r = factor_outer(main_s, next, ob2, 0))) return r;
      }
    }
    // whatever follows on the main stream comes after the deferred work
    MMR_CUDA(cudaEventRecord(e_join, cm_s));
    MMR_CUDA(cudaStreamWaitEvent(main_s, e_join, 0));
    MMR_CUDA(cudaEventRecord(e_join, aux_s));
    MMR_CUDA(cudaStreamWaitEvent(main_s, e_join, 0));
    return 0;
  };
  int* h_info = (int*)ctx->hpin;
  int kb0 = 0, first_inner = 0;
  for (;;) {
    rc = run(kb0, first_inner);
    if (rc) return rc;
    MMR_CUDA(cudaMemcpyAsync(h_info, d_info, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    MMR_CUDA(cudaStreamSynchronize(ctx->stream));
    if (!spec_on || h_info[1] == 0) break;
    // a speculative panel was rejected at column kf: everything that depends on it skipped its stores, so the
    // matrix is exactly in the entry state of run() for that block -- resume there with the pivoting kernels
    const int kf = h_info[1] - 1;
    spec_on = false;
    kb0 = kf / OB * OB;
    first_inner = (kf > kb0) ? 1 : 0;
    MMR_CUDA(cudaMemsetAsync(d_fail, 0, sizeof(int), ctx->stream));
  }
  if (spec_on && h_info[0] == 0) {
    // no panel was rejected: switch the factors to inverse format (diagonal panel blocks <- L11^-1 \ U11^-1)
    lu_store_inverses_kernel<<<npanel_slots, 256, 0, ctx->stream>>>(A, ld, stash, npanel_slots, N, OB, 1, 0, d_fail);
    MMR_CHECK_LAUNCH(ctx);
    lu_set_format_kernel<<<1, 1, 0, ctx->stream>>>(d_ipiv, d_fail);
    MMR_CHECK_LAUNCH(ctx);
  }
  if (info_out) *info_out = h_info[0];
  return 0;
}


// U^T w = b (forward) then L^T v = w (backward), in place on d_b (length N).  Right-looking:
// as soon as a diagonal block of the solution is final it is eliminated from ALL remaining
// equations at once (one warp per remaining column, SB-long contiguous dot products).
int mmr_lu_tri_solves(mmr_ctx* ctx, int N, const double* A, int64_t ld, double* d_b) {
  const size_t smem = (size_t)SB * (SB + 1) * sizeof(double);
  int arc = mmr_func_smem(ctx, (const void*)lu_diag_solve_kernel<true>, smem);
  if (arc) return arc;
  if ((arc = mmr_func_smem(ctx, (const void*)lu_diag_solve_kernel<false>, smem))) return arc;
  for (int k0 = 0; k0 < N; k0 += SB) {
    const int k1 = min(N, k0 + SB);
    lu_diag_solve_kernel<true><<<1, DIAG_THREADS, smem, ctx->stream>>>(A, ld, k0, k1, N, d_b);
    MMR_CHECK_LAUNCH(ctx);
    if (k1 < N) {
      lu_gemv_t_kernel<<<(unsigned)ceil_div64((int64_t)(N - k1) * 32, 256), 256, 0, ctx->stream>>>(
          A, ld, k0, k1, k1, N, d_b, d_b);
      MMR_CHECK_LAUNCH(ctx);
    }
  }
  const int nblk = (int)ceil_div64(N, SB);
  for (int blk = nblk - 1; blk >= 0; --blk) {
    const int k0 = blk * SB;
    const int k1 = min(N, k0 + SB);
    lu_diag_solve_kernel<false><<<1, DIAG_THREADS, smem, ctx->stream>>>(A, ld, k0, k1, N, d_b);
    MMR_CHECK_LAUNCH(ctx);
    if (k0 > 0) {
      lu_gemv_t_kernel<<<(unsigned)ceil_div64((int64_t)k0 * 32, 256), 256, 0, ctx->stream>>>(
          A, ld, k0, k1, 0, k0, d_b, d_b);
      MMR_CHECK_LAUNCH(ctx);
    }
  }
  return 0;
}

extern "C" int mmr_lu_solve(mmr_ctx* ctx, int64_t N64, const double* d_LU, int64_t ld,
                            const int32_t* d_ipiv, double* d_B, int64_t nrhs) {
  MMR_ARG(ctx != nullptr, "ctx is NULL");
  MMR_ARG(N64 >= 0 && N64 < (1LL << 31), "N out of range");
  MMR_ARG(ld >= N64 && nrhs >= 0, "bad ld/nrhs");
  if (N64 == 0 || nrhs == 0) return 0;
  MMR_ARG(d_LU && d_ipiv && d_B, "NULL pointer");
  DeviceGuard g(ctx->device);
  const int N = (int)N64;
  // permutation from the pivot sequence, on the host: x = P v, P = S_0 S_1 ... S_{N-1}
  int rc = mmr_hpin_reserve(ctx, 2 * sizeof(int) * (size_t)N + 256);
  if (rc) return rc;
  rc = mmr_ws_reserve(ctx, sizeof(int) * (size_t)N + sizeof(double) * (size_t)N + 512);
  if (rc) return rc;
  int* h_ipiv = (int*)ctx->hpin;
  int* h_perm = h_ipiv + N;
  MMR_CUDA(cudaMemcpyAsync(h_ipiv, d_ipiv, sizeof(int) * N, cudaMemcpyDeviceToHost, ctx->stream));
  MMR_CUDA(cudaStreamSynchronize(ctx->stream));
  const bool inv_format = (h_ipiv[0] & LU_FORMAT_INVDIAG) != 0;
  h_ipiv[0] &= LU_IPIV_MASK;
  for (int i = 0; i < N; ++i) h_perm[i] = i;
  for (int i = N - 1; i >= 0; --i) {
    const int p = h_ipiv[i];
    if (p != i) {
      const int t = h_perm[i];
      h_perm[i] = h_perm[p];
      h_perm[p] = t;
    }
  }
  // after the loop x[h_perm^-1...]: we applied the swaps to an index array idx (idx = P applied
  // to identity positions), so x[i] = v[idx[i]].
  int* d_perm = (int*)ctx->ws;
  double* d_tmp = (double*)((char*)ctx->ws + ((sizeof(int) * (size_t)N + 255) / 256) * 256);
  MMR_CUDA(cudaMemcpyAsync(d_perm, h_perm, sizeof(int) * N, cudaMemcpyHostToDevice, ctx->stream));
  for (int64_t j = 0; j < nrhs; ++j) {
    double* b = d_B + j * N64;
    const int pt = mmr_prof_begin(ctx, MMR_PROF_TRISOLVE, ctx->stream);
    rc = inv_format ? lu_sweeps_inv(ctx, ctx->stream, N, lu_outer_block(), 1, 0, d_LU, ld, N, b)
                    : mmr_lu_tri_solves(ctx, N, d_LU, ld, b);
    if (rc) return rc;
    mmr_prof_end(ctx, pt, 8.0 * N * (double)N, ctx->stream);
    MMR_CUDA(cudaMemcpyAsync(d_tmp, b, sizeof(double) * N, cudaMemcpyDeviceToDevice, ctx->stream));
    lu_gather_kernel<<<(unsigned)ceil_div64(N, 256), 256, 0, ctx->stream>>>(d_tmp, d_perm, N, b);
    MMR_CHECK_LAUNCH(ctx);
  }
  return 0;
}

// =======================================================================================
// Multi-GPU: 1-D block-cyclic distributed LU + solve (one process per GPU, NCCL).
//
// Global column panel s of A = Q^T (= row block s of Q: the rows of tgt_block cells, width
// nbw = 3*tgt_block <= 128) lives on rank s % R at local column offset (s / R) * nbw.  The pivot
// search of a panel is local to its owner; the factored panel, its pivots and L11^-1 are
// broadcast; every rank then swaps and updates its own trailing columns with the same DMMA GEMM
// as the single-GPU path.  The triangular solves walk the panels in order with one nbw-double
// broadcast per panel.
// =======================================================================================
namespace {

// fold the guard word that arrived with a broadcast record into this rank's sticky guard
__global__ void lu_absorb_fail_kernel(const int* __restrict__ rec_fail, int* __restrict__ fail) {
  const int f = *rec_fail;
  if (f != 0) atomicCAS(fail, 0, f);
}

// b[c0 + c] -= sum_{r in [r0, r1)} Acol[(c0 + c)*ld + r] * x[r]; one CTA per column
__global__ void __launch_bounds__(256)
    lu_gemv_t_block_kernel(const double* __restrict__ Acol, int64_t ld, int r0, int r1, int c0,
                           const double* x, double* b) {
  __shared__ double s_part[8];
  const int c = c0 + blockIdx.x;
  const double* a = Acol + (int64_t)c * ld;
  double acc = 0.0;
  for (int r = r0 + threadIdx.x; r < r1; r += 256) acc += a[r] * x[r];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += s_part[q];
    b[c] -= t;
  }
}

}  // namespace

static int dist_check_args(mmr_ctx* ctx, int64_t N64, int64_t nrows_local, int64_t tgt_block, const double* d_Qlocal,
                           int64_t ld) {
  MMR_ARG(ctx != nullptr, "ctx is NULL");
  MMR_ARG(N64 >= 0 && N64 < (1LL << 31) && N64 % 3 == 0, "N out of range / not a multiple of 3");
  MMR_ARG(ld >= N64 && tgt_block >= 1 && nrows_local >= 0, "bad sizes");
  MMR_ARG(3 * tgt_block <= 2 * NB, "distributed LU needs 3*tgt_block <= 256");
  MMR_ARG(ctx->nranks == 1 || ctx->nccl_comm != nullptr, "communicator not initialised");
  MMR_ARG(d_Qlocal || nrows_local == 0 || N64 == 0, "NULL pointer");
  return 0;
}

extern "C" int mmr_lu_solve_dist(mmr_ctx* ctx, int64_t N64, int64_t nrows_local, int64_t tgt_block,
                                 double* d_Qlocal, int64_t ld, double* d_b, int* info_out) {
  MMR_ARG(ctx != nullptr, "ctx is NULL");
  if (info_out) *info_out = 0;
  if (N64 == 0) return 0;
  MMR_ARG(d_b != nullptr, "NULL right-hand side");
  // pivots of this call live behind the factorization's workspace
  // (pivots + the solve phase's permutation and copy of b, so that mmr_lu_trisolve_dist does not have to grow it)
  int rc = mmr_ws2_reserve(ctx, 2 * (sizeof(int) * (size_t)N64 + 512) + sizeof(double) * (size_t)N64 + 512);
  if (rc) return rc;
  int32_t* d_ipiv = (int32_t*)ctx->ws2;
  rc = mmr_lu_factor_dist(ctx, N64, nrows_local, tgt_block, d_Qlocal, ld, d_ipiv, info_out);
  if (rc) return rc;
  return mmr_lu_trisolve_dist(ctx, N64, nrows_local, tgt_block, d_Qlocal, ld, d_ipiv, d_b);
}

extern "C" int mmr_lu_factor_dist(mmr_ctx* ctx, int64_t N64, int64_t nrows_local, int64_t tgt_block,
                                  double* d_Qlocal, int64_t ld, int32_t* d_ipiv_out, int* info_out) {
  {
    const int rc0 = dist_check_args(ctx, N64, nrows_local, tgt_block, d_Qlocal, ld);
    if (rc0) return rc0;
  }
  if (info_out) *info_out = 0;
  if (N64 == 0) return 0;
  MMR_ARG(d_ipiv_out != nullptr, "NULL pivot pointer");
  DeviceGuard g(ctx->device);
  const int N = (int)N64;
  const int R = ctx->nranks, me = ctx->rank;
  const int obw = (int)(3 * tgt_block);  // outer block = ownership unit = up to two 128-column panels
  const int S = (int)ceil_div64(N, obw);
  {
    int64_t mine = 0;
    for (int s = me; s < S; s += R) mine += min(obw, N - s * obw);
    MMR_ARG(mine == nrows_local, "nrows_local does not match the block-cyclic ownership of this rank");
  }
  const int ncols_local = (int)nrows_local;
  cudaStream_t st = ctx->stream;
  cudaStream_t side = ctx->side;
  static int spec_default = -1;
  if (spec_default < 0) {
    const char* e = getenv("MMR_LU_SPEC");
    spec_default = (e && e[0] == '0') ? 0 : 1;
  }
  bool spec_on = spec_default != 0;

  // workspace: [info | strip scratch | spec scratch | 2 x broadcast record | spec panel]
  // broadcast record of an outer block: the packed panel (m x ob, ld = m) followed by a header (L11^-1 of both
  // panels, the block's pivots, the owner's speculation guard word).  In speculative mode the record IS the
  // panel buffer of the factorization (nothing is packed), panel 0's columns travel while panel 1 is still
  // being factored, and the copy into the matrix happens after the broadcast has been issued.
  const size_t scratch_bytes = (mmr_lu_panel_scratch_bytes() + 255) / 256 * 256;
  const size_t spec_bytes = (mmr_lu_spec_scratch_bytes() + 255) / 256 * 256;
  const size_t hdr_doubles = 2 * (size_t)NB * NB + NB + 2;  // 2 x L11^-1, 2*NB pivot ints, guard word (+ pad)
  const size_t rec_bytes = ((sizeof(double) * (hdr_doubles + (size_t)N * obw) + 255) / 256) * 256;
  // inverse stash of my panels (slot 2 * local block + inner: L11^-1 | U11^-1), see lu_sweep_kernel
  const int npanel_slots = 2 * (int)ceil_div64(ncols_local, obw);
  const size_t stash_bytes = spec_on ? ((sizeof(double) * (size_t)(npanel_slots + 2) * 2 * NB * NB + 255) / 256) * 256 : 0;
  int rc = mmr_ws_reserve(ctx, 256 + scratch_bytes + spec_bytes + 2 * rec_bytes + stash_bytes + 512);
  if (rc) return rc;
  rc = mmr_hpin_reserve(ctx, 2 * sizeof(int) * (size_t)N + 256);
  if (rc) return rc;
  char* base = (char*)ctx->ws;
  int* d_info = (int*)base;  // [0] zero-pivot report, [1] speculation guard (lu_poisoned)
  int* d_fail = d_info + 1;
  void* scratch = base + 256;
  void* d_spec = base + 256 + scratch_bytes;
  int* d_ipiv = d_ipiv_out;
  double* rec[2] = {(double*)(base + 256 + scratch_bytes + spec_bytes),
                    (double*)(base + 256 + scratch_bytes + spec_bytes + rec_bytes)};
  double* stash = (double*)(base + 256 + scratch_bytes + spec_bytes + 2 * rec_bytes);
  auto slot_of = [&](int s, int inner) { return stash + (size_t)(2 * (s / R) + inner) * 2 * NB * NB; };
  auto Mrows = [&](int s) { return N - s * obw; };
  auto Width = [&](int s) { return min(obw, N - s * obw); };
  auto PanelOf = [&](int s) { return rec[s & 1]; };
  auto HdrOf = [&](int s) { return rec[s & 1] + (size_t)Mrows(s) * Width(s); };
  auto LinvOf = [&](int s, int inner) { return HdrOf(s) + (size_t)inner * NB * NB; };
  auto IpivOf = [&](int s) { return (int*)(HdrOf(s) + 2 * (size_t)NB * NB); };
  auto FailOf = [&](int s) { return (int*)(HdrOf(s) + 2 * (size_t)NB * NB + NB); };
  auto pslot = [&](int s, int inner) { return 2 + (((s & 1) << 1) + inner); };
  // start of the last panel of block s: what a consumer of the whole block depends on
  auto last_panel = [&](int s) { return (Width(s) > NB) ? s * obw + NB : s * obw; };
  cudaStream_t cm = ctx->comm;
  cudaEvent_t e_upd = ctx->evd[0], e_free = ctx->evd[1], e_pA = ctx->evd[2], e_pB = ctx->evd[3], e_start = ctx->evd[4],
              e_end = ctx->evd[5];
  cudaEvent_t e_rec[2] = {ctx->evd[6], ctx->evd[7]};
  MMR_CUDA(cudaMemsetAsync(d_info, 0, 2 * sizeof(int), st));
  double* Al = d_Qlocal;

  auto trsm = [&](cudaStream_t s_, const double* Li, int nb, double* Bp, int ncols, int dep) -> int {
    const int ps = mmr_prof_begin(ctx, MMR_PROF_TRSM, s_);
    const int r = mmr_dgemm_launch(ctx, s_, nb, ncols, nb, 1.0, Li, nb, Bp, ld, 0.0, Bp, ld, d_fail, dep);
    mmr_prof_end(ctx, ps, 2.0 * nb * nb * (double)ncols, s_);
    return r;
  };
  auto copy_guarded = [&](cudaStream_t s_, const double* src, int64_t lds_, int rows, int cols, double* dst, int64_t ldd,
                          int dep) -> int {
    if (rows <= 0 || cols <= 0) return 0;
    const unsigned g = (unsigned)std::min<int64_t>((int64_t)4 * ctx->sm_count, ceil_div64((int64_t)rows * cols, 256 * 8));
    lu_copy_guarded_kernel<<<g, 256, 0, s_>>>(src, lds_, rows, cols, dst, ldd, d_fail, dep);
    MMR_CHECK_LAUNCH(ctx);
    return 0;
  };
  // split: panel 0's columns of the record are broadcast on their own as soon as they exist
  auto split_bcast = [&](int s, int first_inner) { return R > 1 && spec_on && first_inner == 0 && Width(s) > NB; };

  // Speculative factorization of my block s straight into its broadcast record (stream-ordered, no packing):
  //   diag+inverses(0) -> L21(0) [+check]  => e_pA
  //   U12 = L11^-1 A12 (into the record) -> A22 -= L21 U12 -> diag+inverses(1) -> L21(1) [+check] -> guard word  => e_pB
  //   (after e_pB, off the chain) record -> matrix
  auto factor_spec = [&](int s, cudaStream_t strm) -> int {
    const int kb = s * obw, ob = Width(s), m = Mrows(s);
    const int nb0 = min(NB, ob), nb1 = ob - nb0, k1 = kb + nb0;
    double* Ash = Al + ((int64_t)(s / R) * obw - kb) * ld;  // global column kb -> my local column
    double* P = PanelOf(s);
    lu_fill_ipiv_kernel<<<(unsigned)ceil_div64(ob, 256), 256, 0, strm>>>(IpivOf(s), d_ipiv + kb, kb, ob);
    MMR_CHECK_LAUNCH(ctx);
    // panel 0; U12 = L11^-1 A12 (out of place, into the record) shares the launch of L21 = A21 U11^-1
    mmr_gemm::GemmAlt u12;
    if (nb1 > 0) {
      u12.m = nb0, u12.n = nb1, u12.k = nb0;
      u12.A = LinvOf(s, 0), u12.lda = nb0;
      u12.B = Ash + (int64_t)k1 * ld + kb, u12.ldb = ld;
      u12.C = P + (size_t)nb0 * m, u12.ldc = m;
    }
    int r = mmr_lu_panel_spec(ctx, strm, Ash, ld, N, kb, nb0, d_ipiv, P, m, LinvOf(s, 0), slot_of(s, 0) + (size_t)NB * NB,
                              d_fail, false, nb1 > 0 ? &u12 : nullptr);
    if (r) return r;
    if (split_bcast(s, 0)) MMR_CUDA(cudaEventRecord(e_pA, strm));
    if (nb1 > 0) {
      int ps = mmr_prof_begin(ctx, MMR_PROF_PANEL_GEMM, strm);
      r = mmr_dgemm_launch(ctx, strm, N - k1, nb1, nb0, -1.0, P + nb0, m, P + (size_t)nb0 * m, m, 1.0,
                           Ash + (int64_t)k1 * ld + k1, ld, d_fail, kb);
      if (r) return r;
      mmr_prof_end(ctx, ps, 2.0 * (N - k1) * (double)nb1 * nb0, strm);
      r = mmr_lu_panel_spec(ctx, strm, Ash, ld, N, k1, nb1, d_ipiv, P + (size_t)nb0 * m + nb0, m, LinvOf(s, 1),
                            slot_of(s, 1) + (size_t)NB * NB, d_fail, false);
      if (r) return r;
    }
    MMR_CUDA(cudaMemcpyAsync(FailOf(s), d_fail, sizeof(int), cudaMemcpyDeviceToDevice, strm));
    MMR_CUDA(cudaEventRecord(e_pB, strm));
    // record -> matrix, per panel (a rejected panel 1 must leave panel 0 and its U12 block committed: the
    // host resumes from panel 1 with the pivoting kernels); L11^-1 -> inverse stash
    MMR_CUDA(cudaMemcpyAsync(slot_of(s, 0), LinvOf(s, 0), sizeof(double) * nb0 * nb0, cudaMemcpyDeviceToDevice, strm));
    if (nb1 > 0)
      MMR_CUDA(cudaMemcpyAsync(slot_of(s, 1), LinvOf(s, 1), sizeof(double) * nb1 * nb1, cudaMemcpyDeviceToDevice, strm));
    if ((r = copy_guarded(strm, P, m, m, nb0, Ash + (int64_t)kb * ld + kb, ld, kb))) return r;
    if (nb1 > 0) {
      if ((r = copy_guarded(strm, P + (size_t)nb0 * m, m, nb0, nb1, Ash + (int64_t)k1 * ld + kb, ld, kb))) return r;
      if ((r = copy_guarded(strm, P + (size_t)nb0 * m + nb0, m, m - nb0, nb1, Ash + (int64_t)k1 * ld + k1, ld, k1))) return r;
    }
    return 0;
  };
  // Pivoting factorization of my block s in place, then packed into its record
  // (first_inner = 1: resuming after a rejected panel 1 -- panel 0 is committed in A, its L11^-1 is rebuilt)
  auto factor_pivoting = [&](int s, cudaStream_t strm, int first_inner) -> int {
    const int kb = s * obw, ob = Width(s), m = Mrows(s);
    const int nb0 = min(NB, ob), nb1 = ob - nb0;
    double* Ash = Al + ((int64_t)(s / R) * obw - kb) * ld;
    int r = 0;
    if (first_inner == 0) {
      if ((r = mmr_lu_panel(ctx, strm, Ash, ld, N, kb, nb0, d_ipiv, d_info, scratch))) return r;
    }
    if ((r = mmr_lu_trtri(ctx, strm, Ash + (int64_t)kb * ld + kb, ld, nb0, LinvOf(s, 0)))) return r;
    if (nb1 > 0) {
      const int k1 = kb + nb0;
      if (first_inner == 0) {
        if ((r = mmr_lu_laswp(ctx, strm, Ash, ld, kb, k1, k1, k1 + nb1, d_ipiv, d_fail))) return r;
        if ((r = trsm(strm, LinvOf(s, 0), nb0, Ash + (int64_t)k1 * ld + kb, nb1, kb))) return r;
        const int pg = mmr_prof_begin(ctx, MMR_PROF_PANEL_GEMM, strm);
        r = mmr_dgemm_launch(ctx, strm, N - k1, nb1, nb0, -1.0, Ash + (int64_t)kb * ld + k1, ld,
                             Ash + (int64_t)k1 * ld + kb, ld, 1.0, Ash + (int64_t)k1 * ld + k1, ld, d_fail, kb);
        if (r) return r;
        mmr_prof_end(ctx, pg, 2.0 * (N - k1) * (double)nb1 * nb0, strm);
      }
      if ((r = mmr_lu_panel(ctx, strm, Ash, ld, N, k1, nb1, d_ipiv, d_info, scratch))) return r;
      if ((r = mmr_lu_trtri(ctx, strm, Ash + (int64_t)k1 * ld + k1, ld, nb1, LinvOf(s, 1)))) return r;
      if ((r = mmr_lu_laswp(ctx, strm, Ash, ld, k1, k1 + nb1, kb, k1, d_ipiv, d_fail))) return r;
    }
    MMR_CUDA(cudaMemcpy2DAsync(PanelOf(s), sizeof(double) * m, Ash + (int64_t)kb * ld + kb, sizeof(double) * ld,
                               sizeof(double) * m, ob, cudaMemcpyDeviceToDevice, strm));
    MMR_CUDA(cudaMemcpyAsync(IpivOf(s), d_ipiv + kb, sizeof(int) * ob, cudaMemcpyDeviceToDevice, strm));
    MMR_CUDA(cudaMemcpyAsync(FailOf(s), d_fail, sizeof(int), cudaMemcpyDeviceToDevice, strm));
    MMR_CUDA(cudaEventRecord(e_pB, strm));
    return 0;
  };
  auto factor_block = [&](int s, cudaStream_t strm, int first_inner) -> int {
    return (spec_on && first_inner == 0) ? factor_spec(s, strm) : factor_pivoting(s, strm, first_inner);
  };
  // apply outer block s (interchanges, U12 in two 128-row steps, one rank-ob update) to my local
  // columns [c0, c1)
  auto update_local = [&](int s, int c0, int c1) -> int {
    if (c1 <= c0) return 0;
    const int kb = s * obw, ob = Width(s), m = Mrows(s);
    const int nb0 = min(NB, ob), nb1 = ob - nb0;
    const int dep = last_panel(s);
    const double* P = PanelOf(s);
    double* B = Al + (int64_t)c0 * ld + kb;
    int ps, r = 0;
    if (!spec_on) {
      ps = mmr_prof_begin(ctx, MMR_PROF_LASWP, st);
      r = mmr_lu_laswp_apply(ctx, st, Al, ld, c0, c1, pslot(s, 0), nb1 > 0 ? pslot(s, 1) : -1, d_fail, dep);
      if (r) return r;
      mmr_prof_end(ctx, ps, 32.0 * ob * (double)(c1 - c0), st);
    }
    bool fused = false;
    if (nb1 > 0 && c1 - c0 <= BLOCK_TRSM_MAX_COLS) {  // (the look-ahead update of my next block: one launch)
      r = mmr_block_trsm_launch(ctx, st, nb0, nb1, LinvOf(s, 0), LinvOf(s, 1), P + nb0, m, B, ld, c1 - c0, d_fail, dep);
      if (r == 0) fused = true;
      else if (r != -2) return r;
    }
    if (!fused) {
      if ((r = trsm(st, LinvOf(s, 0), nb0, B, c1 - c0, dep))) return r;
      if (nb1 > 0) {
        ps = mmr_prof_begin(ctx, MMR_PROF_TRSM, st);
        r = mmr_dgemm_launch(ctx, st, nb1, c1 - c0, nb0, -1.0, P + nb0, m, B, ld, 1.0, B + nb0, ld, d_fail, dep);
        if (r) return r;
        mmr_prof_end(ctx, ps, 2.0 * nb1 * (double)(c1 - c0) * nb0, st);
        if ((r = trsm(st, LinvOf(s, 1), nb1, B + nb0, c1 - c0, dep))) return r;
      }
    }
    if (m > ob) {
      ps = mmr_prof_begin(ctx, MMR_PROF_GEMM, st);
      r = mmr_dgemm_launch(ctx, st, m - ob, c1 - c0, ob, -1.0, P + ob, m, B, ld, 1.0, B + ob, ld, d_fail, dep);
      if (r) return r;
      mmr_prof_end(ctx, ps, 2.0 * (m - ob) * (double)(c1 - c0) * ob, st);
    }
    return 0;
  };
  // broadcast record s on the communication stream (beside the panel chain and the trailing update), then
  // fold the owner's guard word into mine; e_rec[s & 1] marks the arrival.  `split`: panel 0's columns first.
  auto bcast_record = [&](int s, bool split) -> int {
    const int ob = Width(s), m = Mrows(s), owner = s % R;
    const int nb0 = min(NB, ob);
    if (R > 1) {
      const size_t all = sizeof(double) * ((size_t)m * ob + hdr_doubles);
      const size_t part0 = split ? sizeof(double) * (size_t)m * nb0 : 0;
      if (split) {
        if (me == owner) MMR_CUDA(cudaStreamWaitEvent(cm, e_pA, 0));
        const int pc = mmr_prof_begin(ctx, MMR_PROF_COMM, cm);
        const int r = mmr_comm_bcast_bytes(ctx, PanelOf(s), part0, owner, cm);
        if (r) return r;
        mmr_prof_end(ctx, pc, (double)part0, cm);
      }
      if (me == owner) MMR_CUDA(cudaStreamWaitEvent(cm, e_pB, 0));
      const int pc = mmr_prof_begin(ctx, MMR_PROF_COMM, cm);
      const int r = mmr_comm_bcast_bytes(ctx, (char*)PanelOf(s) + part0, all - part0, owner, cm);
      if (r) return r;
      mmr_prof_end(ctx, pc, (double)(all - part0), cm);
    } else {
      MMR_CUDA(cudaStreamWaitEvent(cm, e_pB, 0));
    }
    lu_absorb_fail_kernel<<<1, 1, 0, cm>>>(FailOf(s), d_fail);
    MMR_CHECK_LAUNCH(ctx);
    MMR_CUDA(cudaEventRecord(e_rec[s & 1], cm));
    return 0;
  };

  // One pass over the blocks from s0 on.  Entry state: every block before s0 is factored and applied
  // everywhere, block s0 has received all earlier updates (first_inner = 1: and its panel 0 is done).
  auto run = [&](int s0, int first_inner) -> int {
    MMR_CUDA(cudaEventRecord(e_start, st));
    MMR_CUDA(cudaStreamWaitEvent(side, e_start, 0));  // side and comm streams start after everything queued so far
    MMR_CUDA(cudaStreamWaitEvent(cm, e_start, 0));
    int r = 0;
    if (me == s0 % R && (r = factor_block(s0, side, first_inner))) return r;
    if ((r = bcast_record(s0, split_bcast(s0, first_inner)))) return r;
    for (int s = s0; s < S; ++s) {
      const int kb = s * obw;
      const int ob = Width(s);
      const int nb0 = min(NB, ob), nb1 = ob - nb0;
      MMR_CUDA(cudaStreamWaitEvent(st, e_rec[s & 1], 0));
      // from here on the main stream no longer reads record s-1 = the buffer of record s+1
      MMR_CUDA(cudaEventRecord(e_free, st));
      if (me != s % R)
        MMR_CUDA(cudaMemcpyAsync(d_ipiv + kb, IpivOf(s), sizeof(int) * ob, cudaMemcpyDeviceToDevice, st));
      if (!spec_on) {
        if ((r = mmr_lu_laswp_prepare(ctx, st, d_ipiv, kb, kb + nb0, pslot(s, 0), d_fail))) return r;
        if (nb1 > 0 && (r = mmr_lu_laswp_prepare(ctx, st, d_ipiv, kb + nb0, kb + ob, pslot(s, 1), d_fail))) return r;
        // my already-factored blocks (global index < s): interchanges only
        const int nleft = (s > me) ? ((s - 1 - me) / R + 1) : 0;
        const int ps = mmr_prof_begin(ctx, MMR_PROF_LASWP, st);
        r = mmr_lu_laswp_apply(ctx, st, Al, ld, 0, nleft * obw, pslot(s, 0), nb1 > 0 ? pslot(s, 1) : -1, d_fail,
                               last_panel(s));
        if (r) return r;
        mmr_prof_end(ctx, ps, 32.0 * ob * (double)nleft * obw, st);
      }
      // my trailing blocks (global index > s)
      const int first_trail = (s >= me) ? ((s - me) / R + 1) : 0;
      const int ct0 = first_trail * obw, ct1 = ncols_local;
      const bool have_next = s + 1 < S;
      const bool own_next = have_next && ((s + 1) % R == me);
      if (own_next) {
        // look-ahead: bring my next block up to date first, then factor it on the side stream while
        // the main stream updates the rest of my trailing columns
        const int ob1 = Width(s + 1);
        if ((r = update_local(s, ct0, ct0 + ob1))) return r;
        MMR_CUDA(cudaEventRecord(e_upd, st));
        if ((r = update_local(s, ct0 + ob1, ct1))) return r;
        MMR_CUDA(cudaStreamWaitEvent(side, e_upd, 0));  // (implies e_free: the record buffer of s+1 is free)
        if ((r = factor_block(s + 1, side, 0))) return r;
      } else {
        if ((r = update_local(s, ct0, ct1))) return r;
      }
      if (have_next) {
        MMR_CUDA(cudaStreamWaitEvent(cm, e_free, 0));
        if ((r = bcast_record(s + 1, split_bcast(s + 1, 0)))) return r;
      }
    }
    // whatever follows on the main stream comes after everything on the other two streams
    MMR_CUDA(cudaEventRecord(e_end, side));
    MMR_CUDA(cudaStreamWaitEvent(st, e_end, 0));
    MMR_CUDA(cudaEventRecord(e_end, cm));
    MMR_CUDA(cudaStreamWaitEvent(st, e_end, 0));
    return 0;
  };

  int* h_info = (int*)ctx->hpin;
  int s0 = 0, first_inner = 0;
  for (;;) {
    rc = run(s0, first_inner);
    if (rc) return rc;
    MMR_CUDA(cudaMemcpyAsync(h_info, d_info, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    MMR_CUDA(cudaStreamSynchronize(st));
    if (!spec_on || h_info[1] == 0) break;
    // a speculative panel was rejected at column kf (every rank holds the same guard word: it travels in
    // the records): resume at its block with the pivoting kernels
    const int kf = h_info[1] - 1;
    spec_on = false;
    s0 = kf / obw;
    first_inner = (kf > s0 * obw) ? 1 : 0;
    MMR_CUDA(cudaMemsetAsync(d_fail, 0, sizeof(int), st));
  }
  if (spec_on) {
    // no panel was rejected anywhere (the guard word is the same on every rank): inverse format
    if (npanel_slots > 0) {
      lu_store_inverses_kernel<<<npanel_slots, 256, 0, st>>>(Al, ld, stash, npanel_slots, N, obw, R, me, d_fail);
      MMR_CHECK_LAUNCH(ctx);
    }
    lu_set_format_kernel<<<1, 1, 0, st>>>(d_ipiv, d_fail);
    MMR_CHECK_LAUNCH(ctx);
    MMR_CUDA(cudaStreamSynchronize(st));
  }
  if (info_out) *info_out = h_info[0];  // only the owner of a zero pivot sees it; callers reduce over ranks
  return 0;
}

// Triangular solves with the distributed factors of mmr_lu_factor_dist (d_ipiv: the N replicated pivots it
// returned): U^T w = b, L^T v = w, x = P v.  d_b (N, replicated) is overwritten by x on every rank.
extern "C" int mmr_lu_trisolve_dist(mmr_ctx* ctx, int64_t N64, int64_t nrows_local, int64_t tgt_block,
                                    const double* d_Qlocal, int64_t ld, const int32_t* d_ipiv, double* d_b) {
  {
    const int rc0 = dist_check_args(ctx, N64, nrows_local, tgt_block, d_Qlocal, ld);
    if (rc0) return rc0;
  }
  if (N64 == 0) return 0;
  MMR_ARG(d_b && d_ipiv, "NULL pointer");
  DeviceGuard g(ctx->device);
  const int N = (int)N64;
  const int R = ctx->nranks, me = ctx->rank;
  const int obw = (int)(3 * tgt_block);
  const int S = (int)ceil_div64(N, obw);
  cudaStream_t st = ctx->stream;
  const double* Al = d_Qlocal;
  int rc = mmr_hpin_reserve(ctx, 2 * sizeof(int) * (size_t)N + 256);
  if (rc) return rc;
  // scratch of the solve phase lives in ws2 BEHIND the pivots mmr_lu_solve_dist keeps there
  const size_t off = ((sizeof(int) * (size_t)N + 256 + 255) / 256) * 256;
  const size_t perm_bytes = ((sizeof(int) * (size_t)N + 255) / 256) * 256;
  if (ctx->ws2_bytes < off + perm_bytes + sizeof(double) * (size_t)N) {
    // growing ws2 would free the pivots when they live there: stage them through the pinned buffer
    const bool inside = (const char*)d_ipiv >= (const char*)ctx->ws2 && (const char*)d_ipiv < (const char*)ctx->ws2 + ctx->ws2_bytes;
    if (inside) {
      MMR_CUDA(cudaMemcpyAsync(ctx->hpin, d_ipiv, sizeof(int) * N, cudaMemcpyDeviceToHost, st));
      MMR_CUDA(cudaStreamSynchronize(st));
    }
    rc = mmr_ws2_reserve(ctx, off + perm_bytes + sizeof(double) * (size_t)N);
    if (rc) return rc;
    if (inside) {
      MMR_CUDA(cudaMemcpyAsync(ctx->ws2, ctx->hpin, sizeof(int) * N, cudaMemcpyHostToDevice, st));
      d_ipiv = (const int32_t*)ctx->ws2;
    }
  }
  int* d_perm = (int*)((char*)ctx->ws2 + off);
  double* d_tmp = (double*)((char*)ctx->ws2 + off + perm_bytes);
  int* h_ipiv = (int*)ctx->hpin;
  int* h_perm = h_ipiv + N;
  MMR_CUDA(cudaMemcpyAsync(h_ipiv, d_ipiv, sizeof(int) * N, cudaMemcpyDeviceToHost, st));
  MMR_CUDA(cudaStreamSynchronize(st));
  const bool inv_format = (h_ipiv[0] & LU_FORMAT_INVDIAG) != 0;
  h_ipiv[0] &= LU_IPIV_MASK;
  const int pt = mmr_prof_begin(ctx, MMR_PROF_TRISOLVE, st);
  if (inv_format) {
    int64_t mine = 0;
    for (int s = me; s < S; s += R) mine += min(obw, N - s * obw);
    rc = lu_sweeps_inv(ctx, st, N, obw, R, me, Al, ld, (int)mine, d_b);
    if (rc) return rc;
  } else {

  // ---- triangular solves, block by block (each block in two <= 128-row steps) -----------------
  const size_t diag_smem = (size_t)SB * (SB + 1) * sizeof(double);
  if ((rc = mmr_func_smem(ctx, (const void*)lu_diag_solve_kernel<true>, diag_smem))) return rc;
  if ((rc = mmr_func_smem(ctx, (const void*)lu_diag_solve_kernel<false>, diag_smem))) return rc;
  for (int s = 0; s < S; ++s) {  // U^T w = b
    const int kb = s * obw, ob = min(obw, N - kb), owner = s % R;
    const int nb0 = min(NB, ob), nb1 = ob - nb0;
    if (me == owner) {
      const double* Ash = Al + ((int64_t)(s / R) * obw - kb) * ld;
      if (kb > 0) {
        lu_gemv_t_block_kernel<<<ob, 256, 0, st>>>(Ash, ld, 0, kb, kb, d_b, d_b);
        MMR_CHECK_LAUNCH(ctx);
      }
      lu_diag_solve_kernel<true><<<1, DIAG_THREADS, diag_smem, st>>>(Ash, ld, kb, kb + nb0, 0, d_b);
      MMR_CHECK_LAUNCH(ctx);
      if (nb1 > 0) {
        lu_gemv_t_block_kernel<<<nb1, 256, 0, st>>>(Ash, ld, kb, kb + nb0, kb + nb0, d_b, d_b);
        MMR_CHECK_LAUNCH(ctx);
        lu_diag_solve_kernel<true><<<1, DIAG_THREADS, diag_smem, st>>>(Ash, ld, kb + nb0, kb + ob, 0, d_b);
        MMR_CHECK_LAUNCH(ctx);
      }
    }
    if (R > 1) {
      rc = mmr_comm_bcast_bytes(ctx, d_b + kb, sizeof(double) * ob, owner, st);
      if (rc) return rc;
    }
  }
  for (int s = S - 1; s >= 0; --s) {  // L^T v = w
    const int kb = s * obw, ob = min(obw, N - kb), owner = s % R;
    const int nb0 = min(NB, ob), nb1 = ob - nb0;
    if (me == owner) {
      const double* Ash = Al + ((int64_t)(s / R) * obw - kb) * ld;
      if (kb + ob < N) {
        lu_gemv_t_block_kernel<<<ob, 256, 0, st>>>(Ash, ld, kb + ob, N, kb, d_b, d_b);
        MMR_CHECK_LAUNCH(ctx);
      }
      if (nb1 > 0) {
        lu_diag_solve_kernel<false><<<1, DIAG_THREADS, diag_smem, st>>>(Ash, ld, kb + nb0, kb + ob, 0, d_b);
        MMR_CHECK_LAUNCH(ctx);
        lu_gemv_t_block_kernel<<<nb0, 256, 0, st>>>(Ash, ld, kb + nb0, kb + ob, kb, d_b, d_b);
        MMR_CHECK_LAUNCH(ctx);
      }
      lu_diag_solve_kernel<false><<<1, DIAG_THREADS, diag_smem, st>>>(Ash, ld, kb, kb + nb0, 0, d_b);
      MMR_CHECK_LAUNCH(ctx);
    }
    if (R > 1) {
      rc = mmr_comm_bcast_bytes(ctx, d_b + kb, sizeof(double) * ob, owner, st);
      if (rc) return rc;
    }
  }
  }  // plain format
  mmr_prof_end(ctx, pt, 8.0 * N * (double)nrows_local, st);
  // ---- x = P v (pivots are replicated) ------------------------------------------------------
  for (int i = 0; i < N; ++i) h_perm[i] = i;
  for (int i = N - 1; i >= 0; --i) {
    const int p = h_ipiv[i];
    if (p != i) {
      const int t = h_perm[i];
      h_perm[i] = h_perm[p];
      h_perm[p] = t;
    }
  }
  MMR_CUDA(cudaMemcpyAsync(d_perm, h_perm, sizeof(int) * N, cudaMemcpyHostToDevice, st));
  MMR_CUDA(cudaMemcpyAsync(d_tmp, d_b, sizeof(double) * N, cudaMemcpyDeviceToDevice, st));
  lu_gather_kernel<<<(unsigned)ceil_div64(N, 256), 256, 0, st>>>(d_tmp, d_perm, N, d_b);
  MMR_CHECK_LAUNCH(ctx);
  MMR_CUDA(cudaStreamSynchronize(st));
  return 0;
}
