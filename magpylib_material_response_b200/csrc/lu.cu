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

#include "dgemm.cuh"

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

// Row swaps ipiv[k0..k1) applied to columns [col_begin, col_end): one thread per column.
__global__ void lu_laswp_kernel(double* __restrict__ A, int64_t ld, int k0, int k1, int col_begin,
                                int col_end, const int* __restrict__ ipiv) {
  extern __shared__ int s_piv[];
  for (int i = threadIdx.x; i < k1 - k0; i += blockDim.x) s_piv[i] = ipiv[k0 + i];
  __syncthreads();
  const int col = col_begin + blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= col_end) return;
  double* a = A + (int64_t)col * ld;
  for (int i = 0; i < k1 - k0; ++i) {
    const int p = s_piv[i];
    if (p != k0 + i) {
      const double tmp = a[k0 + i];
      a[k0 + i] = a[p];
      a[p] = tmp;
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
// X = L^-1 for the unit-lower nb x nb diagonal block of a factored panel (column-major L, ldl).
// One CTA, one thread per column of X (forward substitution); X is kept in shared memory and
// written out as a dense nb x nb column-major matrix (upper part zero).  Runs on the look-ahead
// stream, so the main stream's "TRSM" becomes a plain DMMA GEMM U12 = X * A12.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NB)
    lu_trtri_kernel(const double* __restrict__ L, int64_t ldl, int nb, double* __restrict__ X) {
  extern __shared__ __align__(16) double Xs[];  // Xs[c*(nb+1) + r]
  const int c = threadIdx.x;
  if (c < nb) {
    double* x = Xs + c * (nb + 1);
    for (int r = 0; r < c; ++r) x[r] = 0.0;
    x[c] = 1.0;
    for (int r = c + 1; r < nb; ++r) {
      double acc = L[(int64_t)c * ldl + r];  // t = c term (x[c] = 1)
      for (int t = c + 1; t < r; ++t) acc += L[(int64_t)t * ldl + r] * x[t];
      x[r] = -acc;
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < nb * nb; e += blockDim.x) {
    const int r = e % nb, cc = e / nb;
    X[(int64_t)cc * nb + r] = Xs[cc * (nb + 1) + r];
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

// Solve the SB x SB diagonal block in place (single CTA of SB threads, block transposed into
// dynamic shared memory).
//   UPPER_T: w[c] = (b[c] - sum_{k0<=r<c} A[r,c] w[r]) / A[c,c]      ascending c
//   else   : v[c] =  b[c] - sum_{c<r<k1} A[r,c] v[r]                 descending c (unit diag)
template <bool UPPER_T>
__global__ void __launch_bounds__(SB)
    lu_diag_solve_kernel(const double* __restrict__ A, int64_t ld, int k0, int k1,
                         double* __restrict__ b) {
  extern __shared__ __align__(16) double T[];  // T[r*(SB+1) + c] = A[k0+r, k0+c]
  __shared__ double x[SB];
  const int tid = threadIdx.x;
  const int nbk = k1 - k0;
  for (int c = 0; c < nbk; ++c) {
    if (tid < nbk) T[tid * (SB + 1) + c] = A[(int64_t)(k0 + c) * ld + k0 + tid];
  }
  if (tid < nbk) x[tid] = b[k0 + tid];
  __syncthreads();
  if (UPPER_T) {
    for (int r = 0; r < nbk; ++r) {
      if (tid == r) x[r] = x[r] / T[r * (SB + 1) + r];
      __syncthreads();
      if (tid > r && tid < nbk) x[tid] -= T[r * (SB + 1) + tid] * x[r];
      __syncthreads();
    }
  } else {
    for (int r = nbk - 1; r >= 0; --r) {
      if (tid < r) x[tid] -= T[r * (SB + 1) + tid] * x[r];
      __syncthreads();
    }
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
                     double* C, int64_t ldc) {
  using namespace mmr_gemm;
  if (m <= 0 || n <= 0) return 0;
  const bool aligned = (((uintptr_t)A | (uintptr_t)B) % 16 == 0) && (lda % 2 == 0) && (ldb % 2 == 0);
  const bool subtract = (alpha == -1.0 && beta == 1.0);
  typedef void (*kern_t)(int, int, int, double, const double*, int64_t, const double*, int64_t, double,
                         double*, int64_t);
  static const kern_t kerns[4] = {dgemm_nn_kernel<false, false>, dgemm_nn_kernel<false, true>,
                                  dgemm_nn_kernel<true, false>, dgemm_nn_kernel<true, true>};
  static bool attr_set[4] = {false, false, false, false};
  const int which = (aligned ? 2 : 0) + (subtract ? 1 : 0);
  if (!attr_set[which]) {
    MMR_CUDA(cudaFuncSetAttribute(kerns[which], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    attr_set[which] = true;
  }
  dim3 grid((unsigned)ceil_div64(m, BM), (unsigned)ceil_div64(n, BN));
  const int64_t ntiles = (int64_t)grid.x * grid.y;
  static int persistent = -1;
  if (persistent < 0) {
    const char* e = getenv("MMR_DGEMM_PERSISTENT");
    persistent = (e && e[0] == '1') ? 1 : 0;  // experiment, off by default (slower than the fast kernel)
    MMR_CUDA(cudaFuncSetAttribute(dgemm_lu_persistent_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    MMR_CUDA(cudaFuncSetAttribute(dgemm_lu_persistent_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  }
  static int fast = -1;
  if (fast < 0) {
    const char* e = getenv("MMR_DGEMM_FAST");
    fast = (e && e[0] == '0') ? 0 : 1;
    MMR_CUDA(cudaFuncSetAttribute(dgemm_nn_fast_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    MMR_CUDA(cudaFuncSetAttribute(dgemm_nn_fast_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  }
  static int stagger_pct = -1;
  if (stagger_pct < 0) {
    const char* e = getenv("MMR_DGEMM_STAGGER_PCT");
    stagger_pct = e ? atoi(e) : 0;  // measured: no effect on B200, kept as an experiment knob
  }
  if (fast && aligned && k > 0 && k % BK == 0) {
    // tile time when two CTAs share the tensor pipe: 2*BM*BN*k FMA at 64 FMA/clk/SM, ~1.9 GHz
    int stagger_ns = 0;
    if (ntiles > 2 * (int64_t)ctx->sm_count)
      stagger_ns = (int)((2.0 * BM * BN * (double)k / 64.0 / 1.9) * stagger_pct / 100.0);
    if (subtract)
      dgemm_nn_fast_kernel<true><<<grid, THREADS, SMEM_BYTES, s>>>((int)m, (int)n, (int)k, alpha, A, lda, B, ldb, beta, C, ldc, stagger_ns, ctx->sm_count);
    else
      dgemm_nn_fast_kernel<false><<<grid, THREADS, SMEM_BYTES, s>>>((int)m, (int)n, (int)k, alpha, A, lda, B, ldb, beta, C, ldc, stagger_ns, ctx->sm_count);
    MMR_CHECK_LAUNCH(ctx);
    return 0;
  }
  if (subtract && persistent && ntiles > 2 * (int64_t)ctx->sm_count) {
    const unsigned pg = (unsigned)(2 * ctx->sm_count);
    if (aligned)
      dgemm_lu_persistent_kernel<true><<<pg, THREADS, SMEM_BYTES, s>>>((int)m, (int)n, (int)k, A, lda, B, ldb, C, ldc);
    else
      dgemm_lu_persistent_kernel<false><<<pg, THREADS, SMEM_BYTES, s>>>((int)m, (int)n, (int)k, A, lda, B, ldb, C, ldc);
    MMR_CHECK_LAUNCH(ctx);
    return 0;
  }
  kerns[which]<<<grid, THREADS, SMEM_BYTES, s>>>((int)m, (int)n, (int)k, alpha, A, lda, B, ldb, beta, C, ldc);
  MMR_CHECK_LAUNCH(ctx);
  return 0;
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
  static int max_coop_ctas = -1;
  static size_t strip_smem_cap = 0;
  if (max_coop_ctas < 0) {
    strip_smem_cap = 200 * 1024;
    MMR_CUDA(cudaFuncSetAttribute(lu_strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)strip_smem_cap));
    max_coop_ctas = ctx->sm_count;  // 1 CTA per SM at up to 200 KB of shared memory
  }
  static bool reg_ok = false;
  static int cluster_max = -1;  // largest launchable cluster size for the cluster strip kernel
  if (cluster_max < 0) {
    cluster_max = 0;
    if (cudaFuncSetAttribute(lu_strip_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)strip_smem_cap) == cudaSuccess &&
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
          cluster_max = cs;
          break;
        }
      }
    }
    cudaGetLastError();  // clear any probe error
    reg_ok = cudaFuncSetAttribute(lu_strip_reg_kernel<1>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
             cudaFuncSetAttribute(lu_strip_reg_kernel<2>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess &&
             cudaFuncSetAttribute(lu_strip_reg_kernel<3>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
    cudaGetLastError();
  }
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

int mmr_lu_laswp(mmr_ctx* ctx, cudaStream_t strm, double* A, int64_t ld, int k0, int k1, int col_begin, int col_end,
                 const int* d_ipiv) {
  if (col_end <= col_begin || k1 <= k0) return 0;
  const int threads = 128;
  lu_laswp_kernel<<<(unsigned)ceil_div64(col_end - col_begin, threads), threads, (k1 - k0) * sizeof(int),
                    strm>>>(A, ld, k0, k1, col_begin, col_end, d_ipiv);
  MMR_CHECK_LAUNCH(ctx);
  return 0;
}

int mmr_lu_trsm(mmr_ctx* ctx, cudaStream_t strm, const double* L, int64_t ldl, int nb, double* B, int64_t ldb, int ncols) {
  if (ncols <= 0 || nb <= 0) return 0;
  const size_t smem = ((size_t)TRSM_COLS * (nb + 1) + (size_t)nb * (nb + 1)) * sizeof(double);
  static size_t cap = 0;
  if (smem > cap) {
    MMR_CUDA(cudaFuncSetAttribute(lu_trsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cap = smem;
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
  const size_t scratch_bytes = (mmr_lu_panel_scratch_bytes() + 255) / 256 * 256;
  int rc = mmr_ws_reserve(ctx, 256 + scratch_bytes + 2 * sizeof(double) * NB * NB);
  if (rc) return rc;
  rc = mmr_hpin_reserve(ctx, 256);
  if (rc) return rc;
  int* d_info = (int*)ctx->ws;
  void* scratch = (char*)ctx->ws + 256;
  double* Linv[2] = {(double*)((char*)ctx->ws + 256 + scratch_bytes),
                     (double*)((char*)ctx->ws + 256 + scratch_bytes) + NB * NB};
  MMR_CUDA(cudaMemsetAsync(d_info, 0, sizeof(int), ctx->stream));
  double* A = d_Q;
  cudaStream_t main_s = ctx->stream;
  cudaStream_t side_s = ctx->side;
  // trailing update of the column range [cb, ce) by panel kb (swaps, U12 = L11^-1 A12, A22 -= L21 U12)
  auto update_cols = [&](int kb, int nb, int cb, int ce) -> int {
    if (ce <= cb) return 0;
    int ps = mmr_prof_begin(ctx, MMR_PROF_LASWP, main_s);
    int r = mmr_lu_laswp(ctx, main_s, A, ld, kb, kb + nb, cb, ce, d_ipiv);
    if (r) return r;
    mmr_prof_end(ctx, ps, 32.0 * nb * (double)(ce - cb), main_s);
    // U12 = L11^-1 A12 as an in-place DMMA GEMM (one m-tile: each CTA reads its whole B tile
    // into shared memory before it stores, and no other CTA touches those columns)
    ps = mmr_prof_begin(ctx, MMR_PROF_TRSM, main_s);
    r = mmr_dgemm_launch(ctx, main_s, nb, ce - cb, nb, 1.0, Linv[(kb / NB) & 1], nb, A + (int64_t)cb * ld + kb, ld,
                         0.0, A + (int64_t)cb * ld + kb, ld);
    if (r) return r;
    mmr_prof_end(ctx, ps, 2.0 * nb * nb * (ce - cb), main_s);
    const int mrows = N - (kb + nb);
    ps = mmr_prof_begin(ctx, MMR_PROF_GEMM, main_s);
    r = mmr_dgemm_launch(ctx, main_s, mrows, ce - cb, nb, -1.0, A + (int64_t)kb * ld + (kb + nb), ld,
                         A + (int64_t)cb * ld + kb, ld, 1.0, A + (int64_t)cb * ld + (kb + nb), ld);
    if (r) return r;
    mmr_prof_end(ctx, ps, 2.0 * mrows * (double)(ce - cb) * nb, main_s);
    return 0;
  };
  // Look-ahead (depth 1): as soon as the columns of panel k+1 have received the update of panel
  // k, panel k+1 is factored on the high-priority side stream while the main stream applies the
  // rest of update k.  The two touch disjoint column ranges.
  const bool lookahead = ctx->lookahead;
  rc = mmr_lu_panel(ctx, main_s, A, ld, N, 0, min(NB, N), d_ipiv, d_info, scratch);
  if (rc) return rc;
  rc = mmr_lu_trtri(ctx, main_s, A, ld, min(NB, N), Linv[0]);
  if (rc) return rc;
  for (int kb = 0; kb < N; kb += NB) {
    const int nb = min(NB, N - kb);
    const int next = kb + nb;
    // left swaps of panel kb (columns already factored)
    {
      const int ps = mmr_prof_begin(ctx, MMR_PROF_LASWP, main_s);
      rc = mmr_lu_laswp(ctx, main_s, A, ld, kb, kb + nb, 0, kb, d_ipiv);
      if (rc) return rc;
      mmr_prof_end(ctx, ps, 32.0 * nb * (double)kb, main_s);
    }
    if (next >= N) break;
    const int nb2 = min(NB, N - next);
    rc = update_cols(kb, nb, next, next + nb2);
    if (rc) return rc;
    if (lookahead) {
      MMR_CUDA(cudaEventRecord(ctx->ev[0], main_s));
      MMR_CUDA(cudaStreamWaitEvent(side_s, ctx->ev[0], 0));
      rc = mmr_lu_panel(ctx, side_s, A, ld, N, next, nb2, d_ipiv, d_info, scratch);
      if (rc) return rc;
      rc = mmr_lu_trtri(ctx, side_s, A + (int64_t)next * ld + next, ld, nb2, Linv[(next / NB) & 1]);
      if (rc) return rc;
      MMR_CUDA(cudaEventRecord(ctx->ev[1], side_s));
      rc = update_cols(kb, nb, next + nb2, N);
      if (rc) return rc;
      MMR_CUDA(cudaStreamWaitEvent(main_s, ctx->ev[1], 0));
    } else {
      rc = update_cols(kb, nb, next + nb2, N);
      if (rc) return rc;
      rc = mmr_lu_panel(ctx, main_s, A, ld, N, next, nb2, d_ipiv, d_info, scratch);
      if (rc) return rc;
      rc = mmr_lu_trtri(ctx, main_s, A + (int64_t)next * ld + next, ld, nb2, Linv[(next / NB) & 1]);
      if (rc) return rc;
    }
  }
  MMR_CUDA(cudaMemcpyAsync(ctx->hpin, d_info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  MMR_CUDA(cudaStreamSynchronize(ctx->stream));
  if (info_out) *info_out = *(int*)ctx->hpin;
  return 0;
}

// U^T w = b (forward) then L^T v = w (backward), in place on d_b (length N).  Right-looking:
// as soon as a diagonal block of the solution is final it is eliminated from ALL remaining
// equations at once (one warp per remaining column, SB-long contiguous dot products).
int mmr_lu_tri_solves(mmr_ctx* ctx, int N, const double* A, int64_t ld, double* d_b) {
  static bool attr = false;
  const size_t smem = (size_t)SB * (SB + 1) * sizeof(double);
  if (!attr) {
    MMR_CUDA(cudaFuncSetAttribute(lu_diag_solve_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MMR_CUDA(cudaFuncSetAttribute(lu_diag_solve_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  for (int k0 = 0; k0 < N; k0 += SB) {
    const int k1 = min(N, k0 + SB);
    lu_diag_solve_kernel<true><<<1, SB, smem, ctx->stream>>>(A, ld, k0, k1, d_b);
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
    lu_diag_solve_kernel<false><<<1, SB, smem, ctx->stream>>>(A, ld, k0, k1, d_b);
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
    rc = mmr_lu_tri_solves(ctx, N, d_LU, ld, b);
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
#include "comm.h"

namespace {

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

extern "C" int mmr_lu_solve_dist(mmr_ctx* ctx, int64_t N64, int64_t nrows_local, int64_t tgt_block,
                                 double* d_Qlocal, int64_t ld, double* d_b, int* info_out) {
  MMR_ARG(ctx != nullptr, "ctx is NULL");
  MMR_ARG(N64 >= 0 && N64 < (1LL << 31) && N64 % 3 == 0, "N out of range / not a multiple of 3");
  MMR_ARG(ld >= N64 && tgt_block >= 1 && nrows_local >= 0, "bad sizes");
  MMR_ARG(3 * tgt_block <= NB, "distributed LU needs 3*tgt_block <= 128");
  MMR_ARG(ctx->nranks == 1 || ctx->nccl_comm != nullptr, "communicator not initialised");
  if (info_out) *info_out = 0;
  if (N64 == 0) return 0;
  MMR_ARG(d_b && (d_Qlocal || nrows_local == 0), "NULL pointer");
  DeviceGuard g(ctx->device);
  const int N = (int)N64;
  const int R = ctx->nranks, me = ctx->rank;
  const int nbw = (int)(3 * tgt_block);
  const int S = (int)ceil_div64(N, nbw);
  {
    int64_t mine = 0;
    for (int s = me; s < S; s += R) mine += min(nbw, N - s * nbw);
    MMR_ARG(mine == nrows_local, "nrows_local does not match the block-cyclic ownership of this rank");
  }
  const int ncols_local = (int)nrows_local;
  cudaStream_t st = ctx->stream;

  const size_t scratch_bytes = (mmr_lu_panel_scratch_bytes() + 255) / 256 * 256;
  const size_t linv_bytes = sizeof(double) * NB * NB;
  const size_t ipiv_bytes = ((sizeof(int) * (size_t)N + 255) / 256) * 256;
  const size_t p_bytes = ((sizeof(double) * (size_t)N * nbw + 255) / 256) * 256;
  int rc = mmr_ws_reserve(ctx, 256 + scratch_bytes + 2 * linv_bytes + ipiv_bytes + 2 * p_bytes + sizeof(double) * (size_t)N + 512);
  if (rc) return rc;
  rc = mmr_hpin_reserve(ctx, 2 * sizeof(int) * (size_t)N + 256);
  if (rc) return rc;
  char* base = (char*)ctx->ws;
  int* d_info = (int*)base;
  void* scratch = base + 256;
  double* Linv2[2] = {(double*)(base + 256 + scratch_bytes), (double*)(base + 256 + scratch_bytes + linv_bytes)};
  int* d_ipiv = (int*)(base + 256 + scratch_bytes + 2 * linv_bytes);
  double* P2[2] = {(double*)(base + 256 + scratch_bytes + 2 * linv_bytes + ipiv_bytes),
                   (double*)(base + 256 + scratch_bytes + 2 * linv_bytes + ipiv_bytes + p_bytes)};
  double* d_tmp = (double*)(base + 256 + scratch_bytes + 2 * linv_bytes + ipiv_bytes + 2 * p_bytes);
  MMR_CUDA(cudaMemsetAsync(d_info, 0, sizeof(int), st));
  double* Al = d_Qlocal;
  cudaStream_t side = ctx->side;

  // factor global panel s (owned by this rank) on stream `strm`, invert its diagonal block and pack it
  auto factor_and_pack = [&](int s, cudaStream_t strm) -> int {
    const int kb = s * nbw, nb = min(nbw, N - kb), m = N - kb;
    double* Ash = Al + ((int64_t)(s / R) * nbw - kb) * ld;  // global column kb -> my local column
    int r = mmr_lu_panel(ctx, strm, Ash, ld, N, kb, nb, d_ipiv, d_info, scratch);
    if (r) return r;
    r = mmr_lu_trtri(ctx, strm, Ash + (int64_t)kb * ld + kb, ld, nb, Linv2[s & 1]);
    if (r) return r;
    MMR_CUDA(cudaMemcpy2DAsync(P2[s & 1], sizeof(double) * m, Ash + (int64_t)kb * ld + kb, sizeof(double) * ld,
                               sizeof(double) * m, nb, cudaMemcpyDeviceToDevice, strm));
    return 0;
  };
  // apply panel s (pivots, U12 = L11^-1 A12, A22 -= L21 U12) to my local columns [c0, c1)
  auto update_local = [&](int s, int c0, int c1) -> int {
    if (c1 <= c0) return 0;
    const int kb = s * nbw, nb = min(nbw, N - kb), m = N - kb;
    const double* P = P2[s & 1];
    int r = mmr_lu_laswp(ctx, st, Al, ld, kb, kb + nb, c0, c1, d_ipiv);
    if (r) return r;
    int ps = mmr_prof_begin(ctx, MMR_PROF_TRSM, st);
    r = mmr_dgemm_launch(ctx, st, nb, c1 - c0, nb, 1.0, Linv2[s & 1], nb, Al + (int64_t)c0 * ld + kb, ld, 0.0,
                         Al + (int64_t)c0 * ld + kb, ld);
    if (r) return r;
    mmr_prof_end(ctx, ps, 2.0 * nb * nb * (c1 - c0), st);
    if (m > nb) {
      ps = mmr_prof_begin(ctx, MMR_PROF_GEMM, st);
      r = mmr_dgemm_launch(ctx, st, m - nb, c1 - c0, nb, -1.0, P + nb, m, Al + (int64_t)c0 * ld + kb, ld, 1.0,
                           Al + (int64_t)c0 * ld + kb + nb, ld);
      if (r) return r;
      mmr_prof_end(ctx, ps, 2.0 * (m - nb) * (double)(c1 - c0) * nb, st);
    }
    return 0;
  };

  if (me == 0) {
    rc = factor_and_pack(0, st);
    if (rc) return rc;
  }
  for (int s = 0; s < S; ++s) {
    const int kb = s * nbw;
    const int nb = min(nbw, N - kb);
    const int m = N - kb;
    const int owner = s % R;
    if (R > 1) {
      rc = mmr_comm_bcast_bytes(ctx, P2[s & 1], sizeof(double) * (size_t)m * nb, owner, st);
      if (rc) return rc;
      rc = mmr_comm_bcast_bytes(ctx, Linv2[s & 1], sizeof(double) * (size_t)nb * nb, owner, st);
      if (rc) return rc;
      rc = mmr_comm_bcast_bytes(ctx, d_ipiv + kb, sizeof(int) * (size_t)nb, owner, st);
      if (rc) return rc;
    }
    // my already-factored panels (global index < s): row swaps only
    const int nleft = (s > me) ? ((s - 1 - me) / R + 1) : 0;
    rc = mmr_lu_laswp(ctx, st, Al, ld, kb, kb + nb, 0, nleft * nbw, d_ipiv);
    if (rc) return rc;
    // my trailing panels (global index > s)
    const int first_trail = (s >= me) ? ((s - me) / R + 1) : 0;
    const int ct0 = first_trail * nbw, ct1 = ncols_local;
    const bool own_next = (s + 1 < S) && ((s + 1) % R == me);
    if (own_next) {
      // look-ahead: bring my block of panel s+1 up to date first, factor it on the side stream while
      // the main stream updates the rest of my trailing columns
      const int nb1 = min(nbw, N - (s + 1) * nbw);
      rc = update_local(s, ct0, ct0 + nb1);
      if (rc) return rc;
      MMR_CUDA(cudaEventRecord(ctx->ev[0], st));
      MMR_CUDA(cudaStreamWaitEvent(side, ctx->ev[0], 0));
      rc = factor_and_pack(s + 1, side);
      if (rc) return rc;
      MMR_CUDA(cudaEventRecord(ctx->ev[1], side));
      rc = update_local(s, ct0 + nb1, ct1);
      if (rc) return rc;
      MMR_CUDA(cudaStreamWaitEvent(st, ctx->ev[1], 0));
    } else {
      rc = update_local(s, ct0, ct1);
      if (rc) return rc;
    }
  }

  // ---- triangular solves, panel by panel --------------------------------------------------
  const size_t diag_smem = (size_t)SB * (SB + 1) * sizeof(double);
  MMR_CUDA(cudaFuncSetAttribute(lu_diag_solve_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)diag_smem));
  MMR_CUDA(cudaFuncSetAttribute(lu_diag_solve_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)diag_smem));
  for (int s = 0; s < S; ++s) {  // U^T w = b
    const int kb = s * nbw, nb = min(nbw, N - kb), owner = s % R;
    if (me == owner) {
      const double* Ash = Al + ((int64_t)(s / R) * nbw - kb) * ld;
      if (kb > 0) {
        lu_gemv_t_block_kernel<<<nb, 256, 0, st>>>(Ash, ld, 0, kb, kb, d_b, d_b);
        MMR_CHECK_LAUNCH(ctx);
      }
      lu_diag_solve_kernel<true><<<1, SB, diag_smem, st>>>(Ash, ld, kb, kb + nb, d_b);
      MMR_CHECK_LAUNCH(ctx);
    }
    if (R > 1) {
      rc = mmr_comm_bcast_bytes(ctx, d_b + kb, sizeof(double) * nb, owner, st);
      if (rc) return rc;
    }
  }
  for (int s = S - 1; s >= 0; --s) {  // L^T v = w
    const int kb = s * nbw, nb = min(nbw, N - kb), owner = s % R;
    if (me == owner) {
      const double* Ash = Al + ((int64_t)(s / R) * nbw - kb) * ld;
      if (kb + nb < N) {
        lu_gemv_t_block_kernel<<<nb, 256, 0, st>>>(Ash, ld, kb + nb, N, kb, d_b, d_b);
        MMR_CHECK_LAUNCH(ctx);
      }
      lu_diag_solve_kernel<false><<<1, SB, diag_smem, st>>>(Ash, ld, kb, kb + nb, d_b);
      MMR_CHECK_LAUNCH(ctx);
    }
    if (R > 1) {
      rc = mmr_comm_bcast_bytes(ctx, d_b + kb, sizeof(double) * nb, owner, st);
      if (rc) return rc;
    }
  }
  // ---- x = P v (pivots are replicated) ------------------------------------------------------
  int* h_ipiv = (int*)ctx->hpin;
  int* h_perm = h_ipiv + N;
  MMR_CUDA(cudaMemcpyAsync(h_ipiv, d_ipiv, sizeof(int) * N, cudaMemcpyDeviceToHost, st));
  MMR_CUDA(cudaMemcpyAsync(h_perm, d_info, sizeof(int), cudaMemcpyDeviceToHost, st));
  MMR_CUDA(cudaStreamSynchronize(st));
  const int info = h_perm[0];
  for (int i = 0; i < N; ++i) h_perm[i] = i;
  for (int i = N - 1; i >= 0; --i) {
    const int p = h_ipiv[i];
    if (p != i) {
      const int t = h_perm[i];
      h_perm[i] = h_perm[p];
      h_perm[p] = t;
    }
  }
  int* d_perm = d_ipiv;  // pivots no longer needed on the device
  MMR_CUDA(cudaMemcpyAsync(d_perm, h_perm, sizeof(int) * N, cudaMemcpyHostToDevice, st));
  MMR_CUDA(cudaMemcpyAsync(d_tmp, d_b, sizeof(double) * N, cudaMemcpyDeviceToDevice, st));
  lu_gather_kernel<<<(unsigned)ceil_div64(N, 256), 256, 0, st>>>(d_tmp, d_perm, N, d_b);
  MMR_CHECK_LAUNCH(ctx);
  MMR_CUDA(cudaStreamSynchronize(st));
  if (info_out) *info_out = info;  // only the owner of a zero pivot sees it; callers reduce over ranks
  return 0;
}
