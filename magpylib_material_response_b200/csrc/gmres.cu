// Kernel 2b: restarted GMRES on the assembled Q with an iterative-refinement (true residual)
// check at every restart -- the iterative alternative to np.linalg.solve (demag.py:470).
//
// The only O(N^2) work per iteration is the row-block matvec y = Q x, which streams Q once from
// HBM (8 N^2 bytes, the roofline of this solver).  Orthogonalization is classical Gram-Schmidt
// applied twice (CGS2): two multi-dot kernels + two block-axpy kernels over the Krylov basis.
// The (restart+1) x restart Hessenberg least-squares problem lives on the host (Givens).
//
// Multi-GPU: each rank owns a block-cyclic set of rows; the local y slice is all-gathered with
// NCCL and un-permuted to global order; everything else is replicated, so no other collective.
#include <cmath>
#include <vector>

#include "comm.h"
#include "common.cuh"

namespace {

constexpr int MV_THREADS = 128;

// one CTA per row; row-major Q
template <bool VEC2>
__global__ void __launch_bounds__(MV_THREADS)
    matvec_rows_kernel(int64_t nrows, int64_t ncols, const double* __restrict__ Q, int64_t ld,
                       const double* __restrict__ x, double* __restrict__ y) {
  __shared__ double s_part[MV_THREADS / 32];
  const int64_t r = blockIdx.x;
  if (r >= nrows) return;
  const double* q = Q + r * ld;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  if (VEC2) {
    const double2* q2 = reinterpret_cast<const double2*>(q);
    const double2* x2 = reinterpret_cast<const double2*>(x);
    const int64_t n2 = ncols >> 1;
    int64_t c = threadIdx.x;
    for (; c + 3 * MV_THREADS < n2; c += 4 * MV_THREADS) {
      const double2 qa = __ldcs(q2 + c), qb = __ldcs(q2 + c + MV_THREADS);
      const double2 qc = __ldcs(q2 + c + 2 * MV_THREADS), qd = __ldcs(q2 + c + 3 * MV_THREADS);
      const double2 xa = x2[c], xb = x2[c + MV_THREADS], xc = x2[c + 2 * MV_THREADS],
                    xd = x2[c + 3 * MV_THREADS];
      a0 += qa.x * xa.x + qa.y * xa.y;
      a1 += qb.x * xb.x + qb.y * xb.y;
      a2 += qc.x * xc.x + qc.y * xc.y;
      a3 += qd.x * xd.x + qd.y * xd.y;
    }
    for (; c < n2; c += MV_THREADS) {
      const double2 qa = __ldcs(q2 + c);
      const double2 xa = x2[c];
      a0 += qa.x * xa.x + qa.y * xa.y;
    }
    if ((ncols & 1) && threadIdx.x == 0) a1 += q[ncols - 1] * x[ncols - 1];
  } else {
    for (int64_t c = threadIdx.x; c < ncols; c += MV_THREADS) a0 += q[c] * x[c];
  }
  double acc = (a0 + a1) + (a2 + a3);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int q8 = 0; q8 < MV_THREADS / 32; ++q8) t += s_part[q8];
    y[r] = t;
  }
}

// out[i] = V_i . w  for i in [0, nvec)   (V column-major N x *, contiguous basis vectors)
// if also_ww, out[nvec] = w . w
__global__ void __launch_bounds__(512)
    multi_dot_kernel(int64_t N, const double* __restrict__ V, int nvec, const double* __restrict__ w,
                     double* __restrict__ out) {
  __shared__ double s_part[16];
  const int i = blockIdx.x;
  const double* v = (i < nvec) ? (V + (int64_t)i * N) : w;
  double acc = 0.0;
  for (int64_t e = threadIdx.x; e < N; e += 512) acc += v[e] * w[e];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int q = 0; q < 16; ++q) t += s_part[q];
    out[i] = t;
  }
}

// w -= sum_i h[i] V_i ; optionally hsum[i] += h[i]
__global__ void block_axpy_kernel(int64_t N, const double* __restrict__ V, int nvec,
                                  const double* __restrict__ h, double sign, double* __restrict__ w) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N) return;
  double acc = 0.0;
  for (int i = 0; i < nvec; ++i) acc += h[i] * V[(int64_t)i * N + e];
  w[e] += sign * acc;
}

__global__ void scale_copy_kernel(int64_t N, const double* __restrict__ src, const double* __restrict__ nrm2,
                                  double* __restrict__ dst) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N) return;
  const double s = 1.0 / sqrt(*nrm2);
  dst[e] = src[e] * s;
}

__global__ void residual_kernel(int64_t N, const double* __restrict__ b, const double* __restrict__ qx,
                                double* __restrict__ r) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < N) r[e] = b[e] - qx[e];
}

__global__ void add_vec_kernel(int n, const double* __restrict__ a, double* __restrict__ acc) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) acc[e] += a[e];
}

// gathered (rank-major, maxrows per rank, local row order) -> global row order
__global__ void unpermute_kernel(int64_t N, int64_t tgt_block, int nranks, int64_t maxrows,
                                 const double* __restrict__ gathered, double* __restrict__ y) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= N) return;
  const int64_t cell = g / 3, comp = g % 3;
  const int64_t blk = cell / tgt_block;
  const int64_t rank = blk % nranks;
  const int64_t lb = blk / nranks;
  const int64_t lrow = (lb * tgt_block + cell % tgt_block) * 3 + comp;
  y[g] = gathered[rank * maxrows + lrow];
}

// Block-Jacobi right preconditioner M = blockdiag(D_b): the diagonal blocks of Q of selected row ranges (one
// magnet each), LU-factorized with the DMMA LU; identity on the other rows.  GMRES then iterates on
// Q M^-1 u = b, x = M^-1 u; the true residual b - Q x is unaffected.
struct BjBlock {
  int64_t r0, nb, ld;
  double* D;      // nb x ld copy of Q[r0:r0+nb, r0:r0+nb], factorized in place
  int32_t* ipiv;  // nb
};
struct Precond {
  std::vector<BjBlock> blocks;
};

// z = M^-1 v (z may alias v)
int apply_precond(mmr_ctx* ctx, const Precond& pc, int64_t N, const double* v, double* z) {
  if (z != v) MMR_CUDA(cudaMemcpyAsync(z, v, sizeof(double) * N, cudaMemcpyDeviceToDevice, ctx->stream));
  for (const BjBlock& b : pc.blocks) {
    const int rc = mmr_lu_solve(ctx, b.nb, b.D, b.ld, b.ipiv, z + b.r0, 1);
    if (rc) return rc;
  }
  return 0;
}

struct Gm {
  mmr_ctx* ctx;
  int64_t N, nrows_local, tgt_block, ld, maxrows;
  const double* Q;
  double* ylocal;
  double* gathered;
  const Precond* pc = nullptr;
  double* ptmp = nullptr;  // length N scratch of the preconditioned matvec
};

int launch_matvec(mmr_ctx* ctx, int64_t nrows, int64_t ncols, const double* Q, int64_t ld, const double* x,
                  double* y) {
  if (nrows == 0) return 0;
  const bool vec2 = (((uintptr_t)Q | (uintptr_t)x) % 16 == 0) && (ld % 2 == 0);
  const int pidx = mmr_prof_begin(ctx, MMR_PROF_MATVEC, ctx->stream);
  if (vec2)
    matvec_rows_kernel<true><<<(unsigned)nrows, MV_THREADS, 0, ctx->stream>>>(nrows, ncols, Q, ld, x, y);
  else
    matvec_rows_kernel<false><<<(unsigned)nrows, MV_THREADS, 0, ctx->stream>>>(nrows, ncols, Q, ld, x, y);
  mmr_prof_end(ctx, pidx, 8.0 * nrows * (double)ncols, ctx->stream);
  MMR_CHECK_LAUNCH(ctx);
  return 0;
}

// y (global order, length N) = Q x using this rank's rows (+ all-gather when distributed)
int apply_Q(const Gm& g, const double* x, double* y) {
  mmr_ctx* ctx = g.ctx;
  if (ctx->nranks == 1 && g.nrows_local == g.N) return launch_matvec(ctx, g.N, g.N, g.Q, g.ld, x, y);
  int rc = launch_matvec(ctx, g.nrows_local, g.N, g.Q, g.ld, x, g.ylocal);
  if (rc) return rc;
  rc = mmr_comm_allgather_f64(ctx, g.ylocal, g.gathered, (size_t)g.maxrows, ctx->stream);
  if (rc) return rc;
  unpermute_kernel<<<(unsigned)ceil_div64(g.N, 256), 256, 0, ctx->stream>>>(g.N, g.tgt_block, ctx->nranks,
                                                                            g.maxrows, g.gathered, y);
  MMR_CHECK_LAUNCH(ctx);
  return 0;
}

int gmres_impl(mmr_ctx* ctx, int64_t N, int64_t nrows_local, int64_t tgt_block, const double* Q, int64_t ld,
               const double* d_b, double* d_x, double tol, int restart, int maxit, int* iters_out,
               double* relres_out, const Precond* pc = nullptr) {
  if (restart < 1) restart = 1;
  if (restart > maxit) restart = maxit > 0 ? maxit : 1;
  const int m = restart;
  Gm g;
  g.ctx = ctx; g.N = N; g.nrows_local = nrows_local; g.tgt_block = tgt_block; g.ld = ld; g.Q = Q;
  const int R = ctx->nranks;
  const int64_t ncell = N / 3;
  const int64_t nblk = tgt_block > 0 ? ceil_div64(ncell, tgt_block) : 1;
  g.maxrows = 3 * tgt_block * ceil_div64(nblk, R);
  // workspace layout (doubles): V[(m+1)*N] w[N] r[N] ylocal[maxrows] gathered[R*maxrows] hd[2*(m+2)] yd[m]
  const size_t nV = (size_t)(m + 1) * N;
  size_t off = 0;
  auto take = [&](size_t cnt) { size_t o = off; off += (cnt + 31) & ~(size_t)31; return o; };
  const size_t oV = take(nV), ow = take(N), orr = take(N), oyl = take(g.maxrows), og = take((size_t)R * g.maxrows),
               oh = take(m + 2), oh2 = take(m + 2), oy = take(m + 1), opt = take(pc ? N : 0);
  int rc = mmr_ws2_reserve(ctx, off * sizeof(double));
  if (rc) return rc;
  // (with a preconditioner, mmr_lu_solve keeps pivots + permutation of its block in the same pinned buffer:
  // reserve the larger need once, so that the buffer never moves under `hh`)
  rc = mmr_hpin_reserve(ctx, std::max(sizeof(double) * (size_t)(2 * m + 16), 2 * sizeof(int) * (size_t)N + 512));
  if (rc) return rc;
  double* base = (double*)ctx->ws2;
  double *V = base + oV, *w = base + ow, *r = base + orr, *hd = base + oh, *hd2 = base + oh2, *yd = base + oy;
  g.ylocal = base + oyl;
  g.gathered = base + og;
  g.pc = pc;
  g.ptmp = base + opt;
  double* hh = (double*)ctx->hpin;  // pinned host mirror of hd (m+2 entries)
  cudaStream_t s = ctx->stream;
  const unsigned gridN = (unsigned)ceil_div64(N, 256);

  // ||b||
  multi_dot_kernel<<<1, 512, 0, s>>>(N, d_b, 0, d_b, hd);
  MMR_CHECK_LAUNCH(ctx);
  MMR_CUDA(cudaMemcpyAsync(hh, hd, sizeof(double), cudaMemcpyDeviceToHost, s));
  MMR_CUDA(cudaStreamSynchronize(s));
  const double bnorm = std::sqrt(hh[0]);
  int iters = 0;
  double relres = 0.0;
  if (bnorm == 0.0) {
    MMR_CUDA(cudaMemsetAsync(d_x, 0, sizeof(double) * N, s));
    if (iters_out) *iters_out = 0;
    if (relres_out) *relres_out = 0.0;
    return 0;
  }

  std::vector<double> H((size_t)(m + 1) * m, 0.0), cs(m), sn(m), gv(m + 1), yv(m);
  bool converged = false;
  double prev_relres = -1.0;
  int stalled = 0;
  while (true) {
    // true residual r = b - Q x  (iterative-refinement check)
    rc = apply_Q(g, d_x, w);
    if (rc) return rc;
    residual_kernel<<<gridN, 256, 0, s>>>(N, d_b, w, r);
    MMR_CHECK_LAUNCH(ctx);
    multi_dot_kernel<<<1, 512, 0, s>>>(N, r, 0, r, hd);
    MMR_CHECK_LAUNCH(ctx);
    MMR_CUDA(cudaMemcpyAsync(hh, hd, sizeof(double), cudaMemcpyDeviceToHost, s));
    MMR_CUDA(cudaStreamSynchronize(s));
    const double beta = std::sqrt(hh[0]);
    relres = beta / bnorm;
    if (relres <= tol) {
      converged = true;
      break;
    }
    if (iters >= maxit) break;
    // stagnation: the true residual no longer follows the implicit one (attainable accuracy reached above tol);
    // two restart cycles without a 10 % gain end the run instead of grinding to maxit
    if (prev_relres > 0.0 && relres > 0.9 * prev_relres) {
      if (++stalled >= 2) break;
    } else {
      stalled = 0;
    }
    prev_relres = relres;
    scale_copy_kernel<<<gridN, 256, 0, s>>>(N, r, hd, V);  // V_0 = r / beta
    MMR_CHECK_LAUNCH(ctx);
    std::fill(gv.begin(), gv.end(), 0.0);
    gv[0] = beta;
    int j = 0;
    for (; j < m && iters < maxit; ++j) {
      ++iters;
      if (pc) {
        if ((rc = apply_precond(ctx, *pc, N, V + (size_t)j * N, g.ptmp))) return rc;
        rc = apply_Q(g, g.ptmp, w);
      } else {
        rc = apply_Q(g, V + (size_t)j * N, w);
      }
      if (rc) return rc;
      // CGS2
      multi_dot_kernel<<<j + 1, 512, 0, s>>>(N, V, j + 1, w, hd);
      MMR_CHECK_LAUNCH(ctx);
      block_axpy_kernel<<<gridN, 256, 0, s>>>(N, V, j + 1, hd, -1.0, w);
      MMR_CHECK_LAUNCH(ctx);
      multi_dot_kernel<<<j + 1, 512, 0, s>>>(N, V, j + 1, w, hd2);
      MMR_CHECK_LAUNCH(ctx);
      block_axpy_kernel<<<gridN, 256, 0, s>>>(N, V, j + 1, hd2, -1.0, w);
      MMR_CHECK_LAUNCH(ctx);
      add_vec_kernel<<<(unsigned)ceil_div64(j + 1, 128), 128, 0, s>>>(j + 1, hd2, hd);
      MMR_CHECK_LAUNCH(ctx);
      multi_dot_kernel<<<1, 512, 0, s>>>(N, w, 0, w, hd + (j + 1));  // ||w||^2
      MMR_CHECK_LAUNCH(ctx);
      scale_copy_kernel<<<gridN, 256, 0, s>>>(N, w, hd + (j + 1), V + (size_t)(j + 1) * N);
      MMR_CHECK_LAUNCH(ctx);
      MMR_CUDA(cudaMemcpyAsync(hh, hd, sizeof(double) * (j + 2), cudaMemcpyDeviceToHost, s));
      MMR_CUDA(cudaStreamSynchronize(s));
      // Hessenberg column j
      double* Hj = &H[(size_t)j * (m + 1)];
      for (int i = 0; i <= j; ++i) Hj[i] = hh[i];
      Hj[j + 1] = std::sqrt(hh[j + 1]);
      for (int i = 0; i < j; ++i) {
        const double t = cs[i] * Hj[i] + sn[i] * Hj[i + 1];
        Hj[i + 1] = -sn[i] * Hj[i] + cs[i] * Hj[i + 1];
        Hj[i] = t;
      }
      const double denom = std::hypot(Hj[j], Hj[j + 1]);
      cs[j] = denom == 0.0 ? 1.0 : Hj[j] / denom;
      sn[j] = denom == 0.0 ? 0.0 : Hj[j + 1] / denom;
      Hj[j] = cs[j] * Hj[j] + sn[j] * Hj[j + 1];
      Hj[j + 1] = 0.0;
      gv[j + 1] = -sn[j] * gv[j];
      gv[j] = cs[j] * gv[j];
      if (std::fabs(gv[j + 1]) / bnorm <= tol) {
        ++j;
        break;
      }
    }
    // solve the j x j triangular system, x += V y
    const int k = j;
    for (int i = k - 1; i >= 0; --i) {
      double t = gv[i];
      for (int c = i + 1; c < k; ++c) t -= H[(size_t)c * (m + 1) + i] * yv[c];
      const double hii = H[(size_t)i * (m + 1) + i];
      yv[i] = hii != 0.0 ? t / hii : 0.0;
    }
    for (int i = 0; i < k; ++i) hh[i] = yv[i];
    MMR_CUDA(cudaMemcpyAsync(yd, hh, sizeof(double) * k, cudaMemcpyHostToDevice, s));
    if (pc) {
      MMR_CUDA(cudaStreamSynchronize(s));  // yd has left the pinned buffer before the block solves reuse it
      MMR_CUDA(cudaMemsetAsync(g.ptmp, 0, sizeof(double) * N, s));
      block_axpy_kernel<<<gridN, 256, 0, s>>>(N, V, k, yd, 1.0, g.ptmp);
      MMR_CHECK_LAUNCH(ctx);
      if ((rc = apply_precond(ctx, *pc, N, g.ptmp, g.ptmp))) return rc;
      add_vec_kernel<<<gridN, 256, 0, s>>>((int)N, g.ptmp, d_x);
      MMR_CHECK_LAUNCH(ctx);
    } else {
      block_axpy_kernel<<<gridN, 256, 0, s>>>(N, V, k, yd, 1.0, d_x);
      MMR_CHECK_LAUNCH(ctx);
    }
    MMR_CUDA(cudaStreamSynchronize(s));  // hh is reused next round
  }
  if (iters_out) *iters_out = iters;
  if (relres_out) *relres_out = relres;
  if (!converged) {
    mmr_set_error("GMRES did not converge: relres %.3e after %d iterations (tol %.1e)", relres, iters, tol);
    return -4;
  }
  return 0;
}

}  // namespace

extern "C" int mmr_matvec(mmr_ctx* ctx, int64_t nrows, int64_t ncols, const double* d_Q, int64_t ld,
                          const double* d_x, double* d_y) {
  MMR_ARG(ctx != nullptr, "ctx is NULL");
  MMR_ARG(nrows >= 0 && ncols >= 0 && ld >= ncols, "bad sizes");
  if (nrows == 0) return 0;
  MMR_ARG(d_Q && d_x && d_y, "NULL pointer");
  DeviceGuard g(ctx->device);
  return launch_matvec(ctx, nrows, ncols, d_Q, ld, d_x, d_y);
}

extern "C" int mmr_gmres(mmr_ctx* ctx, int64_t N, const double* d_Q, int64_t ld, const double* d_b,
                         double* d_x, double tol, int restart, int maxit, int* iters_out,
                         double* relres_out) {
  MMR_ARG(ctx != nullptr, "ctx is NULL");
  MMR_ARG(N >= 0 && ld >= N, "bad N/ld");
  if (iters_out) *iters_out = 0;
  if (relres_out) *relres_out = 0.0;
  if (N == 0) return 0;
  MMR_ARG(d_Q && d_b && d_x, "NULL pointer");
  MMR_ARG(tol > 0 && maxit >= 0, "bad tol/maxit");
  DeviceGuard g(ctx->device);
  const int saved = ctx->nranks;
  ctx->nranks = 1;  // single-GPU entry point: never touches the communicator
  const int rc = gmres_impl(ctx, N, N, (N + 2) / 3, d_Q, ld, d_b, d_x, tol, restart, maxit, iters_out, relres_out);
  ctx->nranks = saved;
  return rc;
}

extern "C" int mmr_gmres_bj(mmr_ctx* ctx, int64_t N, const double* d_Q, int64_t ld, const double* d_b, double* d_x,
                            double tol, int restart, int maxit, int nblocks, const int64_t* block_row0,
                            const int64_t* block_rows, int* iters_out, double* relres_out) {
  MMR_ARG(ctx != nullptr, "ctx is NULL");
  MMR_ARG(N >= 0 && ld >= N, "bad N/ld");
  if (iters_out) *iters_out = 0;
  if (relres_out) *relres_out = 0.0;
  if (N == 0) return 0;
  MMR_ARG(d_Q && d_b && d_x, "NULL pointer");
  MMR_ARG(tol > 0 && maxit >= 0, "bad tol/maxit");
  MMR_ARG(nblocks >= 0 && (nblocks == 0 || (block_row0 && block_rows)), "bad block list");
  DeviceGuard g(ctx->device);
  Precond pc;
  int rc = 0;
  auto release = [&]() {
    for (BjBlock& b : pc.blocks) {
      if (b.D) cudaFree(b.D);
      if (b.ipiv) cudaFree(b.ipiv);
    }
  };
  for (int i = 0; i < nblocks && rc == 0; ++i) {
    BjBlock b;
    b.r0 = block_row0[i];
    b.nb = block_rows[i];
    if (b.nb <= 0) continue;
    if (b.r0 < 0 || b.r0 + b.nb > N) {
      mmr_set_error("invalid argument: preconditioner block %d outside the matrix", i);
      rc = -1;
      break;
    }
    b.ld = (b.nb + 15) / 16 * 16;
    b.D = nullptr;
    b.ipiv = nullptr;
    cudaError_t e = cudaMalloc((void**)&b.D, sizeof(double) * (size_t)b.nb * b.ld);
    if (e == cudaSuccess) e = cudaMalloc((void**)&b.ipiv, sizeof(int32_t) * (size_t)b.nb);
    pc.blocks.push_back(b);
    if (e != cudaSuccess) {
      rc = mmr_fail_cuda(e, "allocate preconditioner block", __FILE__, __LINE__);
      break;
    }
    e = cudaMemcpy2DAsync(b.D, sizeof(double) * b.ld, d_Q + b.r0 * ld + b.r0, sizeof(double) * ld, sizeof(double) * b.nb,
                          (size_t)b.nb, cudaMemcpyDeviceToDevice, ctx->stream);
    if (e != cudaSuccess) {
      rc = mmr_fail_cuda(e, "copy preconditioner block", __FILE__, __LINE__);
      break;
    }
    int info = 0;
    rc = mmr_lu_factor(ctx, b.nb, b.D, b.ld, b.ipiv, &info);
    if (rc == 0 && info != 0) {
      mmr_set_error("block-Jacobi preconditioner: diagonal block %d is singular (pivot %d)", i, info);
      rc = -8;
    }
  }
  if (rc == 0) {
    const int saved = ctx->nranks;
    ctx->nranks = 1;
    rc = gmres_impl(ctx, N, N, (N + 2) / 3, d_Q, ld, d_b, d_x, tol, restart, maxit, iters_out, relres_out,
                    pc.blocks.empty() ? nullptr : &pc);
    ctx->nranks = saved;
  }
  cudaStreamSynchronize(ctx->stream);
  release();
  return rc;
}

extern "C" int mmr_gmres_dist(mmr_ctx* ctx, int64_t N, int64_t nrows_local, int64_t tgt_block,
                              const double* d_Qlocal, int64_t ld, const double* d_b, double* d_x,
                              double tol, int restart, int maxit, int* iters_out, double* relres_out) {
  MMR_ARG(ctx != nullptr, "ctx is NULL");
  MMR_ARG(N >= 0 && N % 3 == 0 && ld >= N && nrows_local >= 0 && tgt_block >= 1, "bad sizes");
  if (iters_out) *iters_out = 0;
  if (relres_out) *relres_out = 0.0;
  if (N == 0) return 0;
  MMR_ARG(d_b && d_x && (d_Qlocal || nrows_local == 0), "NULL pointer");
  MMR_ARG(tol > 0 && maxit >= 0, "bad tol/maxit");
  MMR_ARG(ctx->nranks == 1 || ctx->nccl_comm != nullptr, "communicator not initialised");
  DeviceGuard g(ctx->device);
  return gmres_impl(ctx, N, nrows_local, tgt_block, d_Qlocal, ld, d_b, d_x, tol, restart, maxit, iters_out,
                    relres_out);
}
