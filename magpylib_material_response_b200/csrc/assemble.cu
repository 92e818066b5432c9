// Kernel 1: dense demag-tensor assembly (replaces the 3 x magpy.getH passes of
// demag.py:182-222, the mu0 scaling/permutation of :451-452 and the dense S@T of :466).
//
// Work decomposition (cell-major layout): a CTA owns a tile of SRC_TILE consecutive source
// cells (one per thread, geometry in registers) x TGT_TILE target cells (geometry staged in
// shared memory and read as warp-uniform broadcasts).  Each thread evaluates the full 3x3
// block of its (source, target) pair once; a warp's 32 blocks form three 96-double row
// segments which are transposed through shared memory so every global store instruction
// writes 256 contiguous bytes (full 32 B sectors).  HBM traffic = 72 B written per pair, the
// source/target geometry (10 doubles per cell) is read once per CTA and stays in L2.
#include "cuboid.cuh"

namespace {

constexpr int SRC_TILE = 128;  // threads per CTA, one source cell each
constexpr int TGT_TILE = 32;   // target cells per CTA

struct AsmParams {
  int64_t n;
  const double* pos;
  const double* quat;
  const double* dim;
  int64_t src_begin, src_end;
  int64_t tgt_first, n_tgt_local, tgt_block, tgt_nparts, tgt_part;
  const double* chi;  // cell-major [3*j+l] or null
  double max_dist;
  double* out;
  int64_t ld;
};

__device__ __forceinline__ int64_t local_to_global_target(const AsmParams& p, int64_t t) {
  return p.tgt_first + ((t / p.tgt_block) * p.tgt_nparts + p.tgt_part) * p.tgt_block +
         t % p.tgt_block;
}

// max_dist predicate, evaluated exactly as demag.py:245-249 does:
// norm(p_j - p_i) / max(dim_j U dim_i) < max_dist   (no FMA contraction, same add order)
__device__ __forceinline__ bool pair_kept(double dx, double dy, double dz, double maxdim,
                                          double max_dist) {
  const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
  return __ddiv_rn(sqrt(d2), maxdim) < max_dist;
}

template <bool ROTATED, bool FUSE_Q, bool MAXDIST>
__global__ void __launch_bounds__(SRC_TILE)
    assemble_cell_major_kernel(const AsmParams p) {
  __shared__ double s_tpos[TGT_TILE][3];
  __shared__ double s_tmaxdim[TGT_TILE];
  __shared__ double s_tchi[TGT_TILE][3];
  __shared__ int64_t s_tglob[TGT_TILE];
  __shared__ double s_stage[SRC_TILE / 32][3][96];

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int64_t t0 = (int64_t)blockIdx.y * TGT_TILE;
  const int ntgt = (int)min((int64_t)TGT_TILE, p.n_tgt_local - t0);

  if (tid < ntgt) {
    const int64_t j = local_to_global_target(p, t0 + tid);
    s_tglob[tid] = j;
    s_tpos[tid][0] = p.pos[3 * j + 0];
    s_tpos[tid][1] = p.pos[3 * j + 1];
    s_tpos[tid][2] = p.pos[3 * j + 2];
    if (MAXDIST)
      s_tmaxdim[tid] = fmax(fmax(p.dim[3 * j + 0], p.dim[3 * j + 1]), p.dim[3 * j + 2]);
    if (FUSE_Q) {
      s_tchi[tid][0] = p.chi[3 * j + 0];
      s_tchi[tid][1] = p.chi[3 * j + 1];
      s_tchi[tid][2] = p.chi[3 * j + 2];
    }
  }

  const int64_t i_warp0 = p.src_begin + (int64_t)blockIdx.x * SRC_TILE + warp * 32;
  const int64_t i = i_warp0 + lane;
  const bool active = i < p.src_end;
  const int64_t ic = active ? i : p.src_end - 1;  // clamp: inactive lanes compute a valid pair
  const double px = p.pos[3 * ic + 0], py = p.pos[3 * ic + 1], pz = p.pos[3 * ic + 2];
  const double d0 = p.dim[3 * ic + 0], d1 = p.dim[3 * ic + 1], d2 = p.dim[3 * ic + 2];
  const double ha = 0.5 * d0, hb = 0.5 * d1, hc = 0.5 * d2;
  const double smaxdim = fmax(fmax(d0, d1), d2);
  Rot3 R;
  if (ROTATED) quat_to_rot(p.quat + 4 * ic, R);
  __syncthreads();

  // number of valid columns (doubles) in this warp's 96-wide row segment
  const int64_t ncols_warp = min((int64_t)96, 3 * (p.src_end - i_warp0));
  const int64_t col0 = 3 * (i_warp0 - p.src_begin) + 3 * p.src_begin;  // = 3*i_warp0

  for (int tt = 0; tt < ntgt; ++tt) {
    const double dx = s_tpos[tt][0] - px, dy = s_tpos[tt][1] - py, dz = s_tpos[tt][2] - pz;
    double blk[9];
    bool keep = true;
    if (MAXDIST) keep = pair_kept(dx, dy, dz, fmax(smaxdim, s_tmaxdim[tt]), p.max_dist);
    if (keep) {
      SymBlock N;
      if (ROTATED) {
        const double lx = R.m[0] * dx + R.m[3] * dy + R.m[6] * dz;  // R^T d
        const double ly = R.m[1] * dx + R.m[4] * dy + R.m[7] * dz;
        const double lz = R.m[2] * dx + R.m[5] * dy + R.m[8] * dz;
        cuboid_block_auto<true>(lx, ly, lz, ha, hb, hc, N);
        rotate_block(R, N, blk);
      } else {
        cuboid_block_auto<true>(dx, dy, dz, ha, hb, hc, N);
        blk[0] = N.xx; blk[1] = N.xy; blk[2] = N.xz;
        blk[3] = N.xy; blk[4] = N.yy; blk[5] = N.yz;
        blk[6] = N.xz; blk[7] = N.yz; blk[8] = N.zz;
      }
    } else {
#pragma unroll
      for (int q = 0; q < 9; ++q) blk[q] = 0.0;
    }
    if (FUSE_Q) {
      const bool diag = (s_tglob[tt] == i);
#pragma unroll
      for (int l = 0; l < 3; ++l) {
        const double chi = s_tchi[tt][l];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const double v = -(chi * blk[3 * l + k]);  // same single product as S@T (Appendix A Q8)
          blk[3 * l + k] = (diag && l == k) ? (1.0 + v) : v;
        }
      }
    }
    // transpose through shared memory: thread-major (3 doubles per lane) -> lane-major
    __syncwarp();
#pragma unroll
    for (int l = 0; l < 3; ++l)
#pragma unroll
      for (int k = 0; k < 3; ++k) s_stage[warp][l][3 * lane + k] = blk[3 * l + k];
    __syncwarp();
    double* row_base = p.out + (3 * (t0 + tt)) * p.ld + col0;
#pragma unroll
    for (int l = 0; l < 3; ++l) {
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int c = q * 32 + lane;
        if (c < ncols_warp) row_base[l * p.ld + c] = s_stage[warp][l][c];
      }
    }
  }
}

// Reference layout (demag.py:224): out[((k*n + i)*n + j)*3 + l], A/m per tesla.  Lanes run
// over TARGET cells here (j is the contiguous index), sources are the uniform loop.
template <bool ROTATED, bool MAXDIST>
__global__ void __launch_bounds__(SRC_TILE)
    assemble_reference_kernel(const AsmParams p) {
  __shared__ double s_spos[TGT_TILE][3];
  __shared__ double s_sdim[TGT_TILE][3];
  __shared__ double s_squat[TGT_TILE][4];
  __shared__ double s_stage[SRC_TILE / 32][3][96];

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int64_t i0 = p.src_begin + (int64_t)blockIdx.y * TGT_TILE;  // source tile (uniform loop)
  const int nsrc = (int)min((int64_t)TGT_TILE, p.src_end - i0);
  if (tid < nsrc) {
    const int64_t i = i0 + tid;
    for (int q = 0; q < 3; ++q) {
      s_spos[tid][q] = p.pos[3 * i + q];
      s_sdim[tid][q] = p.dim[3 * i + q];
    }
    for (int q = 0; q < 4; ++q) s_squat[tid][q] = p.quat[4 * i + q];
  }
  const int64_t j_warp0 = (int64_t)blockIdx.x * SRC_TILE + warp * 32;
  const int64_t j = j_warp0 + lane;
  const int64_t jc = j < p.n ? j : p.n - 1;
  const double tx = p.pos[3 * jc + 0], ty = p.pos[3 * jc + 1], tz = p.pos[3 * jc + 2];
  const double tmaxdim = fmax(fmax(p.dim[3 * jc + 0], p.dim[3 * jc + 1]), p.dim[3 * jc + 2]);
  __syncthreads();
  const int64_t ncols_warp = min((int64_t)96, 3 * (p.n - j_warp0));
  const double inv_mu0 = 1.0 / MMR_MU0;

  for (int ss = 0; ss < nsrc; ++ss) {
    const int64_t i = i0 + ss;
    const double dx = tx - s_spos[ss][0], dy = ty - s_spos[ss][1], dz = tz - s_spos[ss][2];
    const double d0 = s_sdim[ss][0], d1 = s_sdim[ss][1], d2 = s_sdim[ss][2];
    double blk[9];
    bool keep = true;
    if (MAXDIST) keep = pair_kept(dx, dy, dz, fmax(fmax(fmax(d0, d1), d2), tmaxdim), p.max_dist);
    if (keep) {
      SymBlock N;
      if (ROTATED) {
        Rot3 R;
        quat_to_rot(&s_squat[ss][0], R);
        const double lx = R.m[0] * dx + R.m[3] * dy + R.m[6] * dz;
        const double ly = R.m[1] * dx + R.m[4] * dy + R.m[7] * dz;
        const double lz = R.m[2] * dx + R.m[5] * dy + R.m[8] * dz;
        cuboid_block_auto<true>(lx, ly, lz, 0.5 * d0, 0.5 * d1, 0.5 * d2, N);
        rotate_block(R, N, blk);
      } else {
        cuboid_block_auto<true>(dx, dy, dz, 0.5 * d0, 0.5 * d1, 0.5 * d2, N);
        blk[0] = N.xx; blk[1] = N.xy; blk[2] = N.xz;
        blk[3] = N.xy; blk[4] = N.yy; blk[5] = N.yz;
        blk[6] = N.xz; blk[7] = N.yz; blk[8] = N.zz;
      }
    } else {
#pragma unroll
      for (int q = 0; q < 9; ++q) blk[q] = 0.0;
    }
    // for each unit polarization k: row (k,i), contiguous over (j,l)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      __syncwarp();
#pragma unroll
      for (int l = 0; l < 3; ++l) s_stage[warp][0][3 * lane + l] = blk[3 * l + k] * inv_mu0;
      __syncwarp();
      double* row = p.out + ((int64_t)k * p.n + i) * p.n * 3 + 3 * j_warp0;
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int c = q * 32 + lane;
        if (c < ncols_warp) row[c] = s_stage[warp][0][c];
      }
    }
  }
}

template <bool ROTATED, bool FUSE_Q>
void launch_cell_major(const AsmParams& p, dim3 grid, cudaStream_t s) {
  if (p.max_dist != 0.0)
    assemble_cell_major_kernel<ROTATED, FUSE_Q, true><<<grid, SRC_TILE, 0, s>>>(p);
  else
    assemble_cell_major_kernel<ROTATED, FUSE_Q, false><<<grid, SRC_TILE, 0, s>>>(p);
}

__global__ void any_rotated_kernel(const double* __restrict__ quat, int64_t n, int* flag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const double x = quat[4 * i], y = quat[4 * i + 1], z = quat[4 * i + 2];
    if (x != 0.0 || y != 0.0 || z != 0.0) atomicOr(flag, 1);
  }
}

}  // namespace

// Returns 1 in *rotated_out if any cell has a non-identity orientation (host sync).
int mmr_detect_rotated(mmr_ctx* ctx, int64_t n, const double* d_quat, int* rotated_out) {
  MMR_CUDA(cudaMemsetAsync(ctx->ws, 0, sizeof(int), ctx->stream));
  any_rotated_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, ctx->stream>>>(d_quat, n,
                                                                              (int*)ctx->ws);
  MMR_CHECK_LAUNCH(ctx);
  MMR_CUDA(cudaMemcpyAsync(ctx->hpin, ctx->ws, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  MMR_CUDA(cudaStreamSynchronize(ctx->stream));
  *rotated_out = *(int*)ctx->hpin;
  return 0;
}

extern "C" int mmr_assemble(mmr_ctx* ctx, int64_t n, const double* d_pos, const double* d_quat,
                            const double* d_dim, int64_t src_begin, int64_t src_end,
                            int64_t tgt_first, int64_t n_tgt_local, int64_t tgt_block,
                            int64_t tgt_nparts, int64_t tgt_part, const double* d_chi,
                            double max_dist, int layout, double* d_out, int64_t ld) {
  MMR_ARG(ctx != nullptr, "ctx is NULL");
  MMR_ARG(n >= 0 && src_begin >= 0 && src_begin <= src_end && src_end <= n, "bad source range");
  MMR_ARG(n_tgt_local >= 0 && tgt_first >= 0 && tgt_nparts >= 1 && tgt_part >= 0 &&
              tgt_part < tgt_nparts,
          "bad target set");
  if (n == 0 || n_tgt_local == 0 || src_begin == src_end) return 0;
  MMR_ARG(d_pos && d_quat && d_dim && d_out, "NULL geometry/output pointer");
  MMR_ARG(tgt_block >= 1, "tgt_block must be >= 1");
  {
    // last local target must map inside [0,n)
    const int64_t t = n_tgt_local - 1;
    const int64_t jl = tgt_first + ((t / tgt_block) * tgt_nparts + tgt_part) * tgt_block + t % tgt_block;
    MMR_ARG(jl < n, "target set exceeds n");
  }
  DeviceGuard g(ctx->device);
  int rc = mmr_ws_reserve(ctx, 256);
  if (rc) return rc;
  rc = mmr_hpin_reserve(ctx, 256);
  if (rc) return rc;
  int rotated = 0;
  rc = mmr_detect_rotated(ctx, n, d_quat, &rotated);
  if (rc) return rc;

  AsmParams p;
  p.n = n; p.pos = d_pos; p.quat = d_quat; p.dim = d_dim;
  p.src_begin = src_begin; p.src_end = src_end;
  p.tgt_first = tgt_first; p.n_tgt_local = n_tgt_local; p.tgt_block = tgt_block;
  p.tgt_nparts = tgt_nparts; p.tgt_part = tgt_part;
  p.chi = d_chi; p.max_dist = max_dist; p.out = d_out; p.ld = ld;

  const double npairs = (double)(src_end - src_begin) * (double)n_tgt_local;
  if (layout == MMR_LAYOUT_CELL_MAJOR) {
    MMR_ARG(ld >= 3 * n, "ld must be >= 3*n for the cell-major layout");
    const int pidx = mmr_prof_begin(ctx, MMR_PROF_ASSEMBLE, ctx->stream);
    dim3 grid((unsigned)ceil_div64(src_end - src_begin, SRC_TILE),
              (unsigned)ceil_div64(n_tgt_local, TGT_TILE));
    if (rotated) {
      if (d_chi) launch_cell_major<true, true>(p, grid, ctx->stream);
      else launch_cell_major<true, false>(p, grid, ctx->stream);
    } else {
      if (d_chi) launch_cell_major<false, true>(p, grid, ctx->stream);
      else launch_cell_major<false, false>(p, grid, ctx->stream);
    }
    mmr_prof_end(ctx, pidx, npairs, ctx->stream);
    MMR_CHECK_LAUNCH(ctx);
    return 0;
  }
  if (layout == MMR_LAYOUT_REFERENCE) {
    MMR_ARG(d_chi == nullptr, "reference layout does not take chi");
    MMR_ARG(tgt_first == 0 && n_tgt_local == n && tgt_nparts == 1,
            "reference layout needs the full target range");
    dim3 grid((unsigned)ceil_div64(n, SRC_TILE), (unsigned)ceil_div64(src_end - src_begin, TGT_TILE));
    if (rotated) {
      if (max_dist != 0.0) assemble_reference_kernel<true, true><<<grid, SRC_TILE, 0, ctx->stream>>>(p);
      else assemble_reference_kernel<true, false><<<grid, SRC_TILE, 0, ctx->stream>>>(p);
    } else {
      if (max_dist != 0.0) assemble_reference_kernel<false, true><<<grid, SRC_TILE, 0, ctx->stream>>>(p);
      else assemble_reference_kernel<false, false><<<grid, SRC_TILE, 0, ctx->stream>>>(p);
    }
    MMR_CHECK_LAUNCH(ctx);
    return 0;
  }
  mmr_set_error("invalid argument: unknown layout %d", layout);
  return -1;
}
