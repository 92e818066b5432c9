// Field of the (solved) cells at arbitrary observers: B (tesla) or H (A/m), summed over all
// cells -- what `collection.getB(grid)` / `getH` return in the reference's tests
// (tests/test_isotropic_anisotropic.py:23; magpylib call).  Same closed
// form as the assembly kernel (cuboid.cuh); one observer per thread, source geometry staged
// through shared memory, deterministic two-stage reduction over source chunks.
#include "cuboid.cuh"

namespace {

constexpr int OBS_TILE = 128;
constexpr int SRC_CHUNK = 64;

struct SrcRec {
  double px, py, pz, ha, hb, hc, jx, jy, jz;
  double R[9];
};

template <bool FIELD_H>
__global__ void __launch_bounds__(OBS_TILE)
    field_partial_kernel(int64_t n, const double* __restrict__ pos, const double* __restrict__ quat,
                         const double* __restrict__ dim, const double* __restrict__ pol, int64_t m,
                         const double* __restrict__ obs, int64_t src_per_block,
                         double* __restrict__ partial) {
  __shared__ SrcRec s_src[SRC_CHUNK];
  const int64_t o = (int64_t)blockIdx.x * OBS_TILE + threadIdx.x;
  const int64_t oc = o < m ? o : m - 1;
  const double ox = obs[3 * oc], oy = obs[3 * oc + 1], oz = obs[3 * oc + 2];
  const int64_t s_begin = (int64_t)blockIdx.y * src_per_block;
  const int64_t s_end = min(n, s_begin + src_per_block);
  double ax = 0.0, ay = 0.0, az = 0.0;

  for (int64_t s0 = s_begin; s0 < s_end; s0 += SRC_CHUNK) {
    const int cnt = (int)min((int64_t)SRC_CHUNK, s_end - s0);
    __syncthreads();
    if (threadIdx.x < cnt) {
      const int64_t i = s0 + threadIdx.x;
      SrcRec& r = s_src[threadIdx.x];
      r.px = pos[3 * i]; r.py = pos[3 * i + 1]; r.pz = pos[3 * i + 2];
      r.ha = 0.5 * dim[3 * i]; r.hb = 0.5 * dim[3 * i + 1]; r.hc = 0.5 * dim[3 * i + 2];
      r.jx = pol[3 * i]; r.jy = pol[3 * i + 1]; r.jz = pol[3 * i + 2];
      Rot3 R;
      quat_to_rot(quat + 4 * i, R);
#pragma unroll
      for (int q = 0; q < 9; ++q) r.R[q] = R.m[q];
    }
    __syncthreads();
    for (int ss = 0; ss < cnt; ++ss) {
      const SrcRec& r = s_src[ss];
      if (r.jx == 0.0 && r.jy == 0.0 && r.jz == 0.0) continue;  // zero polarization => 0
      const double dx = ox - r.px, dy = oy - r.py, dz = oz - r.pz;
      const double lx = r.R[0] * dx + r.R[3] * dy + r.R[6] * dz;
      const double ly = r.R[1] * dx + r.R[4] * dy + r.R[7] * dz;
      const double lz = r.R[2] * dx + r.R[5] * dy + r.R[8] * dz;
      SymBlock N;
      cuboid_block_auto<FIELD_H>(lx, ly, lz, r.ha, r.hb, r.hc, N);
      const double fx = N.xx * r.jx + N.xy * r.jy + N.xz * r.jz;
      const double fy = N.xy * r.jx + N.yy * r.jy + N.yz * r.jz;
      const double fz = N.xz * r.jx + N.yz * r.jy + N.zz * r.jz;
      ax += r.R[0] * fx + r.R[1] * fy + r.R[2] * fz;
      ay += r.R[3] * fx + r.R[4] * fy + r.R[5] * fz;
      az += r.R[6] * fx + r.R[7] * fy + r.R[8] * fz;
    }
  }
  if (o < m) {
    double* dst = partial + ((int64_t)blockIdx.y * m + o) * 3;
    dst[0] = ax; dst[1] = ay; dst[2] = az;
  }
}

__global__ void field_reduce_kernel(int64_t m3, int nparts, const double* __restrict__ partial,
                                    double scale, double* __restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m3) return;
  double acc = 0.0;
  for (int p = 0; p < nparts; ++p) acc += partial[(int64_t)p * m3 + e];
  out[e] = acc * scale;
}

}  // namespace

extern "C" int mmr_field(mmr_ctx* ctx, int64_t n, const double* d_pos, const double* d_quat,
                         const double* d_dim, const double* d_pol, int64_t m, const double* d_obs,
                         int field, double* d_out) {
  MMR_ARG(ctx != nullptr, "ctx is NULL");
  MMR_ARG(n >= 0 && m >= 0, "negative size");
  MMR_ARG(field == MMR_FIELD_B || field == MMR_FIELD_H, "field must be MMR_FIELD_B or MMR_FIELD_H");
  if (m == 0) return 0;
  MMR_ARG(d_obs && d_out, "NULL observer/output pointer");
  DeviceGuard g(ctx->device);
  if (n == 0) {
    MMR_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double) * 3 * m, ctx->stream));
    return 0;
  }
  MMR_ARG(d_pos && d_quat && d_dim && d_pol, "NULL geometry pointer");
  const int64_t obs_blocks = ceil_div64(m, OBS_TILE);
  // enough source partitions to fill the GPU a few times over
  int64_t nparts = ceil_div64((int64_t)ctx->sm_count * 8, obs_blocks);
  nparts = nparts < 1 ? 1 : nparts;
  int64_t src_per_block = ceil_div64(n, nparts);
  src_per_block = ceil_div64(src_per_block, SRC_CHUNK) * SRC_CHUNK;
  nparts = ceil_div64(n, src_per_block);
  int rc = mmr_ws_reserve(ctx, sizeof(double) * 3 * m * nparts);
  if (rc) return rc;
  dim3 grid((unsigned)obs_blocks, (unsigned)nparts);
  if (field == MMR_FIELD_H)
    field_partial_kernel<true><<<grid, OBS_TILE, 0, ctx->stream>>>(n, d_pos, d_quat, d_dim, d_pol, m,
                                                                  d_obs, src_per_block, (double*)ctx->ws);
  else
    field_partial_kernel<false><<<grid, OBS_TILE, 0, ctx->stream>>>(n, d_pos, d_quat, d_dim, d_pol, m,
                                                                   d_obs, src_per_block, (double*)ctx->ws);
  MMR_CHECK_LAUNCH(ctx);
  const double scale = (field == MMR_FIELD_H) ? 1.0 / MMR_MU0 : 1.0;
  field_reduce_kernel<<<(unsigned)ceil_div64(3 * m, 256), 256, 0, ctx->stream>>>(
      3 * m, (int)nparts, (const double*)ctx->ws, scale, d_out);
  MMR_CHECK_LAUNCH(ctx);
  return 0;
}
