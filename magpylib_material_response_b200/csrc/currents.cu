// Field of current sources at observers: the right-hand-side term of apply_demag for collections that
// contain currents (reference: magpy.getB(currents_list, pos, sumup=True), demag.py:456-463) and
// Collection.getB/getH of such collections.  Closed forms of magpylib's two line-current classes:
//
//   current.Polyline  straight segments p1 -> p2 (Biot-Savart integrated along the segment):
//       B = mu0 I / (4 pi) * (|r1| + |r2|) / (|r1| |r2| (|r1| |r2| + r1.r2)) * (r1 x r2),
//       r1 = p1 - p, r2 = p2 - p;  zero-length segments and observers on the line contribute 0
//       (for r1.r2 < 0 the denominator is rewritten so that it does not cancel next to the wire).
//   current.Circle    loop of radius a in the local xy plane (complete elliptic integrals K, E
//       by the arithmetic-geometric mean; the radial bracket through the AGM sum, see ellip_agm):
//       Bz = mu0 I / (2 pi s) * (K + (a^2 - rho^2 - z^2) / d2 * E),
//       Brho = mu0 I z / (2 pi rho s) * (-K + (a^2 + rho^2 + z^2) / d2 * E),
//       s = sqrt((a + rho)^2 + z^2), d2 = (a - rho)^2 + z^2, m = 4 a rho / s^2;
//       on the axis Brho = 0, Bz = mu0 I a^2 / (2 (a^2 + z^2)^1.5); on the wire 0.
//
// One thread per observer, sources staged through shared memory; the work is O(m * n_src) with
// n_src tiny (a handful of loops/segments) next to the O(n^2) assembly, so this kernel is not a
// roofline subject -- it exists so that no part of the path runs on the CPU.
#include "common.cuh"
#include "../../include/mmr_b200.h"

namespace {

constexpr double MU0_4PI = MMR_MU0 / (4.0 * 3.14159265358979323846);  // scipy.constants.mu_0 = magpy.mu_0
constexpr double PI = 3.14159265358979323846;

// K(m), E(m) and T = sum_{n>=1} 2^(n-1) c_n^2 of the AGM sequence (a0, b0, c0) = (1, sqrt(1-m), sqrt(m)),
// 0 <= m < 1;  E = K (1 - m/2 - T).  c_{n+1} = c_n^2 / (4 a_{n+1}) instead of the cancelling (a_n - b_n)/2.
__device__ __forceinline__ void ellip_agm(double m, double& K, double& E, double& T) {
  const double kc = sqrt(1.0 - m);
  double a = 0.5 * (1.0 + kc), b = sqrt(kc), c = m / (2.0 * (1.0 + kc)), pw = 1.0;
  T = 0.0;
#pragma unroll 1
  for (int it = 0; it < 12; ++it) {
    T += pw * c * c;
    const double an = 0.5 * (a + b);
    b = sqrt(a * b);
    a = an;
    c = c * c / (4.0 * a);
    pw *= 2.0;
    if (pw * c * c <= 1e-18 * T || c == 0.0) break;
  }
  K = 0.5 * PI / a;
  E = K * (1.0 - 0.5 * m - T);
}

__device__ __forceinline__ void quat_rotate(const double* q, const double* v, bool inverse, double* out) {
  // v' = v + 2 w (u x v) + 2 u x (u x v), scalar-last q = (u, w); inverse: u -> -u
  const double s = inverse ? -1.0 : 1.0;
  const double ux = s * q[0], uy = s * q[1], uz = s * q[2], w = q[3];
  const double tx = 2.0 * (uy * v[2] - uz * v[1]);
  const double ty = 2.0 * (uz * v[0] - ux * v[2]);
  const double tz = 2.0 * (ux * v[1] - uy * v[0]);
  out[0] = v[0] + w * tx + (uy * tz - uz * ty);
  out[1] = v[1] + w * ty + (uz * tx - ux * tz);
  out[2] = v[2] + w * tz + (ux * ty - uy * tx);
}

__global__ void __launch_bounds__(128)
    current_field_kernel(int64_t m, const double* __restrict__ obs, int nseg, const double* __restrict__ seg,
                         int nloop, const double* __restrict__ loop, double scale, double* __restrict__ out) {
  extern __shared__ double sh[];  // nseg*7 + nloop*9 doubles
  for (int e = threadIdx.x; e < nseg * 7; e += blockDim.x) sh[e] = seg[e];
  for (int e = threadIdx.x; e < nloop * 9; e += blockDim.x) sh[nseg * 7 + e] = loop[e];
  __syncthreads();
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m) return;
  const double p[3] = {obs[3 * t], obs[3 * t + 1], obs[3 * t + 2]};
  double bx = 0.0, by = 0.0, bz = 0.0;  // in units of mu0/(4 pi)
  for (int s = 0; s < nseg; ++s) {
    const double* g = sh + 7 * s;
    const double r1[3] = {g[0] - p[0], g[1] - p[1], g[2] - p[2]};
    const double r2[3] = {g[3] - p[0], g[4] - p[1], g[5] - p[2]};
    const double cx = r1[1] * r2[2] - r1[2] * r2[1];
    const double cy = r1[2] * r2[0] - r1[0] * r2[2];
    const double cz = r1[0] * r2[1] - r1[1] * r2[0];
    const double n1 = sqrt(r1[0] * r1[0] + r1[1] * r1[1] + r1[2] * r1[2]);
    const double n2 = sqrt(r2[0] * r2[0] + r2[1] * r2[1] + r2[2] * r2[2]);
    const double n12 = n1 * n2;
    const double c2 = cx * cx + cy * cy + cz * cz;
    // on the line through the segment (including its end points) or zero length: no contribution
    if (c2 <= 1e-30 * n12 * n12 || n12 == 0.0) continue;
    const double dot = r1[0] * r2[0] + r1[1] * r2[1] + r1[2] * r2[2];
    // n12 + dot cancels for an observer between the end points close to the wire: there use
    // n12 + dot = |r1 x r2|^2 / (n12 - dot)
    const double f = g[6] * (n1 + n2) * ((dot >= 0.0) ? 1.0 / (n12 * (n12 + dot)) : (n12 - dot) / (n12 * c2));
    bx += f * cx;
    by += f * cy;
    bz += f * cz;
  }
  for (int s = 0; s < nloop; ++s) {
    const double* g = sh + 7 * nseg + 9 * s;
    const double a = 0.5 * fabs(g[7]);
    if (a == 0.0) continue;
    const double d[3] = {p[0] - g[0], p[1] - g[1], p[2] - g[2]};
    double l[3];
    quat_rotate(g + 3, d, true, l);
    const double rho2 = l[0] * l[0] + l[1] * l[1];
    const double rho = sqrt(rho2), z = l[2];
    double bl[3];
    if (rho <= 1e-15 * a) {
      const double q = a * a + z * z;
      bl[0] = bl[1] = 0.0;
      bl[2] = 2.0 * PI * a * a / (q * sqrt(q));
    } else {
      const double s2 = (a + rho) * (a + rho) + z * z;
      const double d2 = (a - rho) * (a - rho) + z * z;
      if (d2 <= 1e-30 * a * a) continue;  // on the wire
      const double sq = sqrt(s2);
      double K, E, T;
      const double mm = 4.0 * a * rho / s2;
      ellip_agm(mm, K, E, T);
      const double bzc = 2.0 / sq * (K + (a * a - rho2 - z * z) / d2 * E);
      // -K + (a^2+rho^2+z^2)/d2 * E cancels like m^2 towards the axis: there (m < 1/2) it is
      // evaluated as K (m^2/2 - (2-m) T) / (2 (1-m)); towards the wire the K/E form is the stable one
      const double brk = (mm < 0.5) ? K * (0.5 * mm * mm - (2.0 - mm) * T) / (2.0 * (1.0 - mm))
                                    : -K + (a * a + rho2 + z * z) / d2 * E;
      const double brc = 2.0 * z / (rho * sq) * brk;
      bl[0] = brc * l[0] / rho;
      bl[1] = brc * l[1] / rho;
      bl[2] = bzc;
    }
    double bg[3];
    quat_rotate(g + 3, bl, false, bg);
    bx += g[8] * bg[0];
    by += g[8] * bg[1];
    bz += g[8] * bg[2];
  }
  out[3 * t] = scale * bx;
  out[3 * t + 1] = scale * by;
  out[3 * t + 2] = scale * bz;
}

}  // namespace

extern "C" int mmr_field_currents(mmr_ctx* ctx, int64_t m, const double* d_obs, int64_t nseg, const double* d_seg,
                                  int64_t nloop, const double* d_loop, int field, double* d_out) {
  MMR_ARG(ctx != nullptr, "ctx is NULL");
  MMR_ARG(m >= 0 && nseg >= 0 && nloop >= 0, "negative size");
  MMR_ARG(field == MMR_FIELD_B || field == MMR_FIELD_H, "field must be MMR_FIELD_B or MMR_FIELD_H");
  MMR_ARG(7 * nseg + 9 * nloop <= 24000, "too many current elements for one call (<= ~3400 segments)");
  if (m == 0) return 0;
  MMR_ARG(d_obs && d_out && (d_seg || nseg == 0) && (d_loop || nloop == 0), "NULL pointer");
  DeviceGuard g(ctx->device);
  const size_t smem = sizeof(double) * (size_t)(7 * nseg + 9 * nloop);
  if (smem > 48 * 1024) {
    const int rc = mmr_func_smem(ctx, (const void*)current_field_kernel, smem);
    if (rc) return rc;
  }
  // B = mu0/(4 pi) * (...);  H = B / mu0 = (...) / (4 pi) outside the (infinitely thin) conductor
  const double scale = (field == MMR_FIELD_B) ? MU0_4PI : 1.0 / (4.0 * PI);
  current_field_kernel<<<(unsigned)ceil_div64(m, 128), 128, smem, ctx->stream>>>(m, d_obs, (int)nseg, d_seg, (int)nloop,
                                                                                 d_loop, scale, d_out);
  MMR_CHECK_LAUNCH(ctx);
  return 0;
}
