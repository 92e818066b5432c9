// Analytic field of a homogeneously polarised cuboid -- device functions shared by the
// assembly kernel (assemble.cu) and the field-evaluation kernel (field.cu).
//
// Follows magpylib's published closed form (magnet_cuboid_Bfield / BHJM_magnet_cuboid, the
// arithmetic underneath the reference's magpy.getH calls at demag.py:190,:216,:220) as stated
// in SURVEY.md 8a row A3: strict octant folding, the 1e-15 relative surface band, on-edge =>
// B = 0, inside => subtract J.  The 9 entries of a pair's 3x3 block share 8 sqrt, 6 log-type
// and 24 atan2 terms, so the block is produced in ONE evaluation (the reference needs three).
#pragma once

#include "common.cuh"

// Symmetric local-frame block: N[l][k] = mu0 * H_l for unit polarization k (dimensionless).
struct SymBlock {
  double xx, yy, zz, xy, xz, yz;
};

// sum_{corners} sigma * atan2(n_{s2 s3}, d_{s1} * r_{s1 s2 s3}) in the reference's order
// mmm, pmm, mpm, ppm, mmp, pmp, mpp, ppp with signs + - - + - + + -.
// (u-,u+) is the axis of the term, (v-,v+), (w-,w+) the two others in the reference's
// corner-index positions.
struct Corners {
  double mmm, pmm, mpm, ppm, mmp, pmp, mpp, ppp;
};

// Out-of-line so the near-field path (24 atan2 + 6 log) stays small: the first version inlined
// all of them and ncu showed 39% of the stall samples as instruction-cache misses.
static __device__ __noinline__ double mmr_atan2(double y, double x) { return atan2(y, x); }
static __device__ __noinline__ double mmr_log(double x) { return log(x); }

// General (non-special-case) evaluation.  (x,y,z): observer in the source's local frame,
// (a,b,c): half edge lengths.  `inside` is subtracted from the diagonal (H = (B - J)/mu0).
__device__ __forceinline__ void cuboid_block_general(double x, double y, double z, double a,
                                                     double b, double c, double inside,
                                                     SymBlock& N) {
  // octant folding, strict inequalities; sign of the off-diagonals is s_k * s_l
  const double sx = (x < 0.0) ? -1.0 : 1.0;
  const double sy = (y > 0.0) ? -1.0 : 1.0;
  const double sz = (z > 0.0) ? -1.0 : 1.0;
  x = (x < 0.0) ? -x : x;
  y = (y > 0.0) ? -y : y;
  z = (z > 0.0) ? -z : z;

  const double xma = x - a, xpa = x + a;
  const double ymb = y - b, ypb = y + b;
  const double zmc = z - c, zpc = z + c;
  const double xma2 = xma * xma, xpa2 = xpa * xpa;
  const double ymb2 = ymb * ymb, ypb2 = ypb * ypb;
  const double zmc2 = zmc * zmc, zpc2 = zpc * zpc;

  // r_{s1 s2 s3}; additions in the reference's order (x^2 + y^2) + z^2, no FMA contraction
  // so that r matches the oracle's sqrt inputs bit for bit.
  Corners r;
  r.mmm = sqrt(__dadd_rn(__dadd_rn(xma2, ymb2), zmc2));
  r.pmm = sqrt(__dadd_rn(__dadd_rn(xpa2, ymb2), zmc2));
  r.mpm = sqrt(__dadd_rn(__dadd_rn(xma2, ypb2), zmc2));
  r.ppm = sqrt(__dadd_rn(__dadd_rn(xpa2, ypb2), zmc2));
  r.mmp = sqrt(__dadd_rn(__dadd_rn(xma2, ymb2), zpc2));
  r.pmp = sqrt(__dadd_rn(__dadd_rn(xpa2, ymb2), zpc2));
  r.mpp = sqrt(__dadd_rn(__dadd_rn(xma2, ypb2), zpc2));
  r.ppp = sqrt(__dadd_rn(__dadd_rn(xpa2, ypb2), zpc2));

  // log terms: numerator corners mmm, ppm, pmp, mpp; denominator pmm, mpm, mmp, ppp
  const double ff2x =
      mmr_log((xma + r.mmm) * (xpa + r.ppm) * (xpa + r.pmp) * (xma + r.mpp)) -
      mmr_log((xpa + r.pmm) * (xma + r.mpm) * (xma + r.mmp) * (xpa + r.ppp));
  const double ff2y =
      mmr_log((r.mmm - ymb) * (r.ppm - ypb) * (r.pmp - ymb) * (r.mpp - ypb)) -
      mmr_log((r.pmm - ymb) * (r.mpm - ypb) * (ymb - r.mmp) * (ypb - r.ppp));
  const double ff2z =
      mmr_log((r.mmm - zmc) * (r.ppm - zmc) * (r.pmp - zpc) * (r.mpp - zpc)) -
      mmr_log((r.pmm - zmc) * (zmc - r.mpm) * (r.mmp - zpc) * (zpc - r.ppp));

  const double yz_mm = ymb * zmc, yz_pm = ypb * zmc, yz_mp = ymb * zpc, yz_pp = ypb * zpc;
  const double ff1x = mmr_atan2(yz_mm, xma * r.mmm) - mmr_atan2(yz_mm, xpa * r.pmm) -
                      mmr_atan2(yz_pm, xma * r.mpm) + mmr_atan2(yz_pm, xpa * r.ppm) -
                      mmr_atan2(yz_mp, xma * r.mmp) + mmr_atan2(yz_mp, xpa * r.pmp) +
                      mmr_atan2(yz_pp, xma * r.mpp) - mmr_atan2(yz_pp, xpa * r.ppp);

  const double xz_mm = xma * zmc, xz_pm = xpa * zmc, xz_mp = xma * zpc, xz_pp = xpa * zpc;
  const double ff1y = mmr_atan2(xz_mm, ymb * r.mmm) - mmr_atan2(xz_pm, ymb * r.pmm) -
                      mmr_atan2(xz_mm, ypb * r.mpm) + mmr_atan2(xz_pm, ypb * r.ppm) -
                      mmr_atan2(xz_mp, ymb * r.mmp) + mmr_atan2(xz_pp, ymb * r.pmp) +
                      mmr_atan2(xz_mp, ypb * r.mpp) - mmr_atan2(xz_pp, ypb * r.ppp);

  const double xy_mm = xma * ymb, xy_pm = xpa * ymb, xy_mp = xma * ypb, xy_pp = xpa * ypb;
  const double ff1z = mmr_atan2(xy_mm, zmc * r.mmm) - mmr_atan2(xy_pm, zmc * r.pmm) -
                      mmr_atan2(xy_mp, zmc * r.mpm) + mmr_atan2(xy_pp, zmc * r.ppm) -
                      mmr_atan2(xy_mm, zpc * r.mmp) + mmr_atan2(xy_pm, zpc * r.pmp) +
                      mmr_atan2(xy_mp, zpc * r.mpp) - mmr_atan2(xy_pp, zpc * r.ppp);

  const double inv4pi = 1.0 / (4.0 * MMR_PI);
  // G = [[ff1x, ff2z, ff2y], [ff2z, ff1y, -ff2x], [ff2y, -ff2x, ff1z]], F[k][l] = s_k s_l
  N.xx = ff1x * inv4pi - inside;
  N.yy = ff1y * inv4pi - inside;
  N.zz = ff1z * inv4pi - inside;
  N.xy = ff2z * (sx * sy) * inv4pi;
  N.xz = ff2y * (sx * sz) * inv4pi;
  N.yz = -ff2x * (sy * sz) * inv4pi;
}

// ---------------------------------------------------------------------------------------
// Far-field evaluation (used when |r|^2 > 9 (a^2+b^2+c^2)).
//
// The reference sums eight atan2 terms per diagonal entry:  ff1 = sum_c sigma_c atan2(n_c, d_c).
// Each term is the argument of w_c = d_c + i sigma_c n_c, so ff1 = arg(prod_c w_c) (mod 2 pi).
// For an observer farther than 3 circumscribed radii the diagonal entries obey
// |N_kk| <= (2/pi) (R/(r-R))^2 < 0.16, i.e. |ff1| = 4 pi |N_kk| < pi, so the principal argument
// of the product IS the reference's sum: ONE atan2 instead of eight.  Likewise each pair of
// logs becomes one log of a ratio.  Rounding differs from the term-by-term sum at the 1e-16
// absolute level -- the same size as the closed form's own FP64 noise and inside the
// parity tolerance |dT| <= 1e-10 |T| + 2e-15.  Returns false (caller falls back to the exact
// term-by-term path) if a factor vanishes or a product is not finite, where the reference's own
// atan2(0,0) / log(0) conventions decide the value.
// ---------------------------------------------------------------------------------------
struct Cplx {
  double re, im;
};
__device__ __forceinline__ Cplx cmul(const Cplx& p, const Cplx& q) {
  Cplx r;
  r.re = p.re * q.re - p.im * q.im;
  r.im = p.re * q.im + p.im * q.re;
  return r;
}
// product over the 8 corners of (d_c + i sigma_c n_c), sigma = + - - + - + + - for
// mmm pmm mpm ppm mmp pmp mpp ppp ; um/up multiply the radius, n_* are the numerators.
__device__ __forceinline__ Cplx corner_product(double um, double up, const Corners& r, double n_mm,
                                               double n_pm, double n_mp, double n_pp,
                                               bool first_axis) {
  // first_axis: corner order for ff1x (u = x: pairs differ in s1, numerator index (s2,s3));
  // for ff1y / ff1z the roles of the corner indices are permuted by the caller through r.
  (void)first_axis;
  const Cplx w0 = {um * r.mmm, n_mm}, w1 = {up * r.pmm, -n_mm};
  const Cplx w2 = {um * r.mpm, -n_pm}, w3 = {up * r.ppm, n_pm};
  const Cplx w4 = {um * r.mmp, -n_mp}, w5 = {up * r.pmp, n_mp};
  const Cplx w6 = {um * r.mpp, n_pp}, w7 = {up * r.ppp, -n_pp};
  return cmul(cmul(cmul(w0, w1), cmul(w2, w3)), cmul(cmul(w4, w5), cmul(w6, w7)));
}

__device__ __forceinline__ bool cuboid_block_far(double x, double y, double z, double a, double b,
                                                 double c, SymBlock& N) {
  const double sx = (x < 0.0) ? -1.0 : 1.0;
  const double sy = (y > 0.0) ? -1.0 : 1.0;
  const double sz = (z > 0.0) ? -1.0 : 1.0;
  x = (x < 0.0) ? -x : x;
  y = (y > 0.0) ? -y : y;
  z = (z > 0.0) ? -z : z;
  // exact power-of-two scaling to O(1) coordinates (keeps the 8-factor products in range for any
  // unit system; the tensor is scale invariant and a power of two changes no rounding)
  {
    const double big = fmax(x, fmax(-y, -z));
    const int e = (__double2hiint(big) >> 20) & 0x7ff;
    const double sc = __hiloint2double((2046 - e) << 20, 0);
    x *= sc; y *= sc; z *= sc; a *= sc; b *= sc; c *= sc;
  }
  const double xma = x - a, xpa = x + a;
  const double ymb = y - b, ypb = y + b;
  const double zmc = z - c, zpc = z + c;
  const double xma2 = xma * xma, xpa2 = xpa * xpa;
  const double ymb2 = ymb * ymb, ypb2 = ypb * ypb;
  const double zmc2 = zmc * zmc, zpc2 = zpc * zpc;
  Corners r;
  r.mmm = sqrt(xma2 + ymb2 + zmc2);
  r.pmm = sqrt(xpa2 + ymb2 + zmc2);
  r.mpm = sqrt(xma2 + ypb2 + zmc2);
  r.ppm = sqrt(xpa2 + ypb2 + zmc2);
  r.mmp = sqrt(xma2 + ymb2 + zpc2);
  r.pmp = sqrt(xpa2 + ymb2 + zpc2);
  r.mpp = sqrt(xma2 + ypb2 + zpc2);
  r.ppp = sqrt(xpa2 + ypb2 + zpc2);

  // log terms as one log of a ratio each
  const double nx = (xma + r.mmm) * (xpa + r.ppm) * (xpa + r.pmp) * (xma + r.mpp);
  const double dx = (xpa + r.pmm) * (xma + r.mpm) * (xma + r.mmp) * (xpa + r.ppp);
  const double ny = (r.mmm - ymb) * (r.ppm - ypb) * (r.pmp - ymb) * (r.mpp - ypb);
  const double dy = (r.pmm - ymb) * (r.mpm - ypb) * (r.mmp - ymb) * (r.ppp - ypb);
  const double nz = (r.mmm - zmc) * (r.ppm - zmc) * (r.pmp - zpc) * (r.mpp - zpc);
  const double dz = (r.pmm - zmc) * (r.mpm - zmc) * (r.mmp - zpc) * (r.ppp - zpc);
  const double qx = nx / dx, qy = ny / dy, qz = nz / dz;

  // atan terms as the argument of the 8-corner complex product
  // ff1x: d = x_{s1} r, n = y_{s2} z_{s3}
  const Cplx px = corner_product(xma, xpa, r, ymb * zmc, ypb * zmc, ymb * zpc, ypb * zpc, true);
  // ff1y: d = y_{s2} r, n = x_{s1} z_{s3}; reference order/signs: + mmm - pmm - mpm + ppm - mmp + pmp + mpp - ppp
  Cplx py, pz;
  {
    const double n_mm = xma * zmc, n_pm = xpa * zmc, n_mp = xma * zpc, n_pp = xpa * zpc;
    const Cplx w0 = {ymb * r.mmm, n_mm}, w1 = {ymb * r.pmm, -n_pm};
    const Cplx w2 = {ypb * r.mpm, -n_mm}, w3 = {ypb * r.ppm, n_pm};
    const Cplx w4 = {ymb * r.mmp, -n_mp}, w5 = {ymb * r.pmp, n_pp};
    const Cplx w6 = {ypb * r.mpp, n_mp}, w7 = {ypb * r.ppp, -n_pp};
    py = cmul(cmul(cmul(w0, w1), cmul(w2, w3)), cmul(cmul(w4, w5), cmul(w6, w7)));
  }
  {
    const double n_mm = xma * ymb, n_pm = xpa * ymb, n_mp = xma * ypb, n_pp = xpa * ypb;
    const Cplx w0 = {zmc * r.mmm, n_mm}, w1 = {zmc * r.pmm, -n_pm};
    const Cplx w2 = {zmc * r.mpm, -n_mp}, w3 = {zmc * r.ppm, n_pp};
    const Cplx w4 = {zpc * r.mmp, -n_mm}, w5 = {zpc * r.pmp, n_pm};
    const Cplx w6 = {zpc * r.mpp, n_mp}, w7 = {zpc * r.ppp, -n_pp};
    pz = cmul(cmul(cmul(w0, w1), cmul(w2, w3)), cmul(cmul(w4, w5), cmul(w6, w7)));
  }
  // a vanished factor / non-finite product: let the exact path decide
  const double chk = (qx * qy * qz) * ((px.re * px.re + px.im * px.im) * (py.re * py.re + py.im * py.im) *
                                       (pz.re * pz.re + pz.im * pz.im));
  if (!(fabs(chk) > 0.0) || !(fabs(chk) < 1.7e308)) return false;

  const double inv4pi = 1.0 / (4.0 * MMR_PI);
  N.xx = atan2(px.im, px.re) * inv4pi;
  N.yy = atan2(py.im, py.re) * inv4pi;
  N.zz = atan2(pz.im, pz.re) * inv4pi;
  N.xy = log(qz) * (sx * sy) * inv4pi;
  N.xz = log(qy) * (sx * sz) * inv4pi;
  N.yz = -log(qx) * (sy * sz) * inv4pi;
  return true;
}

// Full evaluation including the special cases of BHJM_magnet_cuboid.
// If WITH_INSIDE is false the "- J inside" term is left out (that is B, not mu0*H).
template <bool WITH_INSIDE>
__device__ __forceinline__ void cuboid_block(double x, double y, double z, double a, double b,
                                             double c, SymBlock& N) {
  const double RTOL_SURFACE = 1e-15;
  a = fabs(a);
  b = fabs(b);
  c = fabs(c);
  const double xd = fabs(x) - a, yd = fabs(y) - b, zd = fabs(z) - c;
  const bool surf_x = fabs(xd) < RTOL_SURFACE * a;
  const bool surf_y = fabs(yd) < RTOL_SURFACE * b;
  const bool surf_z = fabs(zd) < RTOL_SURFACE * c;
  const bool in_x = xd < RTOL_SURFACE * a;
  const bool in_y = yd < RTOL_SURFACE * b;
  const bool in_z = zd < RTOL_SURFACE * c;
  const bool inside = in_x && in_y && in_z;
  const bool edge = (surf_y && surf_z && in_x) || (surf_x && surf_z && in_y) ||
                    (surf_x && surf_y && in_z);
  const bool dim_null = (a * b * c) == 0.0;
  const double ins = (WITH_INSIDE && inside) ? 1.0 : 0.0;
  if (dim_null || edge) {
    N.xx = N.yy = N.zz = -ins;
    N.xy = N.xz = N.yz = 0.0;
    return;
  }
  cuboid_block_general(x, y, z, a, b, c, ins, N);
}

// Dispatch used by the hot kernels: far pairs take the merged evaluation, the rest (and any
// degenerate far pair) the reference-exact one.
template <bool WITH_INSIDE>
__device__ __forceinline__ void cuboid_block_auto(double x, double y, double z, double a, double b,
                                                  double c, SymBlock& N) {
  const double r2 = x * x + y * y + z * z;
  const double R2 = a * a + b * b + c * c;
  const bool far = (r2 > 9.0 * R2) && (a * b * c != 0.0);
  bool done = false;
  if (far) done = cuboid_block_far(x, y, z, fabs(a), fabs(b), fabs(c), N);
  if (!done) cuboid_block<WITH_INSIDE>(x, y, z, a, b, c, N);
}

// Rotation matrix from a scalar-last quaternion (x,y,z,w), normalised like scipy does.
struct Rot3 {
  double m[9];  // row-major R[r][c]
};

__device__ __forceinline__ void quat_to_rot(const double* __restrict__ q, Rot3& R) {
  double x = q[0], y = q[1], z = q[2], w = q[3];
  const double inv = 1.0 / sqrt(x * x + y * y + z * z + w * w);
  x *= inv;
  y *= inv;
  z *= inv;
  w *= inv;
  R.m[0] = 1 - 2 * (y * y + z * z);
  R.m[1] = 2 * (x * y - z * w);
  R.m[2] = 2 * (x * z + y * w);
  R.m[3] = 2 * (x * y + z * w);
  R.m[4] = 1 - 2 * (x * x + z * z);
  R.m[5] = 2 * (y * z - x * w);
  R.m[6] = 2 * (x * z - y * w);
  R.m[7] = 2 * (y * z + x * w);
  R.m[8] = 1 - 2 * (x * x + y * y);
}

// Global-frame 3x3 block G[l][k] = (R N_loc R^T)[l][k]  (H = R H_loc, J_loc = R^T e_k;
// getBH_level1 frame handling).  out is row-major [l*3+k].
__device__ __forceinline__ void rotate_block(const Rot3& R, const SymBlock& N, double* out) {
  // M = N_loc * R^T   (M[a][k] = sum_b N[a][b] R[k][b])
  double M[9];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double r0 = R.m[3 * k + 0], r1 = R.m[3 * k + 1], r2 = R.m[3 * k + 2];
    M[0 * 3 + k] = N.xx * r0 + N.xy * r1 + N.xz * r2;
    M[1 * 3 + k] = N.xy * r0 + N.yy * r1 + N.yz * r2;
    M[2 * 3 + k] = N.xz * r0 + N.yz * r1 + N.zz * r2;
  }
#pragma unroll
  for (int l = 0; l < 3; ++l)
#pragma unroll
    for (int k = 0; k < 3; ++k)
      out[l * 3 + k] = R.m[3 * l + 0] * M[0 * 3 + k] + R.m[3 * l + 1] * M[1 * 3 + k] +
                       R.m[3 * l + 2] * M[2 * 3 + k];
}
