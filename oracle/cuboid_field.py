"""CPU ORACLE (test infrastructure, NOT product code) -- cuboid magnet field.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product path
(``magpylib_material_response_b200``) never imports it.

What it restates
----------------
The reference's demag tensor is three calls to ``magpylib.getH``
(/root/reference/src/magpylib_material_response/demag.py:190, :216, :220).  magpylib
(dependency ``magpylib>=5.0``, /root/reference/pyproject.toml:33, no lock file) is NOT
vendored in /root/reference and is not installable in this image, so its published
closed form (``magpylib/_src/fields/field_BH_cuboid.py``: ``magnet_cuboid_Bfield`` and
``BHJM_magnet_cuboid``; frame handling of ``field_wrap_BH.getBH_level1``) is restated
here as a numpy ufunc chain of the same shape (same term order, same octant folding,
same surface/edge special cases) -- see SURVEY.md section 8a rows A2/A3.

Pinning status
--------------
* pinned (loosely) by the reference's own goldens: the three ANSYS FEM files of
  /root/reference/tests/test_isotropic_anisotropic.py:9-67 (atol 1.2e-3/1.2e-3/6.5e-3 T),
  copied as fixtures to tests/golden/ansys/ -- see tests/test_oracle_golden.py;
* pinned tightly by independent known answers: tests/golden/cuboid_known_answers.json,
  produced by tests/golden/make_known_answers.py with mpmath (a) from the closed form
  at 50 digits and (b) from direct numerical quadrature of the magnetic surface-charge
  integral, which shares no code or formula with this file.
* At the north-star tolerances (T rtol 1e-10, J rtol 1e-8) the reference itself pins
  nothing ("parity unpinned" by reference tests); the mpmath vectors are what pin it.
"""

from __future__ import annotations

import numpy as np

MU0 = 1.25663706127e-06  # scipy.constants.mu_0 (= magpy.mu_0, demag.py:451); cancels in T
RTOL_SURFACE = 1e-15  # magpylib BHJM_magnet_cuboid surface band

# sign matrices [pol k][field l] accumulated by the octant folding
_QS_FLIPX = np.array([[1, -1, -1], [-1, 1, 1], [-1, 1, 1]], dtype=float)
_QS_FLIPY = np.array([[1, -1, 1], [-1, 1, -1], [1, -1, 1]], dtype=float)
_QS_FLIPZ = np.array([[1, 1, -1], [1, 1, -1], [-1, -1, 1]], dtype=float)


def fold_and_terms(observers, dimensions):
    """Octant folding + the six transcendental combinations of the closed form.

    Returns (ff1x, ff1y, ff1z, ff2x, ff2y, ff2z, qsigns) for m local-frame observers.
    Follows magpylib ``magnet_cuboid_Bfield`` term by term.
    """
    obs = np.array(observers, dtype=float, copy=True)
    dims = np.asarray(dimensions, dtype=float)
    a, b, c = dims.T / 2
    x, y, z = obs.T

    # evaluate in one octant only (strict inequalities)
    maskx = x < 0
    masky = y > 0
    maskz = z > 0
    x = np.where(maskx, -x, x)
    y = np.where(masky, -y, y)
    z = np.where(maskz, -z, z)

    qsigns = np.ones((len(x), 3, 3))
    qsigns[maskx] = qsigns[maskx] * _QS_FLIPX
    qsigns[masky] = qsigns[masky] * _QS_FLIPY
    qsigns[maskz] = qsigns[maskz] * _QS_FLIPZ

    xma, xpa = x - a, x + a
    ymb, ypb = y - b, y + b
    zmc, zpc = z - c, z + c

    xma2, xpa2 = xma**2, xpa**2
    ymb2, ypb2 = ymb**2, ypb**2
    zmc2, zpc2 = zmc**2, zpc**2

    mmm = np.sqrt(xma2 + ymb2 + zmc2)
    pmp = np.sqrt(xpa2 + ymb2 + zpc2)
    pmm = np.sqrt(xpa2 + ymb2 + zmc2)
    mmp = np.sqrt(xma2 + ymb2 + zpc2)
    mpm = np.sqrt(xma2 + ypb2 + zmc2)
    ppp = np.sqrt(xpa2 + ypb2 + zpc2)
    ppm = np.sqrt(xpa2 + ypb2 + zmc2)
    mpp = np.sqrt(xma2 + ypb2 + zpc2)

    with np.errstate(divide="ignore", invalid="ignore"):
        ff2x = np.log((xma + mmm) * (xpa + ppm) * (xpa + pmp) * (xma + mpp)) - np.log(
            (xpa + pmm) * (xma + mpm) * (xma + mmp) * (xpa + ppp)
        )
        ff2y = np.log(
            (-ymb + mmm) * (-ypb + ppm) * (-ymb + pmp) * (-ypb + mpp)
        ) - np.log((-ymb + pmm) * (-ypb + mpm) * (ymb - mmp) * (ypb - ppp))
        ff2z = np.log(
            (-zmc + mmm) * (-zmc + ppm) * (-zpc + pmp) * (-zpc + mpp)
        ) - np.log((-zmc + pmm) * (zmc - mpm) * (-zpc + mmp) * (zpc - ppp))

    ff1x = (
        np.arctan2(ymb * zmc, xma * mmm)
        - np.arctan2(ymb * zmc, xpa * pmm)
        - np.arctan2(ypb * zmc, xma * mpm)
        + np.arctan2(ypb * zmc, xpa * ppm)
        - np.arctan2(ymb * zpc, xma * mmp)
        + np.arctan2(ymb * zpc, xpa * pmp)
        + np.arctan2(ypb * zpc, xma * mpp)
        - np.arctan2(ypb * zpc, xpa * ppp)
    )
    ff1y = (
        np.arctan2(xma * zmc, ymb * mmm)
        - np.arctan2(xpa * zmc, ymb * pmm)
        - np.arctan2(xma * zmc, ypb * mpm)
        + np.arctan2(xpa * zmc, ypb * ppm)
        - np.arctan2(xma * zpc, ymb * mmp)
        + np.arctan2(xpa * zpc, ymb * pmp)
        + np.arctan2(xma * zpc, ypb * mpp)
        - np.arctan2(xpa * zpc, ypb * ppp)
    )
    ff1z = (
        np.arctan2(xma * ymb, zmc * mmm)
        - np.arctan2(xpa * ymb, zmc * pmm)
        - np.arctan2(xma * ypb, zmc * mpm)
        + np.arctan2(xpa * ypb, zmc * ppm)
        - np.arctan2(xma * ymb, zpc * mmp)
        + np.arctan2(xpa * ymb, zpc * pmp)
        + np.arctan2(xma * ypb, zpc * mpp)
        - np.arctan2(xpa * ypb, zpc * ppp)
    )
    return ff1x, ff1y, ff1z, ff2x, ff2y, ff2z, qsigns


def magnet_cuboid_Bfield(observers, dimensions, polarizations):
    """B (tesla) of homogeneously polarised cuboids, local frame, general case only.

    observers, dimensions, polarizations: (m,3) arrays.  Restates magpylib
    ``magnet_cuboid_Bfield``: B_l = 1/(4 pi) sum_k J_k G[k][l] F[k][l].
    """
    pol = np.asarray(polarizations, dtype=float)
    pol_x, pol_y, pol_z = pol.T
    ff1x, ff1y, ff1z, ff2x, ff2y, ff2z, qsigns = fold_and_terms(observers, dimensions)

    bx_pol_x = pol_x * ff1x * qsigns[:, 0, 0]
    by_pol_x = pol_x * ff2z * qsigns[:, 0, 1]
    bz_pol_x = pol_x * ff2y * qsigns[:, 0, 2]
    bx_pol_y = pol_y * ff2z * qsigns[:, 1, 0]
    by_pol_y = pol_y * ff1y * qsigns[:, 1, 1]
    bz_pol_y = -pol_y * ff2x * qsigns[:, 1, 2]
    bx_pol_z = pol_z * ff2y * qsigns[:, 2, 0]
    by_pol_z = -pol_z * ff2x * qsigns[:, 2, 1]
    bz_pol_z = pol_z * ff1z * qsigns[:, 2, 2]

    bx_tot = bx_pol_x + bx_pol_y + bx_pol_z
    by_tot = by_pol_x + by_pol_y + by_pol_z
    bz_tot = bz_pol_x + bz_pol_y + bz_pol_z

    B = np.stack((bx_tot, by_tot, bz_tot), axis=1)
    B /= 4 * np.pi
    return B


def cuboid_masks(observers, dimension, polarization=None):
    """(mask_gen, mask_inside) of magpylib ``BHJM_magnet_cuboid``."""
    obs = np.asarray(observers, dtype=float)
    x, y, z = obs.T
    a, b, c = np.abs(np.asarray(dimension, dtype=float).T) / 2

    mask_dim_not_null = (a * b * c).astype(bool)
    if polarization is None:
        mask_pol_not_null = np.ones(len(x), dtype=bool)
    else:
        pol_x, pol_y, pol_z = np.asarray(polarization, dtype=float).T
        mask_pol_not_null = ~((pol_x == 0) & (pol_y == 0) & (pol_z == 0))

    x_dist = np.abs(x) - a
    y_dist = np.abs(y) - b
    z_dist = np.abs(z) - c
    mask_surf_x = np.abs(x_dist) < RTOL_SURFACE * a
    mask_surf_y = np.abs(y_dist) < RTOL_SURFACE * b
    mask_surf_z = np.abs(z_dist) < RTOL_SURFACE * c
    mask_inside_x = x_dist < RTOL_SURFACE * a
    mask_inside_y = y_dist < RTOL_SURFACE * b
    mask_inside_z = z_dist < RTOL_SURFACE * c
    mask_inside = mask_inside_x & mask_inside_y & mask_inside_z

    mask_xedge = mask_surf_y & mask_surf_z & mask_inside_x
    mask_yedge = mask_surf_x & mask_surf_z & mask_inside_y
    mask_zedge = mask_surf_x & mask_surf_y & mask_inside_z
    mask_not_edge = ~(mask_xedge | mask_yedge | mask_zedge)

    mask_gen = mask_pol_not_null & mask_dim_not_null & mask_not_edge
    return mask_gen, mask_inside


def BHJM_magnet_cuboid(field, observers, dimension, polarization):
    """B (T) or H (A/m) in the cuboid's local frame, with the special cases.

    Restates magpylib ``BHJM_magnet_cuboid``: zero polarization -> 0,
    zero dimension -> 0, on an edge -> B = 0, H = (B - J*[inside]) / mu0.
    """
    if field not in ("B", "H"):
        msg = "oracle supports field 'B' or 'H' only"
        raise ValueError(msg)
    obs = np.atleast_2d(np.asarray(observers, dtype=float))
    dim = np.atleast_2d(np.asarray(dimension, dtype=float))
    pol = np.atleast_2d(np.asarray(polarization, dtype=float))
    m = max(len(obs), len(dim), len(pol))
    obs = np.broadcast_to(obs, (m, 3))
    dim = np.broadcast_to(dim, (m, 3))
    pol = np.broadcast_to(pol, (m, 3))

    mask_gen, mask_inside = cuboid_masks(obs, dim, pol)
    out = np.zeros((m, 3))
    if mask_gen.any():
        out[mask_gen] = magnet_cuboid_Bfield(obs[mask_gen], dim[mask_gen], pol[mask_gen])
    if field == "B":
        return out
    out[mask_inside] -= pol[mask_inside]
    return out / MU0


def quat_to_matrix(quat):
    """Rotation matrices (m,3,3) from scalar-last quaternions (x,y,z,w) as scipy's
    ``Rotation.as_quat()`` returns them (demag.py:178)."""
    q = np.atleast_2d(np.asarray(quat, dtype=float))
    q = q / np.linalg.norm(q, axis=1, keepdims=True)
    x, y, z, w = q.T
    Rm = np.empty((len(q), 3, 3))
    Rm[:, 0, 0] = 1 - 2 * (y * y + z * z)
    Rm[:, 0, 1] = 2 * (x * y - z * w)
    Rm[:, 0, 2] = 2 * (x * z + y * w)
    Rm[:, 1, 0] = 2 * (x * y + z * w)
    Rm[:, 1, 1] = 1 - 2 * (x * x + z * z)
    Rm[:, 1, 2] = 2 * (y * z - x * w)
    Rm[:, 2, 0] = 2 * (x * z - y * w)
    Rm[:, 2, 1] = 2 * (y * z + x * w)
    Rm[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return Rm


def getBH_cuboid_pairs(field, observers, position, quat, dimension, polarization):
    """Pairwise field, global frame: row p = field of cuboid p at observer p.

    Restates magpylib's functional interface ``getH("Cuboid", observers=, position=,
    orientation=, dimension=, polarization=)`` used at demag.py:190-192, i.e.
    ``getBH_level1``: r_loc = R^T (obs - pos); field_loc = BHJM(...); field = R field_loc
    (SURVEY 8a row A2).  ``polarization`` is in the cuboid's LOCAL frame.
    """
    obs = np.atleast_2d(np.asarray(observers, dtype=float))
    pos = np.atleast_2d(np.asarray(position, dtype=float))
    Rm = quat_to_matrix(quat)
    m = max(len(obs), len(pos), len(Rm))
    r = np.broadcast_to(obs, (m, 3)) - np.broadcast_to(pos, (m, 3))
    Rm = np.broadcast_to(Rm, (m, 3, 3))
    r_loc = np.einsum("pji,pj->pi", Rm, r)  # R^T r
    f_loc = BHJM_magnet_cuboid(field, r_loc, dimension, polarization)
    return np.einsum("pij,pj->pi", Rm, f_loc)


def getBH_cuboid_sources(field, observers, position, quat, dimension, polarization):
    """Object-interface shape: (n_src, n_obs, 3), every source at every observer.

    Restates ``magpy.getH(src_list, pos0)`` at demag.py:216,220 (no sumup).
    """
    obs = np.atleast_2d(np.asarray(observers, dtype=float))
    pos = np.atleast_2d(np.asarray(position, dtype=float))
    quat = np.atleast_2d(np.asarray(quat, dtype=float))
    dim = np.atleast_2d(np.asarray(dimension, dtype=float))
    pol = np.atleast_2d(np.asarray(polarization, dtype=float))
    ns, no = len(pos), len(obs)
    out = getBH_cuboid_pairs(
        field,
        np.tile(obs, (ns, 1)),
        np.repeat(pos, no, axis=0),
        np.repeat(quat, no, axis=0),
        np.repeat(dim, no, axis=0),
        np.repeat(pol, no, axis=0),
    )
    return out.reshape(ns, no, 3)


def getBH_collection(field, observers, position, quat, dimension, polarization, chunk=256):
    """Summed field of all cells at the observers, (n_obs,3) -- what
    ``collection.getB(grid)`` returns in tests/test_isotropic_anisotropic.py:23."""
    obs = np.atleast_2d(np.asarray(observers, dtype=float))
    pos = np.atleast_2d(np.asarray(position, dtype=float))
    quat = np.atleast_2d(np.asarray(quat, dtype=float))
    dim = np.atleast_2d(np.asarray(dimension, dtype=float))
    pol = np.atleast_2d(np.asarray(polarization, dtype=float))
    out = np.zeros((len(obs), 3))
    for s in range(0, len(pos), chunk):
        e = s + chunk
        out += getBH_cuboid_sources(field, obs, pos[s:e], quat[s:e], dim[s:e], pol[s:e]).sum(
            axis=0
        )
    return out
