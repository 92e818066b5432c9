"""CPU ORACLE (test infrastructure, NOT product code) -- fields of line-current sources.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module.

What it restates
----------------
``apply_demag`` adds ``S . magpy.getB(currents_list, pos, sumup=True)`` to the right-hand side
(/root/reference/src/magpylib_material_response/demag.py:456-463).  The arithmetic lives in the
absent third-party magpylib (``magpylib/_src/fields/field_BH_polyline.py`` :
``current_polyline_Hfield`` / ``current_vertices_field``; ``field_BH_circle.py`` :
``current_circle_Hfield``), so the published closed forms are restated here:

* straight segment p1 -> p2, current I (Biot-Savart integrated in closed form):
      B = mu0 I/(4 pi) (|r1|+|r2|) / (|r1||r2| (|r1||r2| + r1.r2)) (r1 x r2),  r_i = p_i - p
  (for r1.r2 < 0 the denominator is rewritten with |r1||r2| + r1.r2 = |r1 x r2|^2 / (|r1||r2| - r1.r2),
  which does not cancel next to the wire);
  zero-length segments and observers on the line contribute 0 (magpylib convention);
* circular loop of radius a in the local xy plane (Smythe / Ortner et al.):
      Bz   = mu0 I/(2 pi s) (K(m) + (a^2 - rho^2 - z^2)/d2 E(m)),
      Brho = mu0 I z/(2 pi rho s) (-K(m) + (a^2 + rho^2 + z^2)/d2 E(m)),
      s^2 = (a+rho)^2 + z^2, d2 = (a-rho)^2 + z^2, m = 4 a rho / s^2,
  observers on the wire get 0; frame handling as magpylib's getBH_level1 (observer into the local
  frame, field back).  K and E come from the arithmetic-geometric mean (checked against
  scipy.special in the tests); the radial bracket, which cancels like m^2 towards the axis, is
  evaluated for m < 1/2 through the AGM sum T = sum_{n>=1} 2^(n-1) c_n^2:
  -K + cE = K (m^2/2 - (2-m) T)/(2(1-m))  (magpylib reaches the same stability through Bulirsch's cel).

Pinning status: the reference's tests hold no vector for this term ("parity unpinned" by reference
tests).  Pinned by tests/golden/current_known_answers.json: direct numerical quadrature of the
Biot-Savart line integral (mpmath, 30 digits; tests/golden/make_current_known_answers.py), which
shares no formula with this file or the kernel.
"""

from __future__ import annotations

import numpy as np
from scipy.spatial.transform import Rotation as R

MU0 = 1.25663706127e-06  # scipy.constants.mu_0 = magpy.mu_0 (as in oracle/cuboid_field.py)


def segments_B(points, seg):
    """points (m,3); seg (nseg,7) = (p1, p2, I) in the global frame -> B (m,3) tesla."""
    pts = np.asarray(points, float).reshape(-1, 3)
    out = np.zeros_like(pts)
    for s in np.asarray(seg, float).reshape(-1, 7):
        r1 = s[0:3] - pts
        r2 = s[3:6] - pts
        cr = np.cross(r1, r2)
        n1 = np.linalg.norm(r1, axis=1)
        n2 = np.linalg.norm(r2, axis=1)
        n12 = n1 * n2
        c2 = np.einsum("ij,ij->i", cr, cr)
        ok = (c2 > 1e-30 * n12 * n12) & (n12 > 0)
        dot = np.einsum("ij,ij->i", r1, r2)
        # |r1||r2| + r1.r2 cancels when the observer sits between the end points close to the wire;
        # there use the identity n12 + dot = |r1 x r2|^2 / (n12 - dot)
        with np.errstate(divide="ignore", invalid="ignore"):
            f = np.where(dot >= 0, (n1 + n2) / (n12 * (n12 + dot)), (n1 + n2) * (n12 - dot) / (n12 * c2))
            f = s[6] * f
        out[ok] += (MU0 / (4 * np.pi)) * f[ok, None] * cr[ok]
    return out


def ellip_agm(m):
    """K(m), E(m) and T = sum_{n>=1} 2^(n-1) c_n^2 of the AGM sequence (a0, b0, c0) = (1, sqrt(1-m), sqrt(m));
    E = K (1 - m/2 - T).  c_{n+1} = c_n^2 / (4 a_{n+1}) avoids the cancelling difference (a_n - b_n)/2."""
    m = np.asarray(m, float)
    kc = np.sqrt(1.0 - m)
    a = 0.5 * (1.0 + kc)
    b = np.sqrt(kc)
    c = m / (2.0 * (1.0 + kc))
    T = np.zeros_like(m)
    pw = 1.0
    for _ in range(12):
        T = T + pw * c * c
        an = 0.5 * (a + b)
        b = np.sqrt(a * b)
        a = an
        c = c * c / (4.0 * a)
        pw *= 2.0
    K = 0.5 * np.pi / a
    return K, K * (1.0 - 0.5 * m - T), T


def loops_B(points, loop):
    """points (m,3); loop (nloop,9) = (centre, quat scalar-last, diameter, I) -> B (m,3) tesla."""
    pts = np.asarray(points, float).reshape(-1, 3)
    out = np.zeros_like(pts)
    for g in np.asarray(loop, float).reshape(-1, 9):
        a = 0.5 * abs(g[7])
        if a == 0:
            continue
        rot = R.from_quat(g[3:7])
        loc = rot.apply(pts - g[0:3], inverse=True)
        rho = np.hypot(loc[:, 0], loc[:, 1])
        z = loc[:, 2]
        bl = np.zeros_like(pts)
        axis = rho <= 1e-15 * a
        q = a * a + z * z
        bl[axis, 2] = 2 * np.pi * a * a / (q[axis] * np.sqrt(q[axis]))
        s2 = (a + rho) ** 2 + z * z
        d2 = (a - rho) ** 2 + z * z
        ok = (~axis) & (d2 > 1e-30 * a * a)
        m = 4 * a * rho[ok] / s2[ok]
        K, E, T = ellip_agm(m)
        sq = np.sqrt(s2[ok])
        bz = 2.0 / sq * (K + (a * a - rho[ok] ** 2 - z[ok] ** 2) / d2[ok] * E)
        # radial bracket: AGM-sum form towards the axis (m < 1/2), the K/E form towards the wire
        brk = np.where(m < 0.5, K * (0.5 * m * m - (2.0 - m) * T) / (2.0 * (1.0 - m)),
                       -K + (a * a + rho[ok] ** 2 + z[ok] ** 2) / d2[ok] * E)
        br = 2.0 * z[ok] / (rho[ok] * sq) * brk
        bl[ok, 0] = br * loc[ok, 0] / rho[ok]
        bl[ok, 1] = br * loc[ok, 1] / rho[ok]
        bl[ok, 2] = bz
        out += (MU0 / (4 * np.pi)) * g[8] * rot.apply(bl)
    return out


def currents_B(points, seg=None, loop=None):
    pts = np.asarray(points, float).reshape(-1, 3)
    out = np.zeros_like(pts)
    if seg is not None and len(seg):
        out += segments_B(pts, seg)
    if loop is not None and len(loop):
        out += loops_B(pts, loop)
    return out
