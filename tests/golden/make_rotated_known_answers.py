"""Generate tests/golden/cuboid_rotated_known_answers.json (run once, offline; result committed).

Known answers for a ROTATED source cell: the dimensionless demag block
``N[l][k] = mu0 * H_l`` (GLOBAL frame) of a cuboid with edges ``dim``, centred at the origin, orientation given
by the quaternion ``quat`` (scalar-last, not normalised: integer components make the rotation matrix rational,
i.e. exact in mpmath), for unit GLOBAL polarization e_k, observer at the GLOBAL ``offset``.

(a) ``closed``: R . N_local(R^T r) . R^T with the axis-aligned closed form at 50 digits
    (make_known_answers.closed_form_block) -- the frame handling the reference delegates to magpylib
    (demag.py:183 rot0.inv().apply(unit_pol), getH's R^T(p_j - p_i) and R.H_local);
(b) ``quad``: direct 2-D quadrature of the surface-charge integral over the six faces of the rotated cuboid,
    parametrised in the GLOBAL frame: sigma = e_k . n_face,  H(r) = 1/(4 pi mu0) sum_faces sigma int (r-r')/|r-r'|^3 dS'.
    It never forms a local-frame field and shares no formula with (a), the oracle or the CUDA kernel.

The script asserts (a) == (b) to 1e-20 before writing.  Usage:  python tests/golden/make_rotated_known_answers.py
"""

from __future__ import annotations

import json
import pathlib

import mpmath as mp
from make_known_answers import closed_form_block

CASES = [
    # (dim, quat (x, y, z, w) un-normalised, global offset)
    ((1, 1, 1), (0, 1, 0, 1), ("1.5", 0, "0.25")),            # 90 deg about y (axes permuted)
    ((1, 2, 3), (1, 2, 3, 4), ("2.0", "-1.5", "2.5")),          # generic rotation, outside
    ((1, 2, 3), (1, 2, 3, 4), ("0.2", "-0.3", "0.1")),          # generic rotation, inside the cell
    (("0.3", "0.7", "0.2"), (3, -1, 2, 5), ("-0.4", "0.9", "0.15")),
    ((1, 1, 1), (0, 1, 0, mp.sqrt(2) + 1), ("1.5", 0, 0)),      # 45 deg about y: the soft cube of BASELINE configs[1]
]


def rotation_matrix(q):
    x, y, z, w = (mp.mpf(v) for v in q)
    n = x * x + y * y + z * z + w * w
    return [
        [(w * w + x * x - y * y - z * z) / n, 2 * (x * y - w * z) / n, 2 * (x * z + w * y) / n],
        [2 * (x * y + w * z) / n, (w * w - x * x + y * y - z * z) / n, 2 * (y * z - w * x) / n],
        [2 * (x * z - w * y) / n, 2 * (y * z + w * x) / n, (w * w - x * x - y * y + z * z) / n],
    ]


def closed_rotated_block(dim, q, off):
    Rm = rotation_matrix(q)
    r = [mp.mpf(o) for o in off]
    r_loc = [sum(Rm[m][a] * r[m] for m in range(3)) for a in range(3)]  # R^T r
    Nl = closed_form_block(dim, r_loc)  # [l][k] local
    # N_glob = R N_loc R^T
    return [[sum(Rm[l][a] * Nl[a][b] * Rm[k][b] for a in range(3) for b in range(3)) for k in range(3)] for l in range(3)]


def quadrature_rotated_block(dim, q, off):
    Rm = rotation_matrix(q)
    h = [mp.mpf(d) / 2 for d in dim]
    r = [mp.mpf(o) for o in off]
    p_loc = [sum(Rm[m][a] * r[m] for m in range(3)) for a in range(3)]  # only to split the ranges at the peak
    N = [[mp.mpf(0)] * 3 for _ in range(3)]
    for a in range(3):  # face normal = +-(column a of R)
        u, v = [ax for ax in range(3) if ax != a]

        def pts(ax):
            out = [-h[ax]]
            if -h[ax] < p_loc[ax] < h[ax]:
                out.append(p_loc[ax])
            out.append(h[ax])
            return out

        for sgn in (-1, 1):
            for l in range(3):

                def integrand(su, sv, sgn=sgn, a=a, u=u, v=v, l=l):
                    # point on the face, GLOBAL frame
                    rp = [sgn * h[a] * Rm[m][a] + su * Rm[m][u] + sv * Rm[m][v] for m in range(3)]
                    d = [r[m] - rp[m] for m in range(3)]
                    r2 = d[0] ** 2 + d[1] ** 2 + d[2] ** 2
                    return d[l] / (r2 * mp.sqrt(r2))

                val = mp.quad(integrand, pts(u), pts(v))
                for k in range(3):
                    N[l][k] += sgn * Rm[k][a] * val / (4 * mp.pi)  # sigma = e_k . (sgn * R[:, a])
    return N


def main():
    out = []
    for dim, q, off in CASES:
        mp.mp.dps = 50
        Nc = closed_rotated_block(dim, q, off)
        mp.mp.dps = 30
        Nq = quadrature_rotated_block(dim, q, off)
        err = max(abs(Nc[l][k] - Nq[l][k]) for l in range(3) for k in range(3))
        print(dim, [mp.nstr(mp.mpf(v), 6) for v in q], off, "closed vs quadrature max abs diff:", mp.nstr(err, 3), flush=True)
        assert err < mp.mpf("1e-20"), err
        mp.mp.dps = 50
        qn = mp.sqrt(sum(mp.mpf(v) ** 2 for v in q))
        out.append({
            "dim": [float(mp.mpf(d)) for d in dim],
            "quat": [float(mp.mpf(v) / qn) for v in q],
            "offset": [float(mp.mpf(o)) for o in off],
            "block_l_k": [[mp.nstr(Nc[l][k], 25) for k in range(3)] for l in range(3)],
            "quadrature_max_abs_diff": float(err),
        })
    path = pathlib.Path(__file__).with_name("cuboid_rotated_known_answers.json")
    path.write_text(json.dumps({"convention": "block[l][k] = mu0*H_l (global) for unit global J_k; quat scalar-last",
                                "cases": out}, indent=1))
    print("wrote", path)


if __name__ == "__main__":
    main()
