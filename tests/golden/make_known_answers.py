"""Generate tests/golden/cuboid_known_answers.json (run once, offline; result committed).

Two independent high-precision evaluations of the dimensionless demag block
``N[l][k] = mu0 * H_l`` of a cuboid with unit polarization along k, source centred at the
origin with axis-aligned edges ``dim``, observer at ``offset``:

(a) ``closed``: the cuboid closed form evaluated with mpmath at
    50 significant digits -- removes FP64 rounding from the formula the oracle and the
    CUDA kernel implement;
(b) ``quad``: direct 2-D numerical quadrature (mpmath, 30 digits) of the magnetic
    surface-charge integral  H(r) = 1/(4 pi mu0) * sum_faces  sigma * int (r-r')/|r-r'|^3 dS',
    sigma = J.n  -- shares NO formula with (a), the oracle or the kernel (and needs no
    "subtract J inside" step: it yields H directly, inside and outside).

The script asserts (a) == (b) to 1e-20 before writing.  Usage:
    python tests/golden/make_known_answers.py
"""

from __future__ import annotations

import json
import pathlib

import mpmath as mp

CASES = [
    # (dim, offset)
    ((1, 1, 1), (0, 0, 0)),
    ((1, 1, 1), (1, 0, 0)),
    ((1, 1, 1), (1, 1, 0)),
    ((1, 1, 1), (1, 1, 1)),
    ((1, 1, 1), (2, 1, 0)),
    ((1, 1, 1), (-3, 2, 5)),
    ((1, 1, 1), (10, 7, 3)),
    ((1, 2, 3), (0, 0, 0)),
    ((1, 2, 3), ("0.5", "1.0", "-2.0")),  # on an edge extension: octant folding
    ((1, 2, 3), ("0.25", "-0.5", "0.75")),  # strictly inside, off-centre
    (("0.3", "0.7", "0.2"), ("-0.4", "0.9", "0.15")),
    ((2, 1, "0.5"), ("1.5", "-0.25", "-0.75")),
]


def closed_form_block(dim, off):
    """N[l][k] from the closed form, any precision (mp.dps)."""
    a, b, c = (mp.mpf(d) / 2 for d in dim)
    x, y, z = (mp.mpf(o) for o in off)
    inside = abs(x) < a and abs(y) < b and abs(z) < c
    F = [[mp.mpf(1)] * 3 for _ in range(3)]

    def mul(Fm):
        for k in range(3):
            for l in range(3):
                F[k][l] *= Fm[k][l]

    if x < 0:
        x = -x
        mul([[1, -1, -1], [-1, 1, 1], [-1, 1, 1]])
    if y > 0:
        y = -y
        mul([[1, -1, 1], [-1, 1, -1], [1, -1, 1]])
    if z > 0:
        z = -z
        mul([[1, 1, -1], [1, 1, -1], [-1, -1, 1]])
    X = {-1: x - a, 1: x + a}
    Y = {-1: y - b, 1: y + b}
    Z = {-1: z - c, 1: z + c}
    r = {
        (s1, s2, s3): mp.sqrt(X[s1] ** 2 + Y[s2] ** 2 + Z[s3] ** 2)
        for s1 in (-1, 1)
        for s2 in (-1, 1)
        for s3 in (-1, 1)
    }
    num = [(-1, -1, -1), (1, 1, -1), (1, -1, 1), (-1, 1, 1)]
    den = [(1, -1, -1), (-1, 1, -1), (-1, -1, 1), (1, 1, 1)]
    ff2x = sum(mp.log(X[s[0]] + r[s]) for s in num) - sum(mp.log(X[s[0]] + r[s]) for s in den)
    ff2y = sum(mp.log(-Y[s[1]] + r[s]) for s in num) - sum(
        mp.log(-Y[s[1]] + r[s]) for s in den
    )
    ff2z = sum(mp.log(-Z[s[2]] + r[s]) for s in num) - sum(
        mp.log(-Z[s[2]] + r[s]) for s in den
    )
    ff1x = ff1y = ff1z = mp.mpf(0)
    for s3 in (-1, 1):
        for s2 in (-1, 1):
            for s1 in (-1, 1):
                sg = -s1 * s2 * s3
                rr = r[(s1, s2, s3)]
                ff1x += sg * mp.atan2(Y[s2] * Z[s3], X[s1] * rr)
                ff1y += sg * mp.atan2(X[s1] * Z[s3], Y[s2] * rr)
                ff1z += sg * mp.atan2(X[s1] * Y[s2], Z[s3] * rr)
    G = [[ff1x, ff2z, ff2y], [ff2z, ff1y, -ff2x], [ff2y, -ff2x, ff1z]]
    N = [[None] * 3 for _ in range(3)]
    for k in range(3):
        for l in range(3):
            v = G[k][l] * F[k][l] / (4 * mp.pi)
            if inside and k == l:
                v -= 1
            N[l][k] = v
    return N


def quadrature_block(dim, off):
    """N[l][k] from the surface-charge integral by 2-D quadrature."""
    h = [mp.mpf(d) / 2 for d in dim]
    p = [mp.mpf(o) for o in off]
    N = [[None] * 3 for _ in range(3)]
    for k in range(3):
        u, v = [ax for ax in range(3) if ax != k]
        for l in range(3):
            tot = mp.mpf(0)
            for sgn in (-1, 1):  # face at coordinate k = sgn*h[k], charge sgn

                def integrand(su, sv, sgn=sgn, u=u, v=v, l=l, k=k):
                    d = [None] * 3
                    d[k] = p[k] - sgn * h[k]
                    d[u] = p[u] - su
                    d[v] = p[v] - sv
                    r2 = d[0] ** 2 + d[1] ** 2 + d[2] ** 2
                    return d[l] / (r2 * mp.sqrt(r2))

                # split the range at the observer's projection: the integrand peaks there
                def pts(ax):
                    out = [-h[ax]]
                    if -h[ax] < p[ax] < h[ax]:
                        out.append(p[ax])
                    out.append(h[ax])
                    return out

                tot += sgn * mp.quad(integrand, pts(u), pts(v))
            N[l][k] = tot / (4 * mp.pi)
    return N


def main():
    out = []
    for dim, off in CASES:
        mp.mp.dps = 50
        Nc = closed_form_block(dim, off)
        mp.mp.dps = 30
        Nq = quadrature_block(dim, off)
        err = max(abs(Nc[l][k] - Nq[l][k]) for l in range(3) for k in range(3))
        print(dim, off, "closed vs quadrature max abs diff:", mp.nstr(err, 3))
        assert err < mp.mpf("1e-20"), err
        mp.mp.dps = 50
        out.append(
            {
                "dim": [float(mp.mpf(d)) for d in dim],
                "offset": [float(mp.mpf(o)) for o in off],
                "block_l_k": [[mp.nstr(Nc[l][k], 25) for k in range(3)] for l in range(3)],
                "quadrature_max_abs_diff": float(err),
            }
        )
    path = pathlib.Path(__file__).with_name("cuboid_known_answers.json")
    path.write_text(json.dumps({"convention": "block[l][k] = mu0*H_l for unit J_k", "cases": out}, indent=1))
    print("wrote", path)


if __name__ == "__main__":
    main()
