"""Generate tests/golden/current_known_answers.json (run once, offline; result committed).

Independent known answers for the current-source fields: direct numerical
quadrature (mpmath, 30 digits) of the Biot-Savart line integral
    B(p) = mu0 I / (4 pi) * int dl x (p - l) / |p - l|^3
along straight segments and around circular loops -- shares no formula with the closed forms in
oracle/current_field.py or csrc/currents.cu.  Values are stored in units of mu0/(4 pi) * I
(i.e. the bare integral), so no physical constant enters.

Usage:  python tests/golden/make_current_known_answers.py
"""

from __future__ import annotations

import json
import pathlib

import mpmath as mp

mp.mp.dps = 30

SEGMENTS = [
    # (p1, p2, observer)
    ((0, 0, -1), (0, 0, 1), (1, 0, 0)),
    ((0, 0, 0), ("0.001", 0, 0), ("0.0004", "0.0007", "-0.0002")),
    (("-0.3", "0.2", "0.1"), ("0.5", "-0.4", "0.9"), ("0.25", "0.35", "-0.6")),
    ((1, 2, 3), (4, 6, 3), ("2.5", 4, "3.0001")),  # very close to the wire
    ((0, 0, 0), (1, 1, 1), (3, 3, "3.5")),  # near the extension of the segment
]
LOOPS = [
    # (radius, observer in the loop frame: loop in the xy plane, centred at the origin)
    (1, (0, 0, "0.5")),  # on the axis
    (1, ("0.3", "0.1", "0.2")),
    (1, ("1e-7", 0, "0.3")),  # almost on the axis
    ("0.0015", ("0.0009", "-0.0004", "0.0011")),
    (1, ("2.5", "-1.5", "0.75")),
    (1, ("0.999", 0, "0.002")),  # close to the wire
    (2, ("1.2", "0.9", 0)),  # in the loop plane, inside
]


def cross(a, b):
    return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


def seg_integral(p1, p2, p):
    p1, p2, p = ([mp.mpf(x) for x in v] for v in (p1, p2, p))
    d = [b - a for a, b in zip(p1, p2)]

    def integrand(comp):
        def f(t):
            l = [a + t * dd for a, dd in zip(p1, d)]
            r = [pp - ll for pp, ll in zip(p, l)]
            n = mp.sqrt(sum(x * x for x in r))
            return cross(d, r)[comp] / n**3

        return f

    # break points cluster geometrically around the foot of the perpendicular, where the integrand peaks
    dd = sum(x * x for x in d)
    t0 = sum((pp - a) * x for pp, a, x in zip(p, p1, d)) / dd
    pts = {mp.mpf(0), mp.mpf(1)}
    for k in range(0, 9):
        for sgn in (-1, 1):
            t = t0 + sgn * mp.mpf(10) ** (-k)
            if 0 < t < 1:
                pts.add(t)
    if 0 < t0 < 1:
        pts.add(t0)
    pts = sorted(pts)
    return [mp.quad(integrand(c), pts) for c in range(3)]


def loop_integral(a, p):
    a = mp.mpf(a)
    p = [mp.mpf(x) for x in p]

    def integrand(comp):
        def f(phi):
            l = [a * mp.cos(phi), a * mp.sin(phi), mp.mpf(0)]
            dl = [-a * mp.sin(phi), a * mp.cos(phi), mp.mpf(0)]
            r = [pp - ll for pp, ll in zip(p, l)]
            n = mp.sqrt(sum(x * x for x in r))
            return cross(dl, r)[comp] / n**3

        return f

    # split at the azimuth of the observer, where the integrand peaks
    phi0 = mp.atan2(p[1], p[0])
    pts = [phi0 + k * mp.pi / 8 for k in range(17)]
    return [mp.quad(integrand(c), pts) for c in range(3)]


def main():
    out = {"unit": "integral of dl x r / |r|^3 (multiply by mu0 I / (4 pi) for tesla)", "segments": [], "loops": []}
    for p1, p2, p in SEGMENTS:
        val = seg_integral(p1, p2, p)
        out["segments"].append({"p1": [str(x) for x in p1], "p2": [str(x) for x in p2], "obs": [str(x) for x in p],
                                "value": [mp.nstr(v, 20) for v in val]})
    for a, p in LOOPS:
        val = loop_integral(a, p)
        out["loops"].append({"radius": str(a), "obs": [str(x) for x in p], "value": [mp.nstr(v, 20) for v in val]})
    path = pathlib.Path(__file__).with_name("current_known_answers.json")
    path.write_text(json.dumps(out, indent=1) + "\n")
    print("wrote", path)


if __name__ == "__main__":
    main()
