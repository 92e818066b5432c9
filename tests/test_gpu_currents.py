"""GPU parity of the current-source term: the CUDA kernel behind
``mmr_field_currents`` against (a) Biot-Savart quadrature goldens, (b) the numpy oracle, and the
full ``apply_demag`` of collections with currents (demag.py:456-463) against the oracle solve with
the currents' field folded into H_ext (rhs = J0 + S.(H_ext + B_currents), the same sum)."""

from __future__ import annotations

import json

import numpy as np
import pytest
import torch
from scipy.spatial.transform import Rotation as R

from magpylib_material_response_b200 import (
    Circle,
    Collection,
    Cuboid,
    Polyline,
    _native,
    apply_demag,
    from_json,
    mesh_all,
    mesh_Cuboid,
    to_json,
)
from magpylib_material_response_b200.field import current_arrays
from oracle import cuboid_field as ocf
from oracle import current_field as ocur
from oracle import demag_ref as oref

pytestmark = pytest.mark.gpu


def dev(ctx, a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=float)).to(ctx.device)


def kernel_B(ctx, pts, seg=None, loop=None, field=_native.FIELD_B):
    out = ctx.empty(len(pts), 3)
    ctx.field_currents(dev(ctx, pts), None if seg is None or not len(seg) else dev(ctx, seg),
                       None if loop is None or not len(loop) else dev(ctx, loop), out, field)
    return out.cpu().numpy()


def test_current_kernel_matches_quadrature_goldens(ctx, golden_dir):
    d = json.loads((golden_dir / "current_known_answers.json").read_text())
    unit = ocur.MU0 / (4 * np.pi)
    for s in d["segments"]:
        seg = np.array([[*map(float, s["p1"]), *map(float, s["p2"]), 1.0]])
        got = kernel_B(ctx, np.array([list(map(float, s["obs"]))]), seg=seg)[0] / unit
        ref = np.array(list(map(float, s["value"])))
        assert np.abs(got - ref).max() <= 1e-11 * np.abs(ref).max(), (s, got, ref)
    for s in d["loops"]:
        a = float(s["radius"])
        loop = np.array([[0, 0, 0, 0, 0, 0, 1, 2 * a, 1.0]])
        got = kernel_B(ctx, np.array([list(map(float, s["obs"]))]), loop=loop)[0] / unit
        ref = np.array(list(map(float, s["value"])))
        assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max(), (s, got, ref)


@pytest.mark.parametrize("field", ["B", "H"])
def test_current_kernel_random_sources_vs_oracle(ctx, field):
    rng = np.random.default_rng(7)
    m = 4097
    pts = rng.uniform(-3e-3, 3e-3, (m, 3))
    verts = rng.uniform(-2e-3, 2e-3, (40, 3))
    seg = np.hstack([verts[:-1], verts[1:], rng.normal(0, 3, (39, 1))])
    loop = np.hstack([rng.uniform(-2e-3, 2e-3, (5, 3)), R.random(5, rng).as_quat(), rng.uniform(1e-3, 4e-3, (5, 1)),
                      rng.normal(0, 5, (5, 1))])
    ref = ocur.currents_B(pts, seg, loop)
    scale = 1.0
    if field == "H":
        ref = ref / ocur.MU0
    got = kernel_B(ctx, pts, seg, loop, _native.FIELD_B if field == "B" else _native.FIELD_H) * scale
    err = np.abs(got - ref).max(axis=1) / np.abs(ref).max(axis=1)
    assert err.max() <= 1e-10, err.max()
    # special points: on a vertex, on a segment, on a loop axis, on a loop wire
    sp = np.array([verts[3], 0.5 * (verts[7] + verts[8]), loop[0, :3] + R.from_quat(loop[0, 3:7]).apply([0, 0, 1e-3]),
                   loop[1, :3] + R.from_quat(loop[1, 3:7]).apply([0.5 * loop[1, 7], 0, 0])])
    got = kernel_B(ctx, sp, seg, loop)
    ref = ocur.currents_B(sp, seg, loop)
    assert np.all(np.isfinite(got))
    np.testing.assert_allclose(got[[0, 2]], ref[[0, 2]], rtol=1e-9, atol=1e-16)


def test_collection_getB_with_currents_and_json_roundtrip(ctx):
    cube = Cuboid(polarization=(0.1, -0.2, 0.3), dimension=(1e-3, 2e-3, 1.5e-3), position=(1e-3, 0, 0))
    cube.rotate_from_angax(25, (1, 1, 0))
    pline = Polyline(current=2.5, vertices=[(0, 0, 0), (1e-3, 0, 0), (1e-3, 1e-3, 0), (0, 1e-3, 0)])
    pline.move((1e-3, 0, 2e-3))
    pline.rotate_from_angax(30, "y")
    circle = Circle(current=1.5, diameter=5e-3)
    circle.rotate_from_angax(45, "x")
    coll = Collection(cube, pline, Collection(circle, style_label="sub"), style_label="root")
    coll.move((0, 1e-3, 0))
    probe = np.array([(1e-3, 2e-3, 3e-3), (-1e-3, 0, 5e-3), (0.2e-3, 0.1e-3, -0.3e-3)])
    seg, loop = current_arrays([pline, circle])
    assert seg.shape == (3, 7) and loop.shape == (1, 9)
    ref = ocur.currents_B(probe, seg, loop) + ocf.getBH_collection(
        "B", probe, np.array([cube.position]), np.array([cube.orientation.as_quat()]), np.array([cube.dimension]),
        np.array([cube.polarization]))
    got = coll.getB(probe)
    np.testing.assert_allclose(got, ref, rtol=1e-10, atol=1e-18)
    np.testing.assert_allclose(coll.getH(probe), ref / ocur.MU0, rtol=1e-10, atol=1e-12)  # outside the magnet
    # reference tests/test_serialization.py::_assert_field_equal
    back = from_json(to_json(coll))[0]
    np.testing.assert_allclose(back.getB(probe), got, rtol=1e-13, atol=1e-20)


@pytest.mark.parametrize("solver", ["lu", "gmres"])
def test_apply_demag_iron_core_in_current_loop_vs_oracle(ctx, solver):
    """docs/serialization.md:40-56: soft iron core (chi = 3999) inside a current loop, mesh_all +
    apply_demag; plus a rectangular Polyline coil and an anisotropic hard magnet."""
    cube = Cuboid(polarization=(0, 0, 0), dimension=(0.002, 0.002, 0.004))
    cube.susceptibility = 3999
    hard = Cuboid(polarization=(0, 0.2, 1.0), dimension=(0.001, 0.001, 0.001), position=(0.004, 0, 0))
    hard.susceptibility = (0.3, 0.2, 0.1)
    hard.rotate_from_angax(20, "y")
    loop = Circle(current=10, diameter=0.003)
    coil = Polyline(current=-4.0, vertices=[(-2e-3, -2e-3, 1e-3), (2e-3, -2e-3, 1e-3), (2e-3, 2e-3, 1e-3),
                                            (-2e-3, 2e-3, 1e-3), (-2e-3, -2e-3, 1e-3)])
    coll = mesh_all(Collection(cube, loop, hard, coil, style_label="setup"), target_elems=260)
    cells = [s for s in coll.sources_all if isinstance(s, Cuboid)]
    currents = [s for s in coll.sources_all if not isinstance(s, Cuboid)]
    n = len(cells)
    assert n >= 200 and len(currents) == 2
    pos = np.array([c.position for c in cells])
    quat = np.array([c.orientation.as_quat() for c in cells])
    dim = np.array([c.dimension for c in cells])
    pol = np.array([c.polarization for c in cells])
    chi = np.array([np.broadcast_to(c.susceptibility, 3) for c in cells], dtype=float)
    seg, lp = current_arrays(currents)
    Bcur = ocur.currents_B(pos, seg, lp)
    ref = oref.apply_demag_ref(pos, quat, dim, pol, chi, H_ext=Bcur)
    out = apply_demag(coll, solver=solver, tol=1e-13)
    got = np.array([s.polarization for s in out.sources_all if isinstance(s, Cuboid)])
    scale = np.abs(ref).max(axis=1, keepdims=True)
    assert np.max(np.abs(got - ref) / scale) <= 1e-8
    assert np.abs(ref).max() > 1e-3  # the currents really magnetise the core
    # currents untouched, input untouched
    assert [type(s).__name__ for s in out.sources_all if not isinstance(s, Cuboid)] == ["Circle", "Polyline"]
    assert all(np.all(c.polarization == 0) for c in cells[:5])
    # the field of the solved setup (cells + currents) at a probe
    probe = np.array([(0, 0, 4e-3), (3e-3, 1e-3, -1e-3)])
    Bref = ocf.getBH_collection("B", probe, pos, quat, dim, ref) + ocur.currents_B(probe, seg, lp)
    np.testing.assert_allclose(out.getB(probe), Bref, rtol=1e-7, atol=1e-14)


def test_apply_demag_currents_only_excitation_matches_h_ext_equivalent(ctx):
    """A current far away acts like the uniform H_ext it produces (sanity of sign and units, demag.py:470)."""
    cube = Cuboid(polarization=(0, 0, 0), dimension=(1e-3, 1e-3, 1e-3))
    cube.susceptibility = 50.0
    mesh = mesh_Cuboid(cube, 64)
    loop = Circle(current=1e4, diameter=2.0)  # 1 m radius: B at the centre = mu0 I / (2 a), uniform over 1 mm
    out = apply_demag(Collection(mesh, loop))
    mesh2 = mesh_Cuboid(cube, 64)
    mesh2.H_ext = (0, 0, ocur.MU0 * 1e4 / 2.0)
    out2 = apply_demag(mesh2)
    a = np.array([s.polarization for s in out.sources_all if isinstance(s, Cuboid)])
    b = np.array([s.polarization for s in out2.sources_all])
    np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-9 * np.abs(b).max())
