"""Host-side front end and persistence (no GPU): mesh_all / slice_Cuboid mirror
/root/reference/tests/test_meshing.py, to_json / from_json mirror
/root/reference/tests/test_serialization.py (field equality is replaced by attribute equality here;
the GPU suite checks fields), and the current-source oracle is pinned by quadrature goldens."""

from __future__ import annotations

import json
import pathlib

import numpy as np
import pytest

from magpylib_material_response_b200 import (
    Circle,
    Collection,
    Cuboid,
    Polyline,
    Sensor,
    from_json,
    mesh_all,
    mesh_Cuboid,
    slice_Cuboid,
    to_json,
)
from magpylib_material_response_b200.meshing import get_volume
from magpylib_material_response_b200.serialization import _deserialize_recursive, _serialize_recursive
from oracle import current_field as ocur

GOLDEN = pathlib.Path(__file__).parent / "golden"


# ---- slice_Cuboid (reference tests/test_meshing.py:62-118) ---------------------------------------
def test_slice_cuboid_parts_and_errors():
    c = Cuboid(dimension=(1, 1, 1))
    cm = slice_Cuboid(c, shift=0.5, axis="x")
    np.testing.assert_allclose(cm[0].dimension, [0.5, 1.0, 1.0])
    np.testing.assert_allclose(cm[1].dimension, [0.5, 1.0, 1.0])
    np.testing.assert_allclose(cm[0].position, [-0.25, 0.0, 0.0])
    np.testing.assert_allclose(cm[1].position, [0.25, 0.0, 0.0])
    cm = slice_Cuboid(c, shift=0.2, axis="y")
    np.testing.assert_allclose(cm[0].dimension, [1.0, 0.2, 1.0])
    np.testing.assert_allclose(cm[1].dimension, [1.0, 0.8, 1.0])
    np.testing.assert_allclose(cm[0].position, [0.0, -0.4, 0.0])
    np.testing.assert_allclose(cm[1].position, [0.0, 0.1, 0.0])
    cm = slice_Cuboid(c, shift=0.1, axis="z")
    np.testing.assert_allclose(cm[0].dimension, [1.0, 1.0, 0.1])
    np.testing.assert_allclose(cm[1].dimension, [1.0, 1.0, 0.9])
    np.testing.assert_allclose(cm[0].position, [0.0, 0.0, -0.45])
    np.testing.assert_allclose(cm[1].position, [0.0, 0.0, 0.05])
    with pytest.raises(ValueError, match=r"Shift must be between 0 and 1 *."):
        slice_Cuboid(c, shift=0, axis="y")
    with pytest.raises(TypeError, match="Object to be sliced must be a Cuboid"):
        slice_Cuboid(Circle(current=1, diameter=1), shift=0.5)


def test_slice_cuboid_inherits_pose_style_and_susceptibility():
    c = Cuboid(polarization=(1, 0, 0), dimension=(0.001, 0.002, 0.003), position=[0.004, 0.005, 0.006])
    c.style.label = "Cuboid 1"
    c.style.opacity = 0
    c.rotate_from_angax(76, "z", anchor=0)
    c.susceptibility = 3999
    cm = slice_Cuboid(c, shift=0.2, axis="y", style_opacity=0.5)
    assert all(s.susceptibility == c.susceptibility for s in cm.sources_all)
    assert cm.style.label == c.style.label
    assert cm.style.opacity == 0.5
    np.testing.assert_allclose(cm.position, c.position)
    np.testing.assert_allclose(cm.orientation.as_matrix(), c.orientation.as_matrix())
    assert np.isclose(sum(get_volume(s) for s in cm.sources_all), get_volume(c))


# ---- mesh_all (reference tests/test_meshing.py mesh_all tests; meshing.py:375-454) -----------------
def test_mesh_all_apportions_by_volume_and_keeps_currents():
    big = Cuboid(polarization=(0, 0, 1), dimension=(2e-3, 2e-3, 2e-3))
    small = Cuboid(polarization=(0, 0, 1), dimension=(1e-3, 1e-3, 1e-3), position=(3e-3, 0, 0))
    small.susceptibility = 5.0
    loop = Circle(current=10, diameter=3e-3)
    sub = Collection(small, style_label="sub")
    coll = Collection(big, loop, sub, style_label="setup")
    out = mesh_all(coll, target_elems=900)
    assert out is not coll and out.style.label == "setup"
    assert len(coll.sources_all) == 3  # the input is untouched
    cells = [s for s in out.sources_all if isinstance(s, Cuboid)]
    # volumes 8:1 -> targets int(8/9*900) = 800 and int(1/9*900) = 100 cells, near-cubic grids
    vol = np.array([get_volume(s) for s in cells])
    is_big = np.abs(np.array([s.position for s in cells])).max(axis=1) < 1.01e-3
    n_big = int(is_big.sum())
    assert abs(n_big - 800) <= 100 and abs((len(cells) - n_big) - 100) <= 36
    assert np.allclose(vol[is_big] * n_big, 8e-9, rtol=1e-12, atol=0)
    assert sum(isinstance(s, Circle) for s in out.sources_all) == 1
    assert np.isclose(vol.sum(), 9e-9, rtol=1e-12, atol=0)
    # susceptibility travels with the cells, the sub-collection keeps its label
    sus = [getattr(s, "susceptibility", None) for s in cells]
    assert sus.count(5.0) == len(cells) - n_big
    assert any(c.style.label == "sub" for c in out.collections_all)


def test_mesh_all_single_object_min_elems_per_child_and_inplace():
    c = Cuboid(polarization=(0, 0, 1), dimension=(1e-3, 1e-3, 1e-3))
    out = mesh_all(c, target_elems=27)
    assert isinstance(out, Collection) and len(out.sources_all) == 27
    a = Cuboid(dimension=(1e-3, 1e-3, 1e-3))
    b = Cuboid(dimension=(1e-4, 1e-4, 1e-4), position=(1e-3, 0, 0))
    coll = Collection(a, b)
    out = mesh_all(coll, target_elems=64, min_elems=8)
    sizes = sorted(len(ch.sources_all) for ch in out.children)
    assert sizes == [8, 64] or sizes[0] == 8  # the tiny cuboid is lifted to min_elems
    out2 = mesh_all(coll, target_elems=27, per_child_elems=True)
    assert sorted(len(ch.sources_all) for ch in out2.children) == [27, 27]
    res = mesh_all(coll, target_elems=16, inplace=True)
    assert res is coll and len(coll.sources_all) >= 16


def test_mesh_all_rejects_unsupported_objects():
    class Sphere(Cuboid.__mro__[1]):  # a magnet class the path does not mesh
        pass

    coll = Collection(Cuboid(dimension=(1, 1, 1)), Sphere(), Sphere())
    with pytest.raises(TypeError, match=r"Incompatible objects found: 2 Sphere\nSupported: \['Cuboid'\]"):
        mesh_all(coll, target_elems=10)
    with pytest.raises(TypeError, match="Object to be meshed must be a Cuboid"):
        mesh_Cuboid(Circle(current=1, diameter=1), 8)


# ---- JSON persistence (reference tests/test_serialization.py) -------------------------------------
def test_roundtrip_cuboid():
    src = Cuboid(polarization=(0.1, 0.2, 0.3), dimension=(0.001, 0.002, 0.003), position=(0.0, 0.001, 0.002))
    src.rotate_from_angax(30, "z")
    src.susceptibility = 0.5
    src.style.label = "cube"
    dd = _serialize_recursive(src)
    assert dd["type"] == "magnet.Cuboid"
    assert dd["orientation"]["representation"] == "matrix"
    assert dd["susceptibility"]["value"] == 0.5
    assert dd["polarization"] == {"value": [0.1, 0.2, 0.3], "unit": "T"}
    json.dumps(dd)  # JSON primitives only
    out = _deserialize_recursive(dd)
    assert isinstance(out, Cuboid)
    assert out.susceptibility == 0.5
    assert out.style.label == "cube"
    np.testing.assert_allclose(out.position, src.position)
    np.testing.assert_allclose(out.dimension, src.dimension)
    np.testing.assert_allclose(out.polarization, src.polarization)
    np.testing.assert_allclose(out.orientation.as_matrix(), src.orientation.as_matrix(), atol=1e-15)


def test_roundtrip_currents_and_sensor():
    pl = Polyline(current=2.5, vertices=[(0, 0, 0), (0.001, 0, 0), (0.001, 0.001, 0), (0, 0.001, 0)])
    pl.move((0.001, 0, 0))
    out = _deserialize_recursive(_serialize_recursive(pl))
    assert isinstance(out, Polyline) and out.current == 2.5
    np.testing.assert_allclose(out.vertices, pl.vertices)
    np.testing.assert_allclose(out.position, pl.position)
    ci = Circle(current=1.5, diameter=0.005)
    ci.rotate_from_angax(45, "x")
    out = _deserialize_recursive(_serialize_recursive(ci))
    assert isinstance(out, Circle) and out.current == 1.5 and out.diameter == 0.005
    np.testing.assert_allclose(out.orientation.as_matrix(), ci.orientation.as_matrix(), atol=1e-15)
    se = Sensor(pixel=[(0, 0, 0), (0.001, 0, 0)], position=(0, 0, 0.001))
    out = _deserialize_recursive(_serialize_recursive(se))
    assert isinstance(out, Sensor)
    np.testing.assert_allclose(out.pixel, se.pixel)


def test_roundtrip_collection_mixed_and_json_string():
    cube = Cuboid(polarization=(0, 0, 1), dimension=(0.001, 0.001, 0.001))
    cube.susceptibility = (0.5, 0.4, 0.3)
    pline = Polyline(current=1.0, vertices=[(0, 0, 0), (0.001, 0, 0), (0.001, 0.001, 0)])
    circle = Circle(current=2.0, diameter=0.002)
    sub = Collection(circle, style_label="sub")
    coll = Collection(cube, pline, sub, style_label="root")
    s = to_json(coll, indent=2)
    decoded = json.loads(s)
    assert decoded[0]["type"] == "Collection"
    assert [c["type"] for c in decoded[0]["children"]] == ["magnet.Cuboid", "current.Polyline", "Collection"]
    out = from_json(s)[0]
    assert isinstance(out, Collection) and out.style.label == "root" and len(out.children) == 3
    assert isinstance(out.children[0], Cuboid) and out.children[0].susceptibility == (0.5, 0.4, 0.3)
    assert isinstance(out.children[1], Polyline)
    assert isinstance(out.children[2], Collection) and out.children[2].style.label == "sub"
    assert out.children[0].parent is out
    objs = from_json(to_json(cube.copy(), Sensor()))
    assert len(objs) == 2 and isinstance(objs[0], Cuboid) and isinstance(objs[1], Sensor)


def test_serialization_errors_and_warnings():
    class Sphere(Cuboid.__mro__[1]):
        pass

    with pytest.raises(TypeError, match="Unsupported"):
        _serialize_recursive(Sphere())
    bad = {"type": "magnet.Bogus", "position": {"value": [0, 0, 0], "unit": "m"},
           "orientation": {"value": np.eye(3).tolist(), "representation": "matrix"}}
    with pytest.raises(TypeError, match="Unknown type tag"):
        _deserialize_recursive(bad)
    src = Cuboid(polarization=(0, 0, 1), dimension=(0.001, 0.001, 0.001))
    dd = _serialize_recursive(src)
    dd["position"]["unit"] = "mm"
    with pytest.raises(ValueError, match="Position unit"):
        _deserialize_recursive(dd)
    dd = _serialize_recursive(src)
    dd["orientation"]["representation"] = "quaternion"
    with pytest.raises(ValueError, match="Orientation representation"):
        _deserialize_recursive(dd)
    Collection(src)  # sets src.parent
    with pytest.warns(UserWarning, match="parent"):
        _serialize_recursive(src)


def test_solved_collection_survives_json(tmp_path):
    """docs/serialization.md: a meshed (and solved) collection is stored and reloaded cell by cell."""
    cube = Cuboid(polarization=(0, 0, 0), dimension=(0.002, 0.002, 0.004))
    cube.susceptibility = 3999
    coll = mesh_all(Collection(cube, Circle(current=10, diameter=0.003), style_label="setup"), target_elems=60)
    rng = np.random.default_rng(0)
    for s in coll.sources_all:
        if isinstance(s, Cuboid):
            s.polarization = rng.normal(size=3)  # stands in for the solved polarizations
    path = tmp_path / "demag.json"
    path.write_text(to_json(coll))
    back = from_json(path.read_text())[0]
    a = [s for s in coll.sources_all if isinstance(s, Cuboid)]
    b = [s for s in back.sources_all if isinstance(s, Cuboid)]
    assert len(a) == len(b)
    for x, y in zip(a, b, strict=True):
        assert np.array_equal(x.polarization, y.polarization)  # repr round trip of doubles is exact
        assert np.array_equal(x.position, y.position)
        assert y.susceptibility == 3999


# ---- current-source oracle pinned by Biot-Savart quadrature ---------------------------------------
def test_current_oracle_matches_quadrature_goldens():
    d = json.loads((GOLDEN / "current_known_answers.json").read_text())
    unit = ocur.MU0 / (4 * np.pi)
    for s in d["segments"]:
        seg = np.array([[*map(float, s["p1"]), *map(float, s["p2"]), 1.0]])
        got = ocur.segments_B([list(map(float, s["obs"]))], seg)[0] / unit
        ref = np.array(list(map(float, s["value"])))
        assert np.abs(got - ref).max() <= 1e-11 * np.abs(ref).max(), (s, got, ref)
    for s in d["loops"]:
        a = float(s["radius"])
        loop = np.array([[0, 0, 0, 0, 0, 0, 1, 2 * a, 1.0]])
        got = ocur.loops_B([list(map(float, s["obs"]))], loop)[0] / unit
        ref = np.array(list(map(float, s["value"])))
        assert np.abs(got - ref).max() <= 1e-12 * np.abs(ref).max(), (s, got, ref)


def test_current_oracle_agm_matches_scipy_elliptic_integrals():
    from scipy.special import ellipe, ellipk

    m = np.concatenate([np.logspace(-12, -1, 23), np.linspace(0.1, 0.999999, 50)])
    K, E, T = ocur.ellip_agm(m)
    np.testing.assert_allclose(K, ellipk(m), rtol=4e-16)
    np.testing.assert_allclose(E, ellipe(m), rtol=2e-15)
    assert np.all(T >= 0)


def test_current_oracle_conventions():
    # long straight wire: B = mu0 I / (2 pi r) around the wire, right-hand rule
    seg = np.array([[0, 0, -1e3, 0, 0, 1e3, 2.0]])
    B = ocur.segments_B([[0.5, 0, 0]], seg)[0]
    np.testing.assert_allclose(B, [0, ocur.MU0 * 2.0 / (2 * np.pi * 0.5), 0], rtol=1e-6, atol=1e-20)
    # on the wire / on its extension / zero length: no contribution
    assert np.all(ocur.segments_B([[0, 0, 0.3], [0, 0, 2e3]], seg) == 0)
    assert np.all(ocur.segments_B([[1, 2, 3]], np.array([[1, 1, 1, 1, 1, 1, 5.0]])) == 0)
    # loop centre: B = mu0 I / (2 a); rotated loop turns the axis; point on the wire -> 0
    loop = np.array([[0, 0, 0, 0, 0, 0, 1, 2.0, 3.0]])
    np.testing.assert_allclose(ocur.loops_B([[0, 0, 0]], loop)[0], [0, 0, ocur.MU0 * 3.0 / 2], rtol=1e-15)
    from scipy.spatial.transform import Rotation as R

    q = R.from_euler("y", 90, degrees=True).as_quat()
    loop_r = np.array([[0, 0, 0, *q, 2.0, 3.0]])
    np.testing.assert_allclose(ocur.loops_B([[0, 0, 0]], loop_r)[0], [ocur.MU0 * 3.0 / 2, 0, 0], atol=1e-21)
    assert np.all(ocur.loops_B([[1.0, 0, 0]], loop) == 0)
