"""CPU: structure-of-arrays collections (second half).  ``mesh_Cuboid`` keeps its
cells in a CellBlock; the per-cell objects of the reference (meshing.py:90-93) are views made on demand,
and ``apply_demag``'s host glue (demag.py:380-381, :420-440, :472-476) reads / writes array slices.
Every test compares the array route with the per-object route of the reference."""

from __future__ import annotations

import time

import numpy as np
import pytest
from scipy.spatial.transform import Rotation as R

from magpylib_material_response_b200 import Collection, Cuboid, Sensor, demag
from magpylib_material_response_b200.demag import _cell_geometry, _CellTable, get_H_ext, get_susceptibilities
from magpylib_material_response_b200.field import _observer_array, cell_arrays
from magpylib_material_response_b200.meshing import mesh_Cuboid
from magpylib_material_response_b200.objects import CellView


def _two_magnets(nnn=(4, 3, 2)):
    hard = Cuboid(polarization=(0, 0, 1), dimension=(1e-3, 1e-3, 2e-3)).move((0, 0, 0.5e-3))
    hard.susceptibility = 0.5
    soft = Cuboid(polarization=(0, 0, 0), dimension=(1e-3, 1e-3, 1e-3))
    soft.rotate_from_angax(45, "y").move((1.5e-3, 0, 0))
    soft.susceptibility = (3999, 100.0, 7.0)
    return hard, soft, Collection(mesh_Cuboid(hard, nnn), mesh_Cuboid(soft, nnn))


def test_mesh_cuboid_is_block_backed_and_views_match_reference_cells():
    c = Cuboid(polarization=(0.1, 0.2, 0.3), dimension=(1, 2, 3), position=(4, 5, 6),
               orientation=R.from_euler("xyz", (10, 20, 30), degrees=True))
    cm = mesh_Cuboid(c, (2, 3, 4))
    assert cm._block is not None and cm._kids is None and len(cm) == 24  # no cell object exists yet
    # the reference's construction (meshing.py:60-95): cells on a local grid, then the parent's pose
    axes = [np.linspace(d / 2 * (1 / n - 1), d / 2 * (1 - 1 / n), n) for d, n in zip((1, 2, 3), (2, 3, 4), strict=True)]
    grid = np.stack(np.meshgrid(*axes, indexing="ij"), axis=-1).reshape(-1, 3)
    want = c.orientation.apply(grid) + c.position
    cells = cm.sources_all
    assert all(isinstance(s, CellView) and isinstance(s, Cuboid) for s in cells)
    np.testing.assert_allclose([s.position for s in cells], want, atol=1e-15)
    np.testing.assert_allclose([s.dimension for s in cells], np.tile((0.5, 2 / 3, 0.75), (24, 1)))
    np.testing.assert_allclose([s.polarization for s in cells], np.tile((0.1, 0.2, 0.3), (24, 1)))
    np.testing.assert_allclose([s.orientation.as_quat() for s in cells], np.tile(c.orientation.as_quat(), (24, 1)))
    assert cells[3].parent is cm and cells[3].barycenter is not None
    # a view writes through to the arrays
    cells[5].polarization = (9, 8, 7)
    np.testing.assert_array_equal(cm._block.pol[5], (9, 8, 7))
    cells[5].position = (1, 1, 1)
    np.testing.assert_array_equal(cm._block.pos[5], (1, 1, 1))


def test_collection_rotate_moves_children():
    """rotating a Collection must rotate its children (about the anchor)."""
    coll = Collection(Cuboid(position=(1, 0, 0), dimension=(1, 1, 1)))
    coll.rotate_from_angax(90, "z", anchor=0)
    np.testing.assert_allclose(coll[0].position, (0, 1, 0), atol=1e-15)
    np.testing.assert_allclose(coll[0].orientation.as_rotvec(), (0, 0, np.pi / 2), atol=1e-15)
    # default anchor = the collection's own position
    coll2 = Collection(Cuboid(position=(2, 0, 0), dimension=(1, 1, 1)), position=(1, 0, 0))
    coll2.rotate_from_angax(180, "z")
    np.testing.assert_allclose(coll2[0].position, (0, 0, 0), atol=1e-15)
    np.testing.assert_allclose(coll2.position, (1, 0, 0), atol=1e-15)


@pytest.mark.parametrize("anchor", [None, (0, 0, 0), (1e-3, -2e-3, 0.5e-3)])
def test_rotated_meshed_collection_equals_mesh_of_rotated_cuboid(anchor):
    rot = R.from_euler("zyx", (30, -50, 70), degrees=True)
    base = Cuboid(polarization=(0, 0, 1), dimension=(1e-3, 2e-3, 3e-3), position=(0.5e-3, 0, 0))
    meshed_then_rotated = mesh_Cuboid(base, (3, 2, 4)).rotate(rot, anchor=anchor)
    assert meshed_then_rotated._block is not None  # stayed an array operation
    rotated_then_meshed = mesh_Cuboid(base.copy().rotate(rot, anchor=anchor), (3, 2, 4))
    a, b = _cell_geometry(meshed_then_rotated.sources_all), _cell_geometry(rotated_then_meshed.sources_all)
    np.testing.assert_allclose(a[0], b[0], atol=1e-17)
    q_same = np.abs(np.sum(a[1] * b[1], axis=1))
    np.testing.assert_allclose(q_same, 1.0, atol=1e-15)
    # nested: rotating the parent of two meshes
    parent = Collection(mesh_Cuboid(base, (2, 2, 2)), mesh_Cuboid(base.copy(position=(0, 1e-3, 0)), (2, 2, 2)))
    ref = [rot.apply(s.position - (0 if anchor is None else np.array(anchor))) + (0 if anchor is None else np.array(anchor))
           for s in parent.sources_all]
    parent.rotate(rot, anchor=anchor)
    np.testing.assert_allclose([s.position for s in parent.sources_all], ref, atol=1e-17)


def test_copy_is_deep_and_lazy():
    _, _, coll = _two_magnets()
    cp = coll.copy()
    assert all(ch._block is not None and ch._kids is None for ch in cp.children)
    cp.children[0]._block.pol[:] = 5.0
    cp.position = (1, 1, 1)
    assert not np.any(coll.children[0]._block.pol == 5.0)
    np.testing.assert_allclose(cp.children[0]._block.pos - coll.children[0]._block.pos, 1.0)
    assert cp.children[1].parent is cp and cp.parent is None


def test_structure_changes_and_paths_fall_back_to_plain_cuboids():
    _, _, coll = _two_magnets()
    m0 = coll.children[0]
    before = _cell_geometry(m0.sources_all)
    view = m0[2]
    m0.add(Sensor())
    assert m0._block is None and type(view) is Cuboid and view.parent is m0  # same object, now standalone
    after = _cell_geometry([s for s in m0.sources_all])
    for x, y in zip(before, after, strict=True):
        np.testing.assert_array_equal(x, y)
    assert view.susceptibility == 0.5
    # a path on the collection
    m1 = coll.children[1]
    m1.move(np.linspace((0, 0, 0), (0, 0, 1e-3), 4), start=0)
    assert m1._block is None and m1[0].position.shape == (4, 3)
    # leaving the collection
    _, _, coll2 = _two_magnets()
    cell = coll2.children[0][0]
    cell.parent = None
    assert type(cell) is Cuboid and len(coll2.children[0]) == 23 and coll2.children[0]._block is None


def test_cell_table_matches_the_per_object_route():
    _, _, coll = _two_magnets()
    coll.children[0].H_ext = (0.0, 1.0, 2.0)
    coll.H_ext = (3.0, 0.0, 0.0)
    lone = Cuboid(polarization=(1, 2, 3), dimension=(1e-3, 1e-3, 1e-3), position=(0, 5e-3, 0),
                  orientation=R.from_euler("x", 33, degrees=True))
    lone.susceptibility = 2.0
    coll.add(lone)
    groups = coll._source_groups()
    assert len(groups) == 3  # two blocks + one object, nothing materialised
    tab = _CellTable(groups)
    objs = coll.sources_all
    assert tab.n == len(objs) == 49
    for x, y in zip(tab.geometry(), _cell_geometry(objs), strict=True):
        np.testing.assert_array_equal(x, y)
    np.testing.assert_array_equal(tab.susceptibilities(None), np.reshape(get_susceptibilities(objs), (49, 3), order="F"))
    np.testing.assert_array_equal(tab.susceptibilities((1.0, 2.0, 3.0)),
                                  np.reshape(get_susceptibilities(objs, (1.0, 2.0, 3.0)), (49, 3), order="F"))
    np.testing.assert_array_equal(tab.H_ext(), np.array(get_H_ext(*objs)))
    pol = tab.pol_local()
    np.testing.assert_array_equal(pol, [s.polarization for s in objs])
    rot = R.from_quat(np.array([s.orientation.as_quat() for s in objs]))
    np.testing.assert_allclose(tab.rotate(pol), rot.apply(pol), atol=1e-16)
    np.testing.assert_allclose(tab.rotate(pol, inverse=True), rot.apply(pol, inverse=True), atol=1e-16)
    # the field front end reads the same arrays
    for x, y in zip(cell_arrays(groups), cell_arrays(objs), strict=True):
        np.testing.assert_array_equal(x, y)
    new = np.arange(49 * 3, dtype=float).reshape(49, 3)
    tab.write_polarizations(new)
    np.testing.assert_array_equal([s.polarization for s in coll.sources_all], new)
    # hierarchy errors are the reference's
    bare = mesh_Cuboid(Cuboid(polarization=(0, 0, 1), dimension=(1, 1, 1)), (2, 2, 2))
    with pytest.raises(ValueError, match="No susceptibility defined in any parent collection"):
        _CellTable(bare._source_groups()).susceptibilities(None)


def test_per_cell_attribute_switches_the_block_to_the_object_route():
    _, _, coll = _two_magnets()
    coll.children[0][1].susceptibility = 77.0
    assert coll.children[0]._block.custom
    groups = coll._source_groups()
    assert len(groups) == 24 + 1
    chi = _CellTable(groups).susceptibilities(None)
    assert chi[1, 0] == 77.0 and chi[0, 0] == 0.5 and chi[24, 0] == 3999
    cp = coll.copy()
    assert cp.children[0][1].susceptibility == 77.0 and cp.children[0][0].susceptibility == 0.5


def test_apply_demag_host_glue_uses_arrays_and_is_cheap(monkeypatch):
    """The whole object-API overhead of apply_demag (copy, gather, rotations, chi / H_ext resolution,
    write-back), native solve stubbed out: <= 5 ms at 8000 cells, <= 40 ms at 64000 (VERDICT r1 item 5;
    the per-object route needed ~50 ms / ~1.1 s)."""
    seen = {}

    def fake_solve(pos, quat, dim, pol_global, chi, H_ext, **kw):
        seen["n"] = len(pos)
        return pol_global * 2.0, {}

    monkeypatch.setattr(demag, "solve_cells_global", fake_solve)
    for nnn, limit_ms in (((20, 20, 10), 5.0), ((40, 40, 20), 40.0)):
        hard, soft, coll = _two_magnets(nnn)
        best = 1e9
        for _ in range(5):
            t0 = time.perf_counter()
            out = demag.apply_demag(coll)
            best = min(best, time.perf_counter() - t0)
        n = 2 * int(np.prod(nnn))
        assert seen["n"] == n and all(ch._kids is None for ch in out.children)  # no cell object was created
        np.testing.assert_allclose(out.children[0]._block.pol, np.tile((0, 0, 2.0), (n // 2, 1)), atol=1e-15)
        assert coll.children[0]._block.pol[0, 2] == 1.0  # input untouched (inplace=False)
        assert best * 1e3 <= limit_ms, f"object-API overhead {best * 1e3:.2f} ms at {n} cells"
        t0 = time.perf_counter()
        mesh_Cuboid(hard, nnn)
        assert (time.perf_counter() - t0) * 1e3 <= limit_ms


def test_sensor_observer_shapes():
    """a Sensor with a position path gives (path, *pixel shape) points."""
    s = Sensor(position=np.linspace((-4, 0, -1), (6, 0, -1), 301))
    pts, shape = _observer_array(s)
    assert pts.shape == (301, 3) and shape == (301,)
    np.testing.assert_allclose(pts[:, 0], np.linspace(-4, 6, 301))
    s2 = Sensor(pixel=np.zeros((2, 3, 3)), position=(1, 2, 3))
    pts, shape = _observer_array(s2)
    assert pts.shape == (6, 3) and shape == (2, 3)
    pts, shape = _observer_array([s2, s2])
    assert pts.shape == (12, 3) and shape == (2, 2, 3)
    pts, shape = _observer_array([s, s])
    assert pts.shape == (602, 3) and shape == (301, 2)
