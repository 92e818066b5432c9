"""CPU: host-side logic of the drop-in API.  Mirrors the reference's own unit tests
(/root/reference/tests/test_basic.py:35-185, tests/test_meshing.py:18-56) on the shim
objects, plus the C-ABI export check.  No compute call needs a GPU here."""

from __future__ import annotations

import ctypes
import pathlib
import re

import numpy as np
import pytest
from scipy.spatial.transform import Rotation as R

import magpylib_material_response_b200 as mmr
from magpylib_material_response_b200 import Collection, Cuboid, Sensor, _native
from magpylib_material_response_b200.demag import (
    _cell_geometry,
    apply_demag,
    demag_tensor,
    filter_distance,
    get_H_ext,
    get_susceptibilities,
    match_pairs,
)
from magpylib_material_response_b200.meshing import cells_from_dimension, mesh_Cuboid
from oracle import demag_ref

ROOT = pathlib.Path(__file__).resolve().parent.parent


def _cube(**kw):
    return Cuboid(dimension=(1, 1, 1), polarization=(0, 0, 1), **kw)


# ---- get_susceptibilities (tests/test_basic.py:35-185) ---------------------------------
@pytest.mark.parametrize(
    ("case", "sus", "expected"),
    [
        ("source", [(2.5,), (3.0,)], [2.5, 3.0, 2.5, 3.0, 2.5, 3.0]),
        ("source", [(1.0, 2.0, 3.0), (4.0, 5.0, 6.0)], [1.0, 4.0, 2.0, 5.0, 3.0, 6.0]),
        ("function", 1.5, [1.5] * 6),
        ("function", (2.0, 3.0, 4.0), [2.0, 2.0, 3.0, 3.0, 4.0, 4.0]),
        ("function", [1.5, 2.5], [1.5, 2.5, 1.5, 2.5, 1.5, 2.5]),
    ],
)
def test_get_susceptibilities_basic(case, sus, expected):
    sources = [_cube(), _cube()]
    if case == "source":
        for s, v in zip(sources, sus, strict=True):
            s.susceptibility = v[0] if len(v) == 1 else v
        result = get_susceptibilities(sources)
    else:
        result = get_susceptibilities(sources, susceptibility=sus)
    np.testing.assert_allclose(result, expected)


def test_get_susceptibilities_hierarchy():
    coll = Collection()
    coll.susceptibility = 2.0
    own = _cube()
    own.susceptibility = 5.0
    inherit = _cube()
    coll.add(inherit)
    np.testing.assert_allclose(get_susceptibilities([own, inherit]), [5.0, 2.0, 5.0, 2.0, 5.0, 2.0])
    np.testing.assert_allclose(get_susceptibilities([inherit]), [2.0, 2.0, 2.0])
    # grand-parent
    top = Collection(coll)
    del coll.susceptibility
    top.susceptibility = (1.0, 2.0, 3.0)
    np.testing.assert_allclose(get_susceptibilities([inherit]), [1.0, 2.0, 3.0])


def test_get_susceptibilities_errors():
    with pytest.raises(ValueError, match="No susceptibility defined in any parent collection"):
        get_susceptibilities([_cube()])
    bad = _cube()
    bad.susceptibility = (1, 2)
    with pytest.raises(ValueError, match="susceptibility is not scalar or array of length 3"):
        get_susceptibilities([bad])
    with pytest.raises(ValueError, match="must be scalar, 3-vector, or same length as input Collection"):
        get_susceptibilities([_cube() for _ in range(4)], susceptibility=[1.0, 2.0, 3.0, 4.0, 5.0])
    with pytest.raises(ValueError, match="Apply_demag input susceptibility is ambiguous"):
        get_susceptibilities([_cube() for _ in range(3)], susceptibility=(1.0, 2.0, 3.0))


def test_get_susceptibilities_edge_cases():
    assert len(get_susceptibilities([])) == 0
    z = _cube()
    z.susceptibility = 3.0
    np.testing.assert_allclose(get_susceptibilities([z]), [3.0, 3.0, 3.0])


def test_get_H_ext_resolution():
    coll = Collection()
    coll.H_ext = (0.0, 0.0, 1.0)
    a, b = _cube(), _cube()
    b.H_ext = (1.0, 0.0, 0.0)
    coll.add(a, b)
    lone = _cube()
    assert get_H_ext(a, b, lone) == [(0.0, 0.0, 1.0), (1.0, 0.0, 0.0), (0.0, 0.0, 0.0)]


# ---- mesh_Cuboid (tests/test_meshing.py:18-56) -----------------------------------------
def test_mesh_Cuboid_with_path_and_style():
    N = 5
    c = Cuboid(polarization=(1, 0, 0), dimension=(0.001, 0.002, 0.003), position=[0.004, 0.005, 0.006])
    c.style.label = "Cuboid 1"
    c.style.opacity = 0
    c.rotate_from_angax(np.linspace(0, 76, N), "z", anchor=0, start=0)
    c.move(np.linspace((0, 0, 0), (0, 0, 0.0105), N), start=0)
    c.susceptibility = 3999
    cm = mesh_Cuboid(c, target_elems=8, style_opacity=0.5)
    assert len(cm.sources_all) == 8
    assert all(s.susceptibility == c.susceptibility for s in cm.sources_all)
    assert cm.style.label == c.style.label
    assert cm.style.opacity == 0.5
    # every cell follows the parent's 5-step path: cell centre = parent pose applied to the grid
    assert c.position.shape == (N, 3)
    for s in cm.sources_all:
        assert s.position.shape == (N, 3)
    centre = np.mean([s.position for s in cm.sources_all], axis=0)
    np.testing.assert_allclose(centre, c.position, atol=1e-15)
    np.testing.assert_allclose(cm.sources_all[0].orientation.as_quat(), c.orientation.as_quat())
    with pytest.raises(TypeError, match="Object to be meshed must be a Cuboid"):
        mesh_Cuboid(Sensor(), target_elems=8)


def test_mesh_Cuboid_explicit_divisions_and_order():
    c = Cuboid(dimension=(1, 1, 1))
    cm = mesh_Cuboid(c, target_elems=[1, 2, 3])
    poss = np.array([o.position for o in cm])
    assert [len(np.unique(poss[:, i])) for i in range(3)] == [1, 2, 3]
    # z fastest, x outer (meshing.py:87)
    assert np.all(np.diff(poss[:3, 2]) > 0)
    c2 = Cuboid(dimension=(1e-3, 2e-3, 3e-3), position=(1, 2, 3), orientation=R.from_rotvec((0.1, 0.2, 0.3)))
    c2.polarization = (0.1, 0.2, 0.3)
    cm2 = mesh_Cuboid(c2, 100)
    pos, quat, dim = _cell_geometry(cm2.sources_all)
    p_ref, q_ref, d_ref = demag_ref.mesh_cuboid_arrays(
        (1e-3, 2e-3, 3e-3), 100, position=(1, 2, 3), quat=c2.orientation.as_quat()
    )
    np.testing.assert_allclose(pos, p_ref, rtol=0, atol=1e-15)
    np.testing.assert_allclose(quat, q_ref, atol=1e-15)
    np.testing.assert_allclose(dim, d_ref, rtol=1e-15)
    assert all(np.array_equal(s.polarization, (0.1, 0.2, 0.3)) for s in cm2.sources_all)


@pytest.mark.parametrize(
    ("kw", "expected"),
    [
        ({"parity": None, "strict_max": True}, [4, 9, 25]),
        ({"parity": None, "strict_max": False}, [4, 9, 26]),
        ({"parity": "odd", "strict_max": True}, [3, 11, 27]),
        ({"parity": "odd", "strict_max": False}, [5, 7, 27]),
        ({"parity": "even", "strict_max": True}, [4, 8, 26]),
        ({"parity": "even", "strict_max": False}, [4, 10, 24]),
    ],
)
def test_cells_from_dimension_docstring_examples(kw, expected):
    # the worked examples of meshing_utils.py:59-71
    assert cells_from_dimension([1, 2, 6], 926, **kw).tolist() == expected
    assert demag_ref.cells_from_dimension([1, 2, 6], 926, **kw).tolist() == expected


def test_collection_copy_is_deep_and_detached():
    c = _cube()
    c.susceptibility = 4
    mesh = mesh_Cuboid(c, (2, 2, 2))
    top = Collection(mesh)
    cp = mesh.copy()
    assert cp.parent is None
    assert len(cp.sources_all) == 8
    assert all(a is not b for a, b in zip(cp.sources_all, mesh.sources_all, strict=True))
    assert all(s.parent is cp for s in cp.sources_all)
    cp.sources_all[0].polarization = (9, 9, 9)
    assert np.array_equal(mesh.sources_all[0].polarization, (0, 0, 1))
    assert mesh.parent is top


def test_collection_pose_moves_children():
    a = Cuboid(dimension=(1, 1, 1), position=(1, 0, 0))
    coll = Collection(a)
    coll.position = (0, 0, 2)
    np.testing.assert_allclose(a.position, (1, 0, 2))
    coll.orientation = R.from_euler("z", 90, degrees=True)
    np.testing.assert_allclose(a.position, (0, 1, 2), atol=1e-15)
    np.testing.assert_allclose(a.orientation.as_euler("zyx", degrees=True), (90, 0, 0), atol=1e-12)


# ---- apply_demag argument handling (demag.py:380-408, :165-167, :437-439) ---------------
def test_apply_demag_error_paths_before_any_gpu_work():
    c = _cube()
    c.susceptibility = 1.0
    mesh = mesh_Cuboid(c, (1, 1, 2))
    with pytest.raises(ValueError, match="Pairs matching does not support splitting"):
        apply_demag(mesh, pairs_matching=True, split=2)
    with pytest.raises(ValueError, match="Pairs matching does not support splitting"):
        demag_tensor(mesh.sources_all, pairs_matching=True)  # default split=False (Appendix A Q4)
    pathy = _cube()
    pathy.susceptibility = 1.0
    pathy.move(np.linspace((0, 0, 0), (0, 0, 1), 3), start=0)
    with pytest.raises(ValueError, match="1 objects with paths, found"):
        apply_demag(Collection(pathy))

    class Odd(mmr.objects.BaseSource):
        pass

    class OddColl(Collection):
        @property
        def sources_all(self):
            return [Odd()]

    with pytest.raises(TypeError, match="Only Magnet and Current sources supported. Incompatible objects found: 1 Odd"):
        apply_demag(OddColl())
    with pytest.raises(ValueError, match="No susceptibility defined"):
        apply_demag(mesh_Cuboid(_cube(), (1, 1, 2)))
    with pytest.raises(ValueError, match="solver must be"):
        apply_demag(mesh, solver="cholesky")
    # empty collection: nothing to solve, no GPU needed
    out = apply_demag(Collection())
    assert isinstance(out, Collection)
    assert apply_demag(Collection(), inplace=True) is None


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    c = _cube()
    c.susceptibility = 1.0
    mesh = mesh_Cuboid(c, (1, 1, 2))
    with pytest.raises(_native.NativeUnavailableError, match="no CPU fallback"):
        apply_demag(mesh)
    with pytest.raises(_native.NativeUnavailableError):
        mesh.getB((1, 2, 3))
    # input must not have been touched
    assert np.array_equal(mesh.sources_all[0].polarization, (0, 0, 1))


def test_host_filter_distance_and_match_pairs_agree_with_oracle():
    c = Cuboid(dimension=(1e-3, 2e-3, 1e-3), polarization=(0, 0, 1))
    mesh = mesh_Cuboid(c, (3, 4, 2))
    srcs = mesh.sources_all
    pos, quat, dim = _cell_geometry(srcs)
    mask = filter_distance(srcs, 1.5)
    np.testing.assert_array_equal(mask, demag_ref.filter_distance_mask(pos, dim, 1.5))
    mask2, params, pos0, rot0 = filter_distance(srcs, 1.5, return_params=True, return_base_geo=True)
    assert params["observers"].shape == (mask.sum(), 3)
    _, uniq, inv, _, _ = match_pairs(srcs)
    uniq_ref, inv_ref = demag_ref.match_pairs_indices(pos, quat, dim)
    # identical cells: same classes as the reference's key (Appendix A Q3)
    assert len(uniq) == len(uniq_ref)
    assert np.array_equal(uniq[inv], uniq_ref[inv_ref])


# ---- C ABI ---------------------------------------------------------------------------
def test_c_abi_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "mmr_b200.h").read_text()
    declared = set(re.findall(r"\b(mmr_[a-z0-9_]+)\s*\(", header))
    declared.discard("mmr_ctx")
    assert declared == set(_native.EXPORTED_SYMBOLS), declared ^ set(_native.EXPORTED_SYMBOLS)
    lib = _native.load_library()  # dlopen only; no GPU call
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.mmr_abi_version() == 1
    # argument validation happens before any CUDA call
    assert lib.mmr_lu_factor(None, 3, None, 3, None, None) == -1
    assert b"ctx is NULL" in lib.mmr_last_error()
    assert lib.mmr_ctx_destroy(None) == 0
    handle = ctypes.c_void_p()
    import torch

    if not torch.cuda.is_available():
        status = lib.mmr_ctx_create(0, None, ctypes.byref(handle))
        assert status != 0
        assert b"no CPU fallback" in lib.mmr_last_error()


def test_gmres_block_jacobi_block_selection():
    from magpylib_material_response_b200.demag import _precond_blocks

    chi = np.vstack([np.full((4000, 3), 0.5), np.full((4000, 3), 1000.0)])
    assert _precond_blocks([4000, 4000], chi) == [(12000, 12000)]  # only the soft magnet
    assert _precond_blocks([8000], np.full((8000, 3), 1000.0)) is None  # one magnet: the block would be Q itself
    assert _precond_blocks([4000, 4000], np.full((8000, 3), 0.5)) is None  # nothing ill conditioned
    assert _precond_blocks(None, chi) is None
    many = np.full((2700, 3), 50.0)
    assert _precond_blocks([100] * 27, many) == [(300 * k, 300) for k in range(27)]
    assert _precond_blocks([10] * 270, many) is None  # tiny magnets are left alone
