"""CPU: pin the oracle (oracle/) against every golden the reference holds for the path and
against the independent mpmath known answers (tests/golden/make_known_answers.py)."""

from __future__ import annotations

import json

import numpy as np
import pytest

from oracle.cuboid_field import (
    MU0,
    BHJM_magnet_cuboid,
    getBH_collection,
    getBH_cuboid_sources,
)
from oracle.demag_ref import (
    apply_demag_ref,
    demag_tensor_ref,
    mesh_cuboid_arrays,
    tensor_to_matrix,
)

# (file, J, chi, atol) exactly as /root/reference/tests/test_isotropic_anisotropic.py:9-67
ANSYS_CASES = [
    ("isotropic_results_ansys.txt", (0, 0, 1.1), 0.1, 0.0012),
    ("anisotropic_results_ansys.txt", (0, 0, 1.1), (0.3, 0.2, 0.1), 0.0012),
    ("negative_susceptibility_ansys.txt", (0, 0, -0.1), -1.1, 0.0065),
]


@pytest.mark.parametrize(("fname", "J", "chi", "atol"), ANSYS_CASES)
def test_oracle_reproduces_reference_ansys_goldens(golden_dir, fname, J, chi, atol):
    grid = np.loadtxt(golden_dir / "ansys" / "grid_points.pts")
    field_ansys = np.loadtxt(golden_dir / "ansys" / fname, skiprows=1)[:, 3:]
    pos, quat, dim = mesh_cuboid_arrays((1e-3, 1e-3, 1e-3), 1000)
    assert len(pos) == 1000
    pol = np.tile(np.asarray(J, dtype=float), (len(pos), 1))
    pol_new = apply_demag_ref(pos, quat, dim, pol, chi)
    B = getBH_collection("B", grid, pos, quat, dim, pol_new)
    np.testing.assert_allclose(field_ansys, B, rtol=0, atol=atol)


def test_oracle_matches_mpmath_known_answers(golden_dir):
    data = json.loads((golden_dir / "cuboid_known_answers.json").read_text())
    assert len(data["cases"]) >= 10
    for case in data["cases"]:
        dim = np.array(case["dim"])
        off = np.array(case["offset"])
        ref = np.array([[float(v) for v in row] for row in case["block_l_k"]])
        got = np.empty((3, 3))
        for k in range(3):
            pol = np.zeros(3)
            pol[k] = 1.0
            got[:, k] = MU0 * BHJM_magnet_cuboid("H", off[None, :], dim[None, :], pol[None, :])[0]
        # FP64 closed form carries ~1e-16 absolute error
        np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-15)


def test_oracle_small_known_answer_2x2x2():
    """unit cube meshed (2,2,2), J=(0,0,1), chi=4 (= tests/test_basic.py:18-22)."""
    pos, quat, dim = mesh_cuboid_arrays((1, 1, 1), (2, 2, 2))
    pol = np.tile((0.0, 0.0, 1.0), (8, 1))
    out, parts = apply_demag_ref(pos, quat, dim, pol, 4.0, return_parts=True)
    np.testing.assert_allclose(out[:, 2], 0.435700173485776, rtol=1e-12)
    np.testing.assert_allclose(np.abs(out[:, 0]), 0.0355223040232513, rtol=1e-11)
    assert np.sign(out[:, 0]).tolist() == [1, -1, 1, -1, -1, 1, -1, 1]
    assert np.sign(out[:, 1]).tolist() == [1, -1, -1, 1, 1, -1, -1, 1]
    T = parts["T"]
    assert np.abs(T - T.T).max() < 1e-15
    np.testing.assert_allclose(np.trace(T) / T.shape[0], -1 / 3, rtol=1e-13)
    B = getBH_collection("B", np.array([[1.0, 2.0, 3.0]]), pos, quat, dim, out)[0]
    np.testing.assert_allclose(
        B, (4.2419262588694e-4, 8.4895681674371e-4, 6.1656992445783e-4), rtol=1e-10
    )


def test_oracle_special_cases():
    dim = np.array([[1.0, 2.0, 3.0]])
    pol = np.array([[0.3, -0.2, 0.9]])
    # on an edge (x = a, y = b, |z| < c): B = 0, H = -J/mu0 (inside by the surface band)
    H = BHJM_magnet_cuboid("H", np.array([[0.5, 1.0, 0.2]]), dim, pol)
    np.testing.assert_allclose(H * MU0, -pol)
    B = BHJM_magnet_cuboid("B", np.array([[0.5, 1.0, 0.2]]), dim, pol)
    assert np.all(B == 0)
    # zero polarization / zero dimension
    assert np.all(BHJM_magnet_cuboid("H", np.array([[2.0, 0, 0]]), dim, np.zeros((1, 3))) == 0)
    assert np.all(BHJM_magnet_cuboid("H", np.array([[2.0, 0, 0]]), np.array([[1.0, 0, 1]]), pol) == 0)
    # inside: trace of the block is -1; outside: 0
    for off, tr in (((0.1, -0.3, 0.7), -1.0), ((1.1, 0.3, -2.7), 0.0)):
        blk = np.stack(
            [MU0 * BHJM_magnet_cuboid("H", np.array([off]), dim, np.eye(3)[k][None])[0] for k in range(3)],
            axis=1,
        )
        np.testing.assert_allclose(np.trace(blk), tr, atol=1e-14)


def test_oracle_split_pairs_maxdist_consistency():
    """split / pairs_matching change nothing; max_dist zeroes exactly the masked pairs."""
    rng = np.random.default_rng(3)
    pos, quat, dim = mesh_cuboid_arrays((1e-3, 2e-3, 1e-3), (3, 4, 2))
    T_plain = demag_tensor_ref(pos, quat, dim)
    T_split = demag_tensor_ref(pos, quat, dim, split=5)
    np.testing.assert_array_equal(T_plain, T_split)
    T_pairs = demag_tensor_ref(pos, quat, dim, pairs_matching=True)
    np.testing.assert_allclose(T_pairs, T_plain, rtol=1e-9, atol=1e-9 * np.abs(T_plain).max())
    with pytest.raises(ValueError, match="Pairs matching does not support splitting"):
        demag_tensor_ref(pos, quat, dim, pairs_matching=True, split=2)
    T_md = demag_tensor_ref(pos, quat, dim, max_dist=1.5)
    n = len(pos)
    d = np.linalg.norm(pos[None] - pos[:, None], axis=-1) / dim.max()
    keep = d < 1.5
    assert 0 < keep.sum() < n * n
    np.testing.assert_array_equal(T_md[:, keep, :], T_plain[:, keep, :])
    assert np.all(T_md[:, ~keep, :] == 0)
    del rng


def test_oracle_rotated_frame_handling():
    """Rotating the whole system rotates the tensor blocks: T' = R T R^T per pair."""
    from scipy.spatial.transform import Rotation as R

    rot = R.from_euler("zyx", [33.0, -21.0, 64.0], degrees=True)
    pos, quat, dim = mesh_cuboid_arrays((1e-3, 1.5e-3, 0.5e-3), (2, 3, 2))
    pos_r, quat_r, dim_r = mesh_cuboid_arrays((1e-3, 1.5e-3, 0.5e-3), (2, 3, 2), quat=rot.as_quat())
    np.testing.assert_allclose(pos_r, rot.apply(pos), atol=1e-18)
    T = tensor_to_matrix(demag_tensor_ref(pos, quat, dim))
    Tr = tensor_to_matrix(demag_tensor_ref(pos_r, quat_r, dim_r))
    n = len(pos)
    Rm = rot.as_matrix()
    T4 = T.reshape(3, n, 3, n)  # [l, j, k, i]
    Tr4 = Tr.reshape(3, n, 3, n)
    expect = np.einsum("al,ljki,bk->ajbi", Rm, T4, Rm)
    np.testing.assert_allclose(Tr4, expect, atol=2e-15)


def test_oracle_object_interface_shape():
    pos, quat, dim = mesh_cuboid_arrays((1, 1, 1), (1, 2, 2))
    pol = np.tile((1.0, 0, 0), (4, 1))
    H = getBH_cuboid_sources("H", pos, pos, quat, dim, pol)
    assert H.shape == (4, 4, 3)


# ---- the FEM (ANSYS) comparison datasets the reference ships with its docs (tests/golden/femdata/, copied by
#      copy_from_reference.py): the only reference-held data with ROTATED cells and several magnets -------------
# tolerances = the discretisation gap between the point-matching model at this mesh and the FEM solution,
# measured once (n = 118 cells: 1.5e-3 / 2.2e-3 T on the z = -1 mm line -> 7.3e-4 at n = 944: it converges)
SOFTMAG_TOL = {0: 7.6e-4, 1: 1.7e-5, 2: 3.0e-6}  # sensor line k (z = -1, -3, -5 mm)
CUBOIDS_TOL = (8.5e-4, 4.5e-4, 5.2e-4)  # Bx, By, Bz at 3 x 5x5x5(x2) cells


def softmag_cells(nh=(6, 6, 12), ns=(8, 8, 8)):
    """docs/examples/soft_magnets.md:47-57: hard cuboid chi 0.5 + soft cube chi 3999 rotated 45 deg about y."""
    from scipy.spatial.transform import Rotation as R

    p1, q1, d1 = mesh_cuboid_arrays((1e-3, 1e-3, 2e-3), nh, position=(0, 0, 0.5e-3))
    p2, q2, d2 = mesh_cuboid_arrays((1e-3,) * 3, ns, position=(1.5e-3, 0, 0), quat=R.from_euler("y", 45, degrees=True).as_quat())
    pol = np.vstack([np.tile((0, 0, 1.0), (len(p1), 1)), np.zeros((len(p2), 3))])
    chi = np.concatenate([np.full(len(p1), 0.5), np.full(len(p2), 3999.0)])
    return np.vstack([p1, p2]), np.vstack([q1, q2]), np.vstack([d1, d2]), pol, chi


def cuboids_cells(a=5):
    """docs/examples/cuboids_demagnetization.md:55-72: chi 0.3 / 1.0 (rotated -45 deg about y) / 0.5 (30 deg about z)."""
    from scipy.spatial.transform import Rotation as R

    parts = [mesh_cuboid_arrays((1e-3,) * 3, (a, a, a), position=(-1.5e-3, 0, 0)),
             mesh_cuboid_arrays((1e-3,) * 3, (a, a, a), position=(0, 0, 0.2e-3), quat=R.from_euler("y", -45, degrees=True).as_quat()),
             mesh_cuboid_arrays((1e-3, 1e-3, 2e-3), (a, a, 2 * a), position=(1.6e-3, 0, 0.5e-3),
                                quat=R.from_euler("z", 30, degrees=True).as_quat())]
    pos, quat, dim = (np.vstack([p[k] for p in parts]) for k in range(3))
    ns = [len(p[0]) for p in parts]
    J3 = (0.6 * np.sin(np.deg2rad(30)), 0.6 * np.cos(np.deg2rad(30)), 0.0)
    pol = np.vstack([np.tile((0, 0, 1.0), (ns[0], 1)), np.tile((0.9, 0, 0), (ns[1], 1)), np.tile(J3, (ns[2], 1))])
    chi = np.concatenate([np.full(ns[0], 0.3), np.full(ns[1], 1.0), np.full(ns[2], 0.5)])
    return pos, quat, dim, pol, chi


def test_oracle_vs_fem_softmag_rotated_soft_cube(golden_dir):
    fem = np.loadtxt(golden_dir / "femdata" / "FEMdata_test_softmag.csv", delimiter=",", skiprows=1)
    assert fem.shape == (1001, 7)
    pos, quat, dim, pol, chi = softmag_cells()
    assert len(pos) == 944
    x = apply_demag_ref(pos, quat, dim, pol, chi, split=4)
    xs = np.linspace(-4e-3, 6e-3, 1001)
    for k, z in enumerate((-1e-3, -3e-3, -5e-3)):
        obs = np.stack([xs, np.zeros_like(xs), np.full_like(xs, z)], axis=1)
        B = getBH_collection("B", obs, pos, quat, dim, x)
        np.testing.assert_allclose(B[:, 0], fem[:, 1 + 2 * k], rtol=0, atol=SOFTMAG_TOL[k])
        np.testing.assert_allclose(B[:, 2], fem[:, 2 + 2 * k], rtol=0, atol=SOFTMAG_TOL[k])
        assert np.abs(B[:, 1]).max() < 1e-12  # mirror symmetry in y
    # without the material response the z = -1 mm line is off by ~1.9e-2 T: the fixture does pin the solve
    B0 = getBH_collection("B", obs * 0 + np.stack([xs, 0 * xs, np.full_like(xs, -1e-3)], axis=1), pos, quat, dim, pol)
    assert np.abs(B0[:, 2] - fem[:, 2]).max() > 1e-2


def test_oracle_vs_fem_three_cuboids(golden_dir):
    fem = np.loadtxt(golden_dir / "femdata" / "FEMdata_test_cuboids.csv", delimiter=",", skiprows=1)
    assert fem.shape == (301, 4)
    pos, quat, dim, pol, chi = cuboids_cells()
    x = apply_demag_ref(pos, quat, dim, pol, chi, split=4)
    xs = np.linspace(-4e-3, 4e-3, 301)
    obs = np.stack([xs, np.zeros_like(xs), np.full_like(xs, -1e-3)], axis=1)
    B = getBH_collection("B", obs, pos, quat, dim, x)
    for k in range(3):
        np.testing.assert_allclose(B[:, k], fem[:, 1 + k], rtol=0, atol=CUBOIDS_TOL[k])


def test_oracle_matches_mpmath_known_answers_for_rotated_sources(golden_dir):
    """Rotated source cells: R^T r / R.H_local / R^T e_k frame handling against the global-frame surface-charge
    quadrature of tests/golden/make_rotated_known_answers.py."""
    from oracle.cuboid_field import quat_to_matrix

    data = json.loads((golden_dir / "cuboid_rotated_known_answers.json").read_text())
    assert len(data["cases"]) >= 4
    for case in data["cases"]:
        dim, quat, off = np.array(case["dim"]), np.array(case["quat"]), np.array(case["offset"])
        ref = np.array([[float(v) for v in row] for row in case["block_l_k"]])
        Rm = quat_to_matrix(quat)[0]
        got = np.empty((3, 3))
        for k in range(3):
            pol_local = Rm.T @ np.eye(3)[k]  # demag.py:183
            got[:, k] = MU0 * getBH_cuboid_sources("H", off[None, :], np.zeros((1, 3)), quat[None, :], dim[None, :],
                                                   pol_local[None, :])[0, 0]
        np.testing.assert_allclose(got, ref, rtol=1e-12, atol=2e-15)
