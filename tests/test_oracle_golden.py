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
