"""GPU parity tests (run with -m gpu on a B200).  Every check compares the CUDA path -- called
through the C ABI (ctypes) -- with the CPU oracle on the same seeded inputs, with the committed
goldens, or (at full BASELINE sizes) with size-independent properties.

Tolerances (BASELINE.json north_star / SURVEY 7.2):
  T entries : |dT| <= 1e-10 * |T_ref| + 2e-15   (the FP64 closed form itself carries ~1e-16
              absolute error, so a pure rtol is not attainable for far pairs)
  solutions : |dJ| <= 1e-8 * max|J|  (component-wise against the cell's largest component)
"""

from __future__ import annotations

import json

import numpy as np
import pytest
import torch
from scipy.spatial.transform import Rotation as R

import magpylib_material_response_b200 as mmr
from magpylib_material_response_b200 import Collection, Cuboid, _native, apply_demag, demag_tensor, mesh_Cuboid
from magpylib_material_response_b200.demag import _assemble_device, _cell_geometry
from oracle import cuboid_field as ocf
from oracle import demag_ref as oref

pytestmark = pytest.mark.gpu

T_RTOL, T_ATOL = 1e-10, 2e-15
J_RTOL = 1e-8


def assert_T_close(got, ref):
    err = np.abs(got - ref)
    tol = T_RTOL * np.abs(ref) + T_ATOL
    bad = err > tol
    assert not bad.any(), f"{bad.sum()} entries off; worst abs err {err.max():.3e}"


def random_cells(n, seed=0):
    """SURVEY 8d random-geometry stress: rotations, unequal sizes, overlaps."""
    rng = np.random.default_rng(seed)
    pos = rng.uniform(-1e-3, 1e-3, (n, 3))
    dim = rng.uniform(0.05e-3, 0.3e-3, (n, 3))
    quat = R.random(n, rng).as_quat()
    chi = rng.uniform(-0.5, 5, (n, 3))
    pol = rng.normal(0, 1, (n, 3))
    return pos, quat, dim, chi, pol


def cell_major_to_reference(Nmat, n):
    """device matrix [3j+l, 3i+k] -> reference Tmat [l*n+j, k*n+i] (demag.py:452)."""
    return Nmat.reshape(n, 3, n, 3).transpose(1, 0, 3, 2).reshape(3 * n, 3 * n)


def dev(ctx, a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(ctx.device)


def assemble_numpy(ctx, pos, quat, dim, **kw):
    out, ld = _assemble_device(ctx, pos, quat, dim, **kw)
    ctx.sync()
    return out[:, : 3 * len(pos)].cpu().numpy()


# ---------------------------------------------------------------------------------------
# kernel 1: assembly
# ---------------------------------------------------------------------------------------
def test_native_library_is_loaded_and_counts_launches(ctx):
    assert _native.LIB_PATH.exists()
    before = ctx.launch_count
    pos, quat, dim = oref.mesh_cuboid_arrays((1, 1, 1), (2, 2, 2))
    assemble_numpy(ctx, pos, quat, dim)
    assert ctx.launch_count > before


def test_assembly_matches_mpmath_known_answers(ctx, golden_dir):
    data = json.loads((golden_dir / "cuboid_known_answers.json").read_text())
    for case in data["cases"]:
        dimv = np.array(case["dim"])
        off = np.array(case["offset"])
        ref = np.array([[float(v) for v in row] for row in case["block_l_k"]])
        self_case = not off.any()
        pos = np.zeros((1, 3)) if self_case else np.stack([np.zeros(3), off])
        n = len(pos)
        quat = np.tile((0.0, 0, 0, 1), (n, 1))
        dim = np.tile(dimv, (n, 1))
        Nmat = assemble_numpy(ctx, pos, quat, dim)
        blk = Nmat[0:3, 0:3] if self_case else Nmat[3:6, 0:3]  # target 1, source 0
        np.testing.assert_allclose(blk, ref, rtol=1e-12, atol=1e-15)


def test_assembly_matches_mpmath_known_answers_for_rotated_sources(ctx, golden_dir):
    """Rotated source cell (tests/golden/make_rotated_known_answers.py: global-frame surface-charge quadrature)."""
    data = json.loads((golden_dir / "cuboid_rotated_known_answers.json").read_text())
    for case in data["cases"]:
        ref = np.array([[float(v) for v in row] for row in case["block_l_k"]])
        pos = np.stack([np.zeros(3), np.array(case["offset"])])
        quat = np.stack([np.array(case["quat"]), np.array((0.0, 0, 0, 1))])  # the target's own pose is irrelevant
        dim = np.stack([np.array(case["dim"]), np.array((0.1, 0.2, 0.3))])
        Nmat = assemble_numpy(ctx, pos, quat, dim)
        np.testing.assert_allclose(Nmat[3:6, 0:3], ref, rtol=1e-12, atol=2e-15)  # target 1, source 0


@pytest.mark.parametrize("shape", [(6, 6, 6), (5, 3, 7), (1, 1, 1), (1, 1, 2), (33, 1, 1)])
def test_assembly_aligned_cells_vs_oracle(ctx, shape):
    pos, quat, dim = oref.mesh_cuboid_arrays((1e-3, 1.3e-3, 0.7e-3), shape)
    n = len(pos)
    ref = oref.tensor_to_matrix(oref.demag_tensor_ref(pos, quat, dim))
    got = cell_major_to_reference(assemble_numpy(ctx, pos, quat, dim), n)
    assert_T_close(got, ref)
    # trace identity: every self block has trace -1
    np.testing.assert_allclose(np.trace(got), -n, rtol=1e-12)


@pytest.mark.parametrize("n", [1, 31, 257])
def test_assembly_random_rotated_cells_vs_oracle(ctx, n):
    pos, quat, dim, _, _ = random_cells(n, seed=n)
    ref = oref.tensor_to_matrix(oref.demag_tensor_ref(pos, quat, dim))
    got = cell_major_to_reference(assemble_numpy(ctx, pos, quat, dim), n)
    assert_T_close(got, ref)


def test_assembly_two_magnets_rotated_soft_iron_geometry(ctx):
    """Config-2 geometry at reduced size: hard magnet + soft cube rotated 45 deg about y."""
    p1, q1, d1 = oref.mesh_cuboid_arrays((1e-3, 1e-3, 2e-3), (4, 4, 2), position=(0, 0, 0.5e-3))
    p2, q2, d2 = oref.mesh_cuboid_arrays(
        (1e-3, 1e-3, 1e-3), (4, 4, 2), position=(1.5e-3, 0, 0), quat=R.from_euler("y", 45, degrees=True).as_quat()
    )
    pos, quat, dim = np.vstack([p1, p2]), np.vstack([q1, q2]), np.vstack([d1, d2])
    n = len(pos)
    ref = oref.tensor_to_matrix(oref.demag_tensor_ref(pos, quat, dim))
    got = cell_major_to_reference(assemble_numpy(ctx, pos, quat, dim), n)
    assert_T_close(got, ref)


def test_assembly_touching_cells_of_different_size_hit_surface_and_edge_cases(ctx):
    """A small cell whose centre lies exactly on a face / edge / corner-edge extension of a big
    one (SURVEY 7.2 'edge/surface special cases')."""
    pos = np.array([[0, 0, 0], [0.5, 0, 0], [0.5, 0.5, 0], [0.5, 0.5, 0.5], [1.0, 0.5, 0.25], [0.5, 0.2, 0.1]])
    dim = np.array([[1, 1, 1], [0.2, 0.2, 0.2], [0.2, 0.2, 0.2], [0.2, 0.2, 0.2], [0.2, 0.2, 0.2], [0.1, 0.3, 0.2]])
    quat = np.tile((0.0, 0, 0, 1), (len(pos), 1))
    n = len(pos)
    ref = oref.tensor_to_matrix(oref.demag_tensor_ref(pos, quat, dim))
    got = cell_major_to_reference(assemble_numpy(ctx, pos, quat, dim), n)
    assert np.isfinite(got).all()
    assert_T_close(got, ref)


def test_public_demag_tensor_layout_units_split_and_no_mutation(ctx):
    c = Cuboid(dimension=(1e-3, 2e-3, 1e-3), polarization=(0.1, 0.2, 0.3), orientation=R.from_rotvec((0.2, -0.1, 0.4)))
    c.susceptibility = 1.0
    mesh = mesh_Cuboid(c, (3, 4, 2))
    srcs = mesh.sources_all
    pos, quat, dim = _cell_geometry(srcs)
    ref = oref.demag_tensor_ref(pos, quat, dim)
    got = demag_tensor(srcs)
    assert got.shape == (3, 24, 24, 3)
    assert_T_close(got * ocf.MU0, ref * ocf.MU0)
    got_split = demag_tensor(srcs, split=5)
    np.testing.assert_array_equal(got_split, got)
    # the reference mutates polarizations while assembling (demag.py:200-201); we do not
    assert all(np.array_equal(s.polarization, (0.1, 0.2, 0.3)) for s in srcs)


def test_fused_system_matrix_equals_I_minus_S_T(ctx):
    pos, quat, dim, chi, _ = random_cells(97, seed=5)
    n = len(pos)
    Nmat = assemble_numpy(ctx, pos, quat, dim)
    Q = assemble_numpy(ctx, pos, quat, dim, chi_cell_major=chi.reshape(-1))
    expect = np.eye(3 * n) - chi.reshape(-1)[:, None] * Nmat
    np.testing.assert_array_equal(Q, expect)  # the same single FP64 product (Appendix A Q8)
    # and against the reference's dense S@T build
    T = oref.tensor_to_matrix(oref.demag_tensor_ref(pos, quat, dim))
    S = np.diag(oref.susceptibility_vector(chi, n))
    Qref = np.eye(3 * n) - S @ T
    assert_T_close(cell_major_to_reference(Q, n), Qref)


@pytest.mark.parametrize("max_dist", [1.5, 2.0, 3.7, -1.0])
def test_max_dist_masks_exactly_the_reference_pairs(ctx, max_dist):
    pos, quat, dim = oref.mesh_cuboid_arrays((1e-3, 2e-3, 1e-3), (4, 5, 3))
    n = len(pos)
    ref = oref.tensor_to_matrix(oref.demag_tensor_ref(pos, quat, dim, max_dist=max_dist))
    got = cell_major_to_reference(assemble_numpy(ctx, pos, quat, dim, max_dist=max_dist), n)
    # masked pairs (flat index i*n+j, demag.py:245-249) are EXACTLY zero, kept pairs are evaluated
    keep = oref.filter_distance_mask(pos, dim, max_dist).reshape(n, n)  # [i source, j observer]
    blocks = got.reshape(3, n, 3, n).transpose(3, 1, 0, 2)  # [i, j, l, k]
    assert not blocks[~keep].any()
    if keep.any():
        assert np.abs(blocks[keep]).max(axis=(1, 2)).min() > 0
    assert_T_close(got, ref)


def test_pairs_matching_table_finds_the_reference_classes(ctx):
    """Config 3 at reduced size: 2x2x2 lattice of cubes (pitch = multiple of the cell size), each
    5x5x5 cells, anisotropic chi.  The device table must find exactly the classes np.unique finds
    in the reference (demag.py:306-308) and reproduce the dense matrix."""
    parts = []
    for ix in range(2):
        for iy in range(2):
            for iz in range(2):
                parts.append(oref.mesh_cuboid_arrays((1e-3,) * 3, (5, 5, 5), position=(1.5e-3 * ix, 1.5e-3 * iy, 1.5e-3 * iz)))
    pos, quat, dim = (np.vstack([p[k] for p in parts]) for k in range(3))
    n = len(pos)
    chi = np.tile((0.3, 0.2, 0.1), n)
    uniq_ref, inv_ref = oref.match_pairs_indices(pos, quat, dim)
    from magpylib_material_response_b200.demag import _source_classes

    cls = torch.from_numpy(_source_classes(quat, dim)).to(ctx.device)
    d_pos, d_quat, d_dim, d_chi = (dev(ctx, a) for a in (pos, quat, dim, chi))
    ld = 3 * n + 8
    out = ctx.empty(3 * n, ld)
    nuniq = ctx.assemble_matched(d_pos, d_quat, d_dim, cls, out, ld, chi=d_chi)
    assert nuniq == len(uniq_ref)
    assert nuniq <= (2 * 20 - 1) ** 3
    dense = assemble_numpy(ctx, pos, quat, dim, chi_cell_major=chi)
    got = out[:, : 3 * n].cpu().numpy()
    assert_T_close(got, dense)
    # members of one class carry bit-identical blocks (copied from the representative)
    blocks = got.reshape(n, 3, n, 3).transpose(2, 0, 1, 3).reshape(n * n, 9)  # flat index i*n+j
    offdiag = np.ones(n * n, dtype=bool)
    offdiag[:: n + 1] = False  # the self pairs carry the "+1" of Q = I - chi N
    rep_blocks = blocks[uniq_ref][inv_ref]
    # every cell has the same chi triple, so same class => bit-identical Q block
    np.testing.assert_array_equal(blocks[offdiag], rep_blocks[offdiag])
    # a too-small table is reported, not silently wrong
    assert ctx.assemble_matched(d_pos, d_quat, d_dim, cls, out, ld, chi=d_chi, table_cap_slots=1024) is None


def test_block_cyclic_target_sets_tile_the_full_matrix(ctx):
    """Row-block sharding used by the multi-GPU path: the union of all parts == the full matrix."""
    pos, quat, dim, chi, _ = random_cells(100, seed=11)
    n, N = 100, 300
    full = assemble_numpy(ctx, pos, quat, dim, chi_cell_major=chi.reshape(-1))
    d_pos, d_quat, d_dim, d_chi = (dev(ctx, a) for a in (pos, quat, dim, chi.reshape(-1)))
    tgt_block, nparts = 8, 3
    nblk = -(-n // tgt_block)
    rebuilt = np.zeros_like(full)
    for part in range(nparts):
        blocks = list(range(part, nblk, nparts))
        cells = np.concatenate([np.arange(b * tgt_block, min(n, (b + 1) * tgt_block)) for b in blocks])
        out = ctx.empty(3 * len(cells), 304)
        ctx.assemble(d_pos, d_quat, d_dim, out, 304, chi=d_chi, n_tgt_local=len(cells), tgt_block=tgt_block,
                     tgt_nparts=nparts, tgt_part=part)
        ctx.sync()
        loc = out[:, :N].cpu().numpy()
        rows = (3 * cells[:, None] + np.arange(3)[None, :]).reshape(-1)
        rebuilt[rows] = loc
    np.testing.assert_array_equal(rebuilt, full)


# ---------------------------------------------------------------------------------------
# field kernel
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("field", ["B", "H"])
def test_field_kernel_vs_oracle(ctx, field):
    pos, quat, dim, _, pol = random_cells(150, seed=2)
    rng = np.random.default_rng(9)
    obs = np.vstack([rng.uniform(-2e-3, 2e-3, (200, 3)), pos[:5], pos[:5] + dim[:5] * 0.25])
    ref = ocf.getBH_collection(field, obs, pos, quat, dim, pol)
    out = ctx.empty(len(obs), 3)
    ctx.field(dev(ctx, pos), dev(ctx, quat), dev(ctx, dim), dev(ctx, pol), dev(ctx, obs), out,
              _native.FIELD_B if field == "B" else _native.FIELD_H)
    got = out.cpu().numpy()
    scale = np.abs(ref).max()
    np.testing.assert_allclose(got, ref, rtol=1e-9, atol=1e-12 * scale)


# ---------------------------------------------------------------------------------------
# kernel 2a: DGEMM (DMMA) and LU
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize(("m", "n", "k"), [(128, 64, 16), (200, 130, 48), (777, 333, 128), (64, 64, 7), (1500, 1100, 128)])
def test_dgemm_dmma_vs_fp64_reference(ctx, m, n, k):
    rng = np.random.default_rng(m + n + k)
    A = rng.normal(size=(m, k))
    B = rng.normal(size=(k, n))
    C = rng.normal(size=(m, n))
    lda, ldb, ldc = m + 2, k + 4, m + 6
    # column-major buffers with padding
    dA = torch.zeros(k, lda, dtype=torch.float64, device=ctx.device)
    dA[:, :m] = dev(ctx, A.T)
    dB = torch.zeros(n, ldb, dtype=torch.float64, device=ctx.device)
    dB[:, :k] = dev(ctx, B.T)
    dC = torch.zeros(n, ldc, dtype=torch.float64, device=ctx.device)
    dC[:, :m] = dev(ctx, C.T)
    ctx.dgemm_nn(m, n, k, -1.0, dA, lda, dB, ldb, 1.0, dC, ldc)
    ctx.sync()
    got = dC[:, :m].cpu().numpy().T
    ref = C - A @ B
    np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-12 * np.sqrt(k))
    assert torch.all(dC[:, m:] == 0)  # padding untouched


@pytest.mark.parametrize(("m", "n", "k"), [(2048, 1280, 256), (2000, 1400, 256), (4102, 2050, 48), (2304, 2112, 16),
                                         (2001, 1400, 64)])
def test_dgemm_tma_kernels_vs_fp64_reference(ctx, m, n, k):
    """C -= A B on grids of at least one full wave (296 tiles of 128 x 64): the TMA kernels of dgemm_tma.cuh --
    A / B by tensor-map loads, C by bulk copies through shared memory -- with ragged last tiles in m and n;
    an odd m must fall back to the cp.async kernel (bulk copies move multiples of 16 bytes)."""
    rng = np.random.default_rng(m + n + k)
    lda, ldb, ldc = m + 2 + (m % 2), k + 4, m + 6 + (m % 2)
    dA = torch.zeros(k, lda, dtype=torch.float64, device=ctx.device)
    dA[:, :m] = torch.from_numpy(rng.normal(size=(k, m))).to(ctx.device)
    dB = torch.zeros(n, ldb, dtype=torch.float64, device=ctx.device)
    dB[:, :k] = torch.from_numpy(rng.normal(size=(n, k))).to(ctx.device)
    dC = torch.zeros(n, ldc, dtype=torch.float64, device=ctx.device)
    dC[:, :m] = torch.from_numpy(rng.normal(size=(n, m))).to(ctx.device)
    P = dB[:, :k] @ dA[:, :m]  # (column-major views: (A B)^T = B^T A^T)
    ref = dC[:, :m] - 2.0 * P
    for _ in range(2):  # two updates in a row
        ctx.dgemm_nn(m, n, k, -1.0, dA, lda, dB, ldb, 1.0, dC, ldc)
    ctx.sync()
    torch.testing.assert_close(dC[:, :m], ref, rtol=1e-12, atol=1e-12 * np.sqrt(k))
    assert torch.all(dC[:, m:] == 0)  # padding untouched


@pytest.mark.parametrize(("nb1", "ncols"), [(128, 256), (112, 240), (64, 37), (128, 512)])
def test_block_trsm_one_launch_vs_triangular_solve(ctx, nb1, ncols):
    """U12 = L_blk^-1 A12 for a two-panel block in one launch (block_trsm.cuh) against a triangular solve."""
    nb0 = 128
    ob = nb0 + nb1
    g = torch.Generator(device=ctx.device)
    g.manual_seed(nb1 * 1000 + ncols)
    L = torch.tril(torch.randn(ob, ob, dtype=torch.float64, device=ctx.device, generator=g) * 0.05, -1)  # (well conditioned)
    L += torch.eye(ob, dtype=torch.float64, device=ctx.device)
    Bm = torch.randn(ob, ncols, dtype=torch.float64, device=ctx.device, generator=g)
    ref = torch.linalg.solve_triangular(L, Bm, upper=False, unitriangular=True)
    # column-major operands = transposed row-major torch tensors
    Linv0 = torch.linalg.inv(L[:nb0, :nb0]).T.contiguous()
    Linv1 = torch.linalg.inv(L[nb0:, nb0:]).T.contiguous()
    ldl = ob + 6
    L10 = torch.zeros(nb0, ldl, dtype=torch.float64, device=ctx.device)
    L10[:, :nb1] = L[nb0:, :nb0].T
    ldb = ob + 10
    Bc = torch.full((ncols + 3, ldb), 7.0, dtype=torch.float64, device=ctx.device)
    Bc[:ncols, :ob] = Bm.T
    ctx.block_trsm(nb0, nb1, Linv0, Linv1, L10, ldl, Bc, ldb, ncols)
    ctx.sync()
    torch.testing.assert_close(Bc[:ncols, :ob].T, ref, rtol=1e-12, atol=1e-12 * float(ref.abs().max()))
    assert torch.all(Bc[ncols:] == 7.0) and torch.all(Bc[:, ob:] == 7.0)  # nothing outside the block was touched


@pytest.mark.parametrize("N", [5184, 6400])
def test_lu_speculative_with_lookahead_is_reproducible(ctx, N):
    """Diagonally dominant system (speculative panels, deferred TRSM, look-ahead on a second stream: the demag code
    path) at a size where the trailing updates run on the barrier-free tensor-map kernel.  The factorization is
    repeated: every run must give the same bits and a residual at rounding level.  (A stage of that kernel released
    too early showed up exactly here: residual 1e-8 in 4 runs of 5, different every time.)"""
    g = torch.Generator(device=ctx.device)
    g.manual_seed(N)
    ld = (N + 15) // 16 * 16
    Q0 = torch.randn(N, ld, dtype=torch.float64, device=ctx.device, generator=g) * (0.5 / N)
    Q0[:, :N] += torch.eye(N, dtype=torch.float64, device=ctx.device)
    b = torch.randn(N, dtype=torch.float64, device=ctx.device, generator=g)
    first = None
    for _ in range(4):
        Q = Q0.clone()
        x = b.clone()
        ipiv = torch.zeros(N, dtype=torch.int32, device=ctx.device)
        assert ctx.lu_factor(Q, ld, ipiv) == 0
        ctx.lu_solve(Q, ld, ipiv, x)
        ctx.sync()
        r = Q0[:, :N] @ x - b  # (the rows of the buffer are the rows of Q; the LU works on the column-major view Q^T)
        assert float(torch.linalg.norm(r) / torch.linalg.norm(b)) < 1e-13
        if first is None:
            first = Q
        else:
            assert torch.equal(Q, first)


@pytest.mark.parametrize("N", [1, 5, 16, 130, 777, 1000, 2049])
def test_lu_solve_random_dense_vs_numpy(ctx, N):
    """General (non diagonally dominant) matrices: partial pivoting is required."""
    rng = np.random.default_rng(N)
    Q = rng.normal(size=(N, N))
    b = rng.normal(size=(N,))
    ld = (N + 15) // 16 * 16
    dQ = torch.zeros(N, ld, dtype=torch.float64, device=ctx.device)
    dQ[:, :N] = dev(ctx, Q)
    ipiv = torch.zeros(N, dtype=torch.int32, device=ctx.device)
    info = ctx.lu_factor(dQ, ld, ipiv)
    assert info == 0
    dx = dev(ctx, b)
    ctx.lu_solve(dQ, ld, ipiv, dx)
    x = dx.cpu().numpy()
    ref = np.linalg.solve(Q, b)
    resid = np.linalg.norm(Q @ x - b) / (np.linalg.norm(Q) * np.linalg.norm(x))
    assert resid < 1e-14, resid
    cond = np.linalg.cond(Q)
    np.testing.assert_allclose(x, ref, rtol=0, atol=1e-13 * cond * np.abs(ref).max())
    # pivots are a valid LAPACK-style sequence (bit 30 of ipiv[0] is the factor-format flag)
    p = ipiv.cpu().numpy()
    p[0] &= (1 << 30) - 1
    assert np.all(p >= np.arange(N)) and np.all(p < N)


@pytest.mark.parametrize(("N", "ndom"), [(300, 300), (777, 777), (900, 384), (1500, 130), (2049, 1100)])
def test_lu_speculative_panels_then_pivoting(ctx, N, ndom):
    """The first `ndom` columns of Q^T are diagonally dominant (the speculative diagonal-pivot panels
    pass their |l| <= 1 check, pivots = identity there, as LAPACK would choose), the rest is a
    general matrix that needs interchanges (falls back to the pivoting strip kernels)."""
    rng = np.random.default_rng(N + ndom)
    Q = rng.normal(size=(N, N))
    Q[np.arange(ndom), np.arange(ndom)] += 8.0 * np.sqrt(N)
    b = rng.normal(size=(N,))
    ld = (N + 15) // 16 * 16
    dQ = torch.zeros(N, ld, dtype=torch.float64, device=ctx.device)
    dQ[:, :N] = dev(ctx, Q)
    ipiv = torch.zeros(N, dtype=torch.int32, device=ctx.device)
    assert ctx.lu_factor(dQ, ld, ipiv) == 0
    dx = dev(ctx, b)
    ctx.lu_solve(dQ, ld, ipiv, dx)
    x = dx.cpu().numpy()
    resid = np.linalg.norm(Q @ x - b) / (np.linalg.norm(Q) * np.linalg.norm(x))
    assert resid < 1e-14, resid
    p = ipiv.cpu().numpy()
    # bit 30 of ipiv[0]: "inverse format" -- set exactly when no panel needed an interchange; the diagonal
    # panel blocks then hold L11^-1 \ U11^-1 (include/mmr_b200.h, mmr_lu_factor)
    inv_format = bool(p[0] & (1 << 30))
    p[0] &= (1 << 30) - 1
    assert inv_format == (ndom == N)
    nfull = (ndom // 128) * 128 if ndom < N else N  # panels that lie entirely in the dominant part
    assert np.array_equal(p[:nfull], np.arange(nfull))
    # packed factors reproduce Q^T = P L U
    LU = dQ[:, :N].cpu().numpy().T.copy()  # column-major view A = Q^T
    if inv_format:
        for kb in range(0, N, 128):
            nb = min(128, N - kb)
            D = LU[kb:kb + nb, kb:kb + nb]
            L11 = np.linalg.inv(np.tril(D, -1) + np.eye(nb))
            U11 = np.linalg.inv(np.triu(D))
            LU[kb:kb + nb, kb:kb + nb] = np.tril(L11, -1) + U11
    L = np.tril(LU, -1) + np.eye(N)
    U = np.triu(LU)
    PA = Q.T.copy()
    for i in range(N):
        if p[i] != i:
            PA[[i, p[i]]] = PA[[p[i], i]]
    assert np.abs(L @ U - PA).max() < 1e-11 * np.abs(Q).max() * np.sqrt(N)
    assert np.abs(np.tril(LU, -1)).max() <= 1.0 + 1e-12


def test_lu_reports_singular_matrix(ctx):
    N = 40
    Q = np.random.default_rng(0).normal(size=(N, N))
    Q[:, 7] = 0.0  # zero column of Q = zero row of Q^T
    Q[7, :] = 0.0
    dQ = torch.zeros(N, 48, dtype=torch.float64, device=ctx.device)
    dQ[:, :N] = dev(ctx, Q)
    ipiv = torch.zeros(N, dtype=torch.int32, device=ctx.device)
    assert ctx.lu_factor(dQ, 48, ipiv) > 0


# ---------------------------------------------------------------------------------------
# kernel 2b: GMRES
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize(("N", "restart"), [(300, 50), (1001, 20)])
def test_gmres_vs_numpy_solve(ctx, N, restart):
    rng = np.random.default_rng(N)
    Q = np.eye(N) + 0.5 * rng.normal(size=(N, N)) / np.sqrt(N)
    b = rng.normal(size=(N,))
    ld = (N + 15) // 16 * 16
    dQ = torch.zeros(N, ld, dtype=torch.float64, device=ctx.device)
    dQ[:, :N] = dev(ctx, Q)
    db = dev(ctx, b)
    dx = torch.zeros_like(db)
    iters, relres = ctx.gmres(dQ, ld, db, dx, tol=1e-13, restart=restart, maxit=500)
    assert relres <= 1e-13 and 0 < iters < 500
    np.testing.assert_allclose(dx.cpu().numpy(), np.linalg.solve(Q, b), rtol=1e-10, atol=1e-12)
    # matvec export
    dy = torch.zeros_like(db)
    ctx.matvec(dQ, ld, dx, dy)
    np.testing.assert_allclose(dy.cpu().numpy(), b, rtol=1e-11, atol=1e-12)


# ---------------------------------------------------------------------------------------
# end to end through the reference-facing API
# ---------------------------------------------------------------------------------------
ANSYS_CASES = [
    ("isotropic_results_ansys.txt", (0, 0, 1.1), 0.1, 0.0012),
    ("anisotropic_results_ansys.txt", (0, 0, 1.1), (0.3, 0.2, 0.1), 0.0012),
    ("negative_susceptibility_ansys.txt", (0, 0, -0.1), -1.1, 0.0065),
]


@pytest.mark.parametrize("solver", ["lu", "gmres"])
@pytest.mark.parametrize(("fname", "J", "chi", "atol"), ANSYS_CASES)
def test_reference_ansys_goldens_through_the_api(golden_dir, fname, J, chi, atol, solver):
    """The reference's own tests (tests/test_isotropic_anisotropic.py:9-67), verbatim flow:
    Cuboid -> mesh_Cuboid(1000) -> apply_demag(inplace) -> getB(grid) vs ANSYS."""
    magnet = Cuboid(dimension=(1e-3, 1e-3, 1e-3), polarization=J)
    grid = np.loadtxt(golden_dir / "ansys" / "grid_points.pts")
    field_ansys = np.loadtxt(golden_dir / "ansys" / fname, skiprows=1)[:, 3:]
    magnet.susceptibility = chi
    meshed = mesh_Cuboid(magnet, 1000)
    assert apply_demag(meshed, inplace=True, solver=solver) is None
    field = meshed.getB(grid)
    np.testing.assert_allclose(field_ansys, field, rtol=0, atol=atol)
    # and tight parity with the oracle on the same mesh (config 1)
    pos, quat, dim = oref.mesh_cuboid_arrays((1e-3,) * 3, 1000)
    ref = oref.apply_demag_ref(pos, quat, dim, np.tile(np.asarray(J, float), (1000, 1)), chi)
    got = np.array([s.polarization for s in meshed.sources_all])
    scale = np.abs(ref).max(axis=1, keepdims=True)
    assert np.all(np.abs(got - ref) <= J_RTOL * scale)


def test_apply_demag_integration_equivalent_susceptibility_inputs():
    """tests/test_basic.py:16-32 + the SURVEY 8c small known answer."""
    zone = Cuboid(dimension=(1, 1, 1), polarization=(0, 0, 1))
    mesh = mesh_Cuboid(zone, (2, 2, 2))
    dm1 = apply_demag(mesh, susceptibility=4)
    dm2 = apply_demag(mesh, susceptibility=(4, 4, 4))
    zone.susceptibility = 4
    dm3 = apply_demag(mesh_Cuboid(zone, (2, 2, 2)))
    b_ref = dm1.getB((1, 2, 3))
    np.testing.assert_allclose(dm2.getB((1, 2, 3)), b_ref)
    np.testing.assert_allclose(dm3.getB((1, 2, 3)), b_ref)
    np.testing.assert_allclose(b_ref, (4.2419262588694e-4, 8.4895681674371e-4, 6.1656992445783e-4), rtol=1e-10)
    J = np.array([s.polarization for s in dm1.sources_all])
    np.testing.assert_allclose(J[:, 2], 0.435700173485776, rtol=1e-11)
    # input not mutated when inplace=False
    assert all(np.array_equal(s.polarization, (0, 0, 1)) for s in mesh.sources_all)


@pytest.mark.parametrize("solver", ["lu", "gmres"])
@pytest.mark.parametrize("mode", ["plain", "split", "pairs_matching", "max_dist"])
def test_apply_demag_two_magnets_rotated_soft_iron_vs_oracle(solver, mode):
    """Config 2 at reduced size (2 x 6x6x3 cells): hard magnet chi=0.5 + soft cube chi=1000 rotated
    45 deg about y, with H_ext on the soft part; all assembly modes of the reference API."""
    hard = Cuboid(polarization=(0, 0, 1), dimension=(1e-3, 1e-3, 2e-3), position=(0, 0, 0.5e-3))
    hard.susceptibility = 0.5
    soft = Cuboid(polarization=(0, 0, 0), dimension=(1e-3, 1e-3, 1e-3))
    soft.rotate_from_angax(45, "y").move((1.5e-3, 0, 0))
    soft.susceptibility = 1000
    mh, ms = mesh_Cuboid(hard, (6, 6, 3)), mesh_Cuboid(soft, (6, 6, 3))
    ms.H_ext = (0.0, 0.1, 0.2)
    coll = Collection(mh, ms)
    kw = {"plain": {}, "split": {"split": 7}, "pairs_matching": {"pairs_matching": True}, "max_dist": {"max_dist": 6.0}}[mode]
    out = apply_demag(coll, solver=solver, **kw)
    srcs = coll.sources_all
    pos, quat, dim = _cell_geometry(srcs)
    n = len(srcs)
    pol = np.array([s.polarization for s in srcs])
    chi = np.array([[0.5] * 3] * (n // 2) + [[1000.0] * 3] * (n // 2))
    H_ext = np.array([[0.0, 0, 0]] * (n // 2) + [[0.0, 0.1, 0.2]] * (n // 2))
    okw = {"max_dist": 6.0} if mode == "max_dist" else {}
    ref = oref.apply_demag_ref(pos, quat, dim, pol, chi, H_ext=H_ext, **okw)
    got = np.array([s.polarization for s in out.sources_all])
    scale = np.abs(ref).max(axis=1, keepdims=True)
    assert np.all(np.abs(got - ref) <= J_RTOL * scale), np.abs(got - ref).max()


def test_apply_demag_random_anisotropic_stress_vs_oracle():
    pos, quat, dim, chi, pol = random_cells(512, seed=0)
    chi = np.abs(chi) * 0.2  # keep Q comfortably non-singular for a tight comparison
    cells = []
    for i in range(len(pos)):
        c = Cuboid(position=pos[i], orientation=R.from_quat(quat[i]), dimension=dim[i], polarization=pol[i])
        c.susceptibility = tuple(chi[i])
        cells.append(c)
    out = apply_demag(Collection(cells))
    ref = oref.apply_demag_ref(pos, quat, dim, pol, chi)
    got = np.array([s.polarization for s in out.sources_all])
    scale = np.abs(ref).max(axis=1, keepdims=True)
    assert np.all(np.abs(got - ref) <= J_RTOL * scale), np.abs(got - ref).max()


# ---------------------------------------------------------------------------------------
# full BASELINE size (n = 8000, N = 24000): size-independent properties
# ---------------------------------------------------------------------------------------
def test_config3_lattice_anisotropic_pairs_matching_reduced_vs_oracle():
    """BASELINE configs[2] geometry at reduced size (2x2x2 magnets x 5^3 cells; the CPU match_pairs needs
    n^2 x 14 doubles): anisotropic chi, pairs_matching=True, LU and GMRES, against the oracle's own
    match_pairs restatement (demag.py:280-321)."""
    from bench import make_workload

    w = make_workload("lattice_2x5")
    n = len(w["pos"])
    assert n == 1000 and w["pairs_matching"]
    ref = oref.apply_demag_ref(w["pos"], w["quat"], w["dim"], w["pol"], w["chi"], pairs_matching=True)
    ref_dense = oref.apply_demag_ref(w["pos"], w["quat"], w["dim"], w["pol"], w["chi"], split=4)
    scale = np.abs(ref).max(axis=1, keepdims=True)
    assert np.max(np.abs(ref - ref_dense) / scale) < 1e-6  # the reference's own 1e-8 m key rounding
    for solver in ("lu", "gmres"):
        got, info = mmr.demag.apply_demag_arrays(w["pos"], w["quat"], w["dim"], w["pol"], w["chi"], w["H_ext"],
                                                 pairs_matching=True, solver=solver, tol=1e-13, return_info=True)
        assert np.max(np.abs(got - ref_dense) / scale) <= J_RTOL, solver
    # mirror symmetry of the lattice: J_z even, J_x odd under x -> -x
    ctr = w["pos"].mean(axis=0)
    key = np.round((w["pos"] - ctr) * 1e9).astype(np.int64)
    mir = key * np.array([-1, 1, 1])
    perm = np.empty(n, dtype=np.int64)
    perm[np.lexsort(key.T[::-1])] = np.lexsort(mir.T[::-1])
    assert np.abs(got[:, 2] - got[perm][:, 2]).max() <= 1e-12 and np.abs(got[:, 0] + got[perm][:, 0]).max() <= 1e-12


def test_config3_full_size_27000_cells_properties(ctx):
    """BASELINE configs[2] at FULL size (27000 cells, Q = 52.5 GB, anisotropic chi, pairs_matching=True; the
    CPU match_pairs would need 81.6 GB): size-independent properties -- the solution has the mirror symmetries
    of the geometry and satisfies sampled rows of a DENSELY re-assembled system (independent of the dedupe
    table and of GMRES)."""
    from bench import make_workload

    free, _ = torch.cuda.mem_get_info()
    if free < 70e9:
        pytest.skip("needs ~60 GB of device memory")
    w = make_workload("lattice_27000")
    n = len(w["pos"])
    N = 3 * n
    got, info = mmr.demag.apply_demag_arrays(w["pos"], w["quat"], w["dim"], w["pol"], w["chi"], w["H_ext"],
                                             pairs_matching=True, solver="gmres", tol=1e-12, return_info=True)
    assert info["gmres_relres"] <= 1e-12 and info["gmres_iters"] < 30
    # mirror planes through the lattice centre: J_z even, J_x odd under x -> -x; J_z even, J_y odd under y -> -y
    ctr = w["pos"].mean(axis=0)
    key = np.round((w["pos"] - ctr) * 1e9).astype(np.int64)
    order = np.lexsort(key.T[::-1])
    for ax in (0, 1):
        mir = key.copy()
        mir[:, ax] *= -1
        perm = np.empty(n, dtype=np.int64)
        perm[order] = np.lexsort(mir.T[::-1])
        gm = got[perm]
        assert np.abs(got[:, 2] - gm[:, 2]).max() <= 1e-12 * np.abs(got).max()
        assert np.abs(got[:, ax] + gm[:, ax]).max() <= 1e-12 * np.abs(got).max()
    # sampled rows of a dense re-assembly
    ld = (N + 15) // 16 * 16
    d_pos, d_quat, d_dim = ctx.to_device(w["pos"]), ctx.to_device(w["quat"]), ctx.to_device(w["dim"])
    d_chi = ctx.to_device(w["chi"].reshape(N))
    xg = got.reshape(N)  # cells are axis-aligned: local = global frame
    rhs = w["pol"].reshape(N)
    rng = np.random.default_rng(1)
    worst = 0.0
    for c in rng.choice(n, size=24, replace=False):
        row = ctx.empty(3, ld)
        ctx.assemble(d_pos, d_quat, d_dim, row, ld, chi=d_chi, tgt_first=int(c), n_tgt_local=1)
        r = row[:, :N].cpu().numpy() @ xg - rhs[3 * c:3 * c + 3]
        worst = max(worst, float(np.abs(r).max()))
    assert worst <= 1e-10 * np.abs(rhs).max(), worst


def test_config5_magnet_array_max_dist_split_reduced_vs_oracle():
    """BASELINE configs[4] geometry at reduced size (3x3 magnets x 4^3 cells): max_dist truncation with the
    repaired semantics and split chunking, against the oracle."""
    from bench import make_workload

    w = make_workload("array_3x4")
    n = len(w["pos"])
    assert n == 576 and w["max_dist"] == 12.0 and w["split"] == 8
    ref = oref.apply_demag_ref(w["pos"], w["quat"], w["dim"], w["pol"], w["chi"], max_dist=12.0, split=8)
    full = oref.apply_demag_ref(w["pos"], w["quat"], w["dim"], w["pol"], w["chi"])
    scale = np.abs(ref).max(axis=1, keepdims=True)
    assert np.max(np.abs(ref - full) / scale) > 1e-6  # the truncation really removes pairs
    for solver, split in (("lu", 8), ("lu", 1), ("gmres", 8)):
        got = mmr.demag.apply_demag_arrays(w["pos"], w["quat"], w["dim"], w["pol"], w["chi"], w["H_ext"],
                                           max_dist=12.0, split=split, solver=solver, tol=1e-13)
        assert np.max(np.abs(got - ref) / scale) <= J_RTOL, (solver, split)


@pytest.mark.parametrize("solver", ["lu", "gmres"])
def test_demag_store_reuses_the_factorized_system(solver):
    """demag_store (north-star keyword): the second call on the same cells / chi only changes the
    right-hand side -- it must hit the store, skip assembly + factorization and still match the oracle."""
    from bench import make_workload

    w = make_workload("two_magnets_432")
    store = {}
    kw = dict(solver=solver, tol=1e-13, return_info=True, demag_store=store)
    x1, info1 = mmr.demag.apply_demag_arrays(w["pos"], w["quat"], w["dim"], w["pol"], w["chi"], w["H_ext"], **kw)
    assert info1["demag_store"] == "miss" and len(store) == 1
    rng = np.random.default_rng(3)
    pol2 = rng.normal(size=w["pol"].shape)
    H2 = rng.normal(size=w["H_ext"].shape) * 0.1
    x2, info2 = mmr.demag.apply_demag_arrays(w["pos"], w["quat"], w["dim"], pol2, w["chi"], H2, **kw)
    assert info2["demag_store"] == "hit" and len(store) == 1
    assert info2["assemble_ms"] < 0.5 * info1["assemble_ms"]
    for x, pol, H in ((x1, w["pol"], w["H_ext"]), (x2, pol2, H2)):
        ref = oref.apply_demag_ref(w["pos"], w["quat"], w["dim"], pol, w["chi"], H_ext=H)
        assert np.max(np.abs(x - ref) / np.abs(ref).max(axis=1, keepdims=True)) <= J_RTOL
    # a different chi is a different system
    chi3 = w["chi"] * 1.5
    _, info3 = mmr.demag.apply_demag_arrays(w["pos"], w["quat"], w["dim"], pol2, chi3, H2, **kw)
    assert info3["demag_store"] == "miss" and len(store) == 2


def test_demag_store_through_the_object_api():
    cube = Cuboid(polarization=(0, 0, 0), dimension=(1e-3, 1e-3, 2e-3))
    cube.susceptibility = 80.0
    mesh = mesh_Cuboid(cube, 96)
    store = {}
    outs = []
    for hz in (0.1, 0.2):
        mesh.H_ext = (0, 0, hz)
        out, info = apply_demag(mesh, demag_store=store, return_info=True)
        outs.append(np.array([s.polarization for s in out.sources_all]))
    assert info["demag_store"] == "hit" and len(store) == 1
    np.testing.assert_allclose(outs[1], 2.0 * outs[0], rtol=1e-10)  # linear in the excitation
    with pytest.raises(TypeError, match="demag_store must be None or a dict-like mapping"):
        apply_demag(mesh, demag_store=42)


def test_config2prime_8000_cells_properties(ctx):
    """North-star target config: soft-iron cube 20x20x20, chi=1000, H_ext=(0,0,1)."""
    cube = Cuboid(dimension=(1e-3,) * 3, polarization=(0, 0, 0))
    cube.susceptibility = 1000
    mesh = mesh_Cuboid(cube, (20, 20, 20))
    mesh.H_ext = (0.0, 0.0, 1.0)
    n, N = 8000, 24000
    pos, quat, dim = _cell_geometry(mesh.sources_all)
    # (1) tensor: trace of every self block is -1; symmetric for identical cells; rows sum rule
    Nmat, ld = _assemble_device(ctx, pos, quat, dim)
    A = Nmat[:, :N]
    assert abs(float(torch.diagonal(A).sum()) + n) < 1e-9
    assert float((A - A.T).abs().max()) < 4e-15  # 2x the closed form's absolute error
    # translation invariance: block (i, j) depends on the displacement only
    blk = A.reshape(n, 3, n, 3)
    np.testing.assert_allclose(blk[0, :, 1, :].cpu().numpy(), blk[400, :, 401, :].cpu().numpy(), rtol=1e-12, atol=2e-15)
    # spot-check 64 random rows against the oracle
    rng = np.random.default_rng(1)
    jj = rng.choice(n, 8, replace=False)
    H = np.stack([
        ocf.getBH_cuboid_sources("H", pos[jj], pos, quat, dim, np.tile(np.eye(3)[k], (n, 1))) for k in range(3)
    ])  # [k, i, jsel, l]
    ref_rows = ocf.MU0 * H.transpose(2, 3, 1, 0).reshape(len(jj) * 3, N)  # [(jsel,l), (i,k)]
    rows = (3 * jj[:, None] + np.arange(3)[None, :]).reshape(-1)
    assert_T_close(A[rows].cpu().numpy(), ref_rows)
    del Nmat, A, blk
    # (2) solve: both solvers agree with each other and satisfy the system
    out_lu, info_lu = apply_demag(mesh, solver="lu", return_info=True)
    out_gm, info_gm = apply_demag(mesh, solver="gmres", return_info=True)
    J_lu = np.array([s.polarization for s in out_lu.sources_all])
    J_gm = np.array([s.polarization for s in out_gm.sources_all])
    assert np.abs(J_lu - J_gm).max() <= J_RTOL * np.abs(J_lu).max()
    assert info_gm["gmres_relres"] <= 1e-12
    chi = np.full(N, 1000.0)
    Q, ld = _assemble_device(ctx, pos, quat, dim, chi_cell_major=chi)
    x = torch.from_numpy(J_lu.reshape(-1)).to(ctx.device)
    y = torch.empty_like(x)
    ctx.matvec(Q, ld, x, y)
    rhs = np.tile((0.0, 0.0, 1000.0), n)
    resid = np.linalg.norm(y.cpu().numpy() - rhs) / np.linalg.norm(rhs)
    assert resid < 1e-11, resid
    # physics: mean J_z between the n=1000 (3.5716) and n=2744 (3.5948) values of SURVEY 8c, x/y vanish
    assert 3.55 < J_lu[:, 2].mean() < 3.65
    assert abs(J_lu[:, 0].mean()) < 1e-9


# ---------------------------------------------------------------------------------------
# distributed entry points on ONE rank (the panel-by-panel code path; R > 1 is exercised by
# tests/multi_gpu_parity.py under torchrun and by the gloo CPU tests)
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize(("n", "tgt_block"), [(50, 8), (333, 40), (700, 42), (700, 80), (1000, 85), (640, 64)])
def test_lu_solve_dist_single_rank_vs_numpy(ctx, n, tgt_block):
    N = 3 * n
    rng = np.random.default_rng(n)
    Q = rng.normal(size=(N, N))
    b = rng.normal(size=N)
    ld = (N + 15) // 16 * 16
    dQ = torch.zeros(N, ld, dtype=torch.float64, device=ctx.device)
    dQ[:, :N] = dev(ctx, Q)
    dx = dev(ctx, b)
    info = ctx.lu_solve_dist(N, N, tgt_block, dQ, ld, dx)
    assert info == 0
    x = dx.cpu().numpy()
    resid = np.linalg.norm(Q @ x - b) / (np.linalg.norm(Q) * np.linalg.norm(x))
    assert resid < 1e-14, resid


def test_gmres_dist_single_rank_matches_gmres(ctx):
    n, tgt_block = 200, 40
    pos, quat, dim = oref.mesh_cuboid_arrays((1e-3, 2e-3, 1e-3), (5, 8, 5))
    chi = np.full(3 * n, 50.0)
    Q = assemble_numpy(ctx, pos, quat, dim, chi_cell_major=chi)
    N = 3 * n
    ld = (N + 15) // 16 * 16
    dQ = torch.zeros(N, ld, dtype=torch.float64, device=ctx.device)
    dQ[:, :N] = dev(ctx, Q)
    b = np.tile((0.0, 0.0, 50.0), n)
    db = dev(ctx, b)
    dx = torch.zeros_like(db)
    iters, relres = ctx.gmres_dist(N, N, tgt_block, dQ, ld, db, dx, tol=1e-13, restart=100, maxit=1000)
    assert relres <= 1e-13
    np.testing.assert_allclose(dx.cpu().numpy(), np.linalg.solve(Q, b), rtol=1e-9, atol=1e-11)


def test_lu_large_panel_fallback_path(ctx):
    """N > 25600 rows: the early strips exceed the single-cluster capacity and go through the
    cooperative grid-barrier kernel (the path every early panel of the 64k-cell config takes)."""
    N = 26000
    g = torch.Generator(device=ctx.device)
    g.manual_seed(1)
    ld = (N + 15) // 16 * 16
    Q = torch.randn(N, ld, dtype=torch.float64, device=ctx.device, generator=g)
    b = torch.randn(N, dtype=torch.float64, device=ctx.device, generator=g)
    Q0 = Q.clone()
    ipiv = torch.zeros(N, dtype=torch.int32, device=ctx.device)
    assert ctx.lu_factor(Q, ld, ipiv) == 0
    x = b.clone()
    ctx.lu_solve(Q, ld, ipiv, x)
    y = torch.empty_like(b)
    ctx.matvec(Q0, ld, x, y)
    resid = float(torch.linalg.norm(y - b) / (torch.linalg.norm(Q0[:, :N]) * torch.linalg.norm(x)))
    assert resid < 1e-14, resid


def test_pairs_matching_row_shards_tile_the_dense_matrix(ctx):
    """Multi-GPU pairs_matching (mmr_assemble_matched_rows): each rank's unique-pair table covers its own
    block-cyclic target rows; the union of the shards reproduces the dense matrix, and a shard's table is
    smaller than the full one."""
    from magpylib_material_response_b200.demag import _source_classes

    parts = [oref.mesh_cuboid_arrays((1e-3,) * 3, (4, 4, 4), position=(1.5e-3 * ix, 0, 1.5e-3 * iz))
             for ix in range(2) for iz in range(2)]
    pos, quat, dim = (np.vstack([p[k] for p in parts]) for k in range(3))
    n, N = len(pos), 3 * len(pos)
    chi = np.tile((0.3, 0.2, 0.1), n)
    cls = torch.from_numpy(_source_classes(quat, dim)).to(ctx.device)
    d_pos, d_quat, d_dim, d_chi = (dev(ctx, a) for a in (pos, quat, dim, chi))
    dense = assemble_numpy(ctx, pos, quat, dim, chi_cell_major=chi)
    ld = N + 8
    full = ctx.empty(N, ld)
    nuniq_full = ctx.assemble_matched(d_pos, d_quat, d_dim, cls, full, ld, chi=d_chi)
    tgt_block, nparts = 24, 3
    nblk = -(-n // tgt_block)
    rebuilt = np.zeros_like(dense)
    for part in range(nparts):
        cells = np.concatenate([np.arange(b * tgt_block, min(n, (b + 1) * tgt_block)) for b in range(part, nblk, nparts)])
        out = ctx.empty(3 * len(cells), ld)
        nuniq = ctx.assemble_matched(d_pos, d_quat, d_dim, cls, out, ld, chi=d_chi, n_tgt_local=len(cells),
                                     tgt_block=tgt_block, tgt_nparts=nparts, tgt_part=part)
        assert 0 < nuniq <= nuniq_full
        rows = (3 * cells[:, None] + np.arange(3)[None, :]).reshape(-1)
        rebuilt[rows] = out[:, :N].cpu().numpy()
    assert_T_close(rebuilt, dense)
    assert_T_close(full[:, :N].cpu().numpy(), dense)


@pytest.mark.parametrize(("n", "tgt_block"), [(333, 40), (700, 80)])
def test_lu_dist_factor_once_solve_many(ctx, n, tgt_block):
    """The two phases of the distributed LU on their own (demag_store on several GPUs keeps the factors)."""
    N = 3 * n
    rng = np.random.default_rng(n + 1)
    Q = rng.normal(size=(N, N))
    ld = (N + 15) // 16 * 16
    dQ = torch.zeros(N, ld, dtype=torch.float64, device=ctx.device)
    dQ[:, :N] = dev(ctx, Q)
    ipiv = torch.zeros(N, dtype=torch.int32, device=ctx.device)
    assert ctx.lu_factor_dist(N, N, tgt_block, dQ, ld, ipiv) == 0
    for seed in range(3):
        b = np.random.default_rng(seed).normal(size=N)
        dx = dev(ctx, b)
        ctx.lu_trisolve_dist(N, N, tgt_block, dQ, ld, ipiv, dx)
        x = dx.cpu().numpy()
        resid = np.linalg.norm(Q @ x - b) / (np.linalg.norm(Q) * np.linalg.norm(x))
        assert resid < 1e-14, resid


def test_two_magnets_8000_full_size_vs_oracle(ctx, golden_dir):
    """The bench workload itself (BASELINE configs[1], N = 24000) against the CPU oracle: 64 sampled rows of
    the assembled tensor (the rotated chi = 1000 cube's rows included) computed by the oracle here, and the
    FULL solution against the oracle solve committed in tests/golden/two_magnets_8000_solution.npz
    (generated by tests/golden/make_two_magnets_8000_solution.py: 66 s of CPU on 8 cores)."""
    import bench

    w = bench.make_workload("two_magnets_8000")
    pos, quat, dim = w["pos"], w["quat"], w["dim"]
    n, N = len(pos), 3 * len(pos)
    Nmat, ld = _assemble_device(ctx, pos, quat, dim)
    rng = np.random.default_rng(7)
    jj = np.sort(np.concatenate([rng.choice(4000, 32, replace=False), 4000 + rng.choice(4000, 32, replace=False)]))
    H = np.stack([ocf.getBH_cuboid_sources("H", pos[jj], pos, quat, dim,
                                           R.from_quat(quat).apply(np.tile(np.eye(3)[k], (n, 1)), inverse=True))
                  for k in range(3)])  # [k, i, jsel, l]
    ref_rows = ocf.MU0 * H.transpose(2, 3, 1, 0).reshape(len(jj) * 3, N)  # [(jsel,l), (i,k)]
    rows = (3 * jj[:, None] + np.arange(3)[None, :]).reshape(-1)
    assert_T_close(Nmat[rows][:, :N].cpu().numpy(), ref_rows)
    del Nmat
    gold = np.load(golden_dir / "two_magnets_8000_solution.npz")["x_local"]
    meta = json.loads((golden_dir / "two_magnets_8000_solution.json").read_text())
    import hashlib

    assert hashlib.sha256(np.ascontiguousarray(gold).tobytes()).hexdigest() == meta["sha256_x_local"]
    from magpylib_material_response_b200.demag import apply_demag_arrays

    for solver in ("lu", "gmres"):
        x = apply_demag_arrays(pos, quat, dim, w["pol"], w["chi"], w["H_ext"], solver=solver, tol=1e-13)
        err = np.abs(x - gold).max() / np.abs(gold).max()
        assert err <= J_RTOL, (solver, err)
    # and through the object API (SoA collection of mesh_Cuboid cells)
    out = apply_demag(bench.make_collection("two_magnets_8000"))
    err = np.abs(bench.collection_polarizations(out) - gold).max() / np.abs(gold).max()
    assert err <= J_RTOL, err


@pytest.mark.parametrize("nb", [128, 112, 96, 33, 1])
def test_lu_diag_block_cluster_kernel_vs_numpy(ctx, nb):
    """The one-launch diagonal-block step of the speculative panel (panel128.cuh): no-pivot LU + L^-1 + U^-1
    against a numpy elimination, and the rejection of a block that needs an interchange."""
    rng = np.random.default_rng(nb)
    A = rng.normal(size=(nb, nb)) * 0.2
    A[np.arange(nb), np.arange(nb)] += 2.0 + nb * 0.2  # diagonally dominant: every multiplier stays below 1
    ld = nb + 5
    dA = torch.zeros(nb, ld, dtype=torch.float64, device=ctx.device)  # column-major A = rows of this tensor
    dA[:, :nb] = dev(ctx, A.T)
    dS = torch.zeros(nb, ld, dtype=torch.float64, device=ctx.device)
    dL = torch.zeros(nb, nb, dtype=torch.float64, device=ctx.device)
    dU = torch.zeros(nb, nb, dtype=torch.float64, device=ctx.device)
    assert ctx.lu_diag_block(dA, ld, nb, dS, ld, dL, dU) is False
    LU = A.copy()
    for j in range(nb):
        LU[j + 1:, j] /= LU[j, j]
        LU[j + 1:, j + 1:] -= np.outer(LU[j + 1:, j], LU[j, j + 1:])
    L = np.tril(LU, -1) + np.eye(nb)
    U = np.triu(LU)
    got = dS[:, :nb].cpu().numpy().T
    np.testing.assert_allclose(got, LU, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(dL.cpu().numpy().T, np.linalg.inv(L), rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(dU.cpu().numpy().T, np.linalg.inv(U), rtol=1e-11, atol=1e-13)
    if nb > 1:
        B = A.copy()
        B[nb // 2, nb // 2] = 1e-3  # the column below holds larger entries: partial pivoting would interchange
        dA[:, :nb] = dev(ctx, B.T)
        assert ctx.lu_diag_block(dA, ld, nb, dS, ld, dL, dU) is True


# ---- FEM (ANSYS) datasets of the reference's docs through the OBJECT API (rotated Cuboids, Collections,
#      Sensors with position paths) -- same fixtures and tolerances as tests/test_oracle_golden.py ---------------
@pytest.mark.parametrize("solver", ["lu", "gmres"])
def test_fem_softmag_dataset_through_the_object_api(golden_dir, solver):
    from magpylib_material_response_b200 import Sensor
    from test_oracle_golden import SOFTMAG_TOL, softmag_cells

    fem = np.loadtxt(golden_dir / "femdata" / "FEMdata_test_softmag.csv", delimiter=",", skiprows=1)
    cube1 = Cuboid(polarization=(0, 0, 1), dimension=(0.001, 0.001, 0.002))
    cube1.move((0, 0, 0.0005))
    cube1.susceptibility = 0.5
    cube2 = Cuboid(polarization=(0, 0, 0), dimension=(0.001, 0.001, 0.001))
    cube2.rotate_from_angax(angle=45, axis="y").move((0.0015, 0, 0))
    cube2.susceptibility = 3999
    coll = Collection(mesh_Cuboid(cube1, (6, 6, 12)), mesh_Cuboid(cube2, (8, 8, 8)), style_label="meshed")
    sensors = [Sensor(position=np.linspace((-0.004, 0, z), (0.006, 0, z), 1001)) for z in (-0.001, -0.003, -0.005)]
    out = apply_demag(coll, solver=solver, tol=1e-13)
    B = out.getB(sensors)  # (path, sensor, 3)
    assert B.shape == (1001, 3, 3)
    for k in range(3):
        np.testing.assert_allclose(B[:, k, 0], fem[:, 1 + 2 * k], rtol=0, atol=SOFTMAG_TOL[k])
        np.testing.assert_allclose(B[:, k, 2], fem[:, 2 + 2 * k], rtol=0, atol=SOFTMAG_TOL[k])
    # and the oracle on the same cells: solved polarizations within the north-star tolerance
    pos, quat, dim, pol, chi = softmag_cells()
    ref = oref.apply_demag_ref(pos, quat, dim, pol, chi, split=4)
    got = np.array([s.polarization for s in out.sources_all])
    assert np.abs(got - ref).max() <= J_RTOL * np.abs(ref).max()
    assert out.getB(sensors[0]).shape == (1001, 3)


def test_fem_three_cuboids_dataset_through_the_object_api(golden_dir):
    from magpylib_material_response_b200 import Sensor, mesh_all
    from test_oracle_golden import CUBOIDS_TOL, cuboids_cells

    fem = np.loadtxt(golden_dir / "femdata" / "FEMdata_test_cuboids.csv", delimiter=",", skiprows=1)
    cube1 = Cuboid(polarization=(0, 0, 1), dimension=(0.001, 0.001, 0.001))
    cube1.move((-0.0015, 0, 0))
    cube1.susceptibility = 0.3
    cube2 = Cuboid(polarization=(0.9, 0, 0), dimension=(0.001, 0.001, 0.001))
    cube2.rotate_from_angax(-45, "y").move((0, 0, 0.0002))
    cube2.susceptibility = 1.0
    mx, my = 0.6 * np.sin(np.deg2rad(30)), 0.6 * np.cos(np.deg2rad(30))
    cube3 = Cuboid(polarization=(mx, my, 0), dimension=(0.001, 0.001, 0.002))
    cube3.move((0.0016, 0, 0.0005)).rotate_from_angax(30, "z")
    cube3.susceptibility = 0.5
    coll = Collection(mesh_Cuboid(cube1, (5, 5, 5)), mesh_Cuboid(cube2, (5, 5, 5)), mesh_Cuboid(cube3, (5, 5, 10)))
    sensor = Sensor(position=np.linspace((-0.004, 0, -0.001), (0.004, 0, -0.001), 301))
    out = apply_demag(coll)
    B = out.getB(sensor)
    assert B.shape == (301, 3)
    for k in range(3):
        np.testing.assert_allclose(B[:, k], fem[:, 1 + k], rtol=0, atol=CUBOIDS_TOL[k])
    pos, quat, dim, pol, chi = cuboids_cells()
    ref = oref.apply_demag_ref(pos, quat, dim, pol, chi, split=4)
    got = np.array([s.polarization for s in out.sources_all])
    assert np.abs(got - ref).max() <= J_RTOL * np.abs(ref).max()
    # the docs' own route: mesh_all over the un-meshed collection (volume-apportioned cell counts)
    meshed = mesh_all(Collection(cube1.copy(), cube2.copy(), cube3.copy()), target_elems=400, per_child_elems=False)
    B2 = apply_demag(meshed).getB(sensor)
    assert np.abs(B2 - B).max() < 6e-4  # two different meshes of the same problem


def test_gmres_block_jacobi_preconditioner_two_magnets():
    """Hard magnet + rotated soft cube (chi = 1000): block-Jacobi over the soft magnet cuts the iteration count
    several-fold and reaches the same solution (true-residual stopping test unchanged)."""
    import bench
    from magpylib_material_response_b200.demag import apply_demag_arrays

    w = bench.make_workload("two_magnets_1728")
    geo = (w["pos"], w["quat"], w["dim"], w["pol"], w["chi"], w["H_ext"])
    ref = oref.apply_demag_ref(*geo[:5], H_ext=w["H_ext"], split=4)
    n = len(ref)
    x0, i0 = apply_demag_arrays(*geo, solver="gmres", tol=1e-12, return_info=True)
    x1, i1 = apply_demag_arrays(*geo, solver="gmres", tol=1e-12, return_info=True, groups=[n // 2, n // 2])
    assert i1["gmres_precond_blocks"] == [(3 * (n // 2), 3 * (n // 2))]
    assert i1["gmres_relres"] <= 1e-12 and i0["gmres_relres"] <= 1e-12
    assert i1["gmres_iters"] * 3 <= i0["gmres_iters"], (i0["gmres_iters"], i1["gmres_iters"])
    for x in (x0, x1):
        assert np.abs(x - ref).max() <= J_RTOL * np.abs(ref).max()
    # the object API passes the magnets of the collection by itself
    out, info = apply_demag(bench.make_collection("two_magnets_1728"), solver="gmres", tol=1e-12, return_info=True)
    assert info["gmres_precond_blocks"] == i1["gmres_precond_blocks"] and info["gmres_iters"] == i1["gmres_iters"]
    assert np.abs(bench.collection_polarizations(out) - ref).max() <= J_RTOL * np.abs(ref).max()
