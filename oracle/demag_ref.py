"""CPU ORACLE (test infrastructure, NOT product code) -- the demagnetization path.

Array-level restatement of /root/reference/src/magpylib_material_response/demag.py
(demag_tensor :124-224, filter_distance :227-277, match_pairs :280-321, the system
build and solve of apply_demag :420-476) and of mesh_Cuboid's cell generation
(meshing.py:60-95, meshing_utils.py:8-110).  Objects are replaced by plain arrays:
``pos`` (n,3) global cell barycenters, ``quat`` (n,4) scalar-last orientation,
``dim`` (n,3) cell dimensions, ``pol`` (n,3) LOCAL-frame polarizations (tesla).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this file.  See oracle/cuboid_field.py for the pinning status.
"""

from __future__ import annotations

from itertools import product

import numpy as np

from .cuboid_field import MU0, getBH_cuboid_pairs, getBH_cuboid_sources, quat_to_matrix


# ----------------------------------------------------------------------------------
# meshing (meshing_utils.py:8-110, meshing.py:60-95)
# ----------------------------------------------------------------------------------
def apportion_triple(triple, min_val=1, max_iter=30):
    """meshing_utils.py:8-23"""
    triple = np.abs(np.array(triple, dtype=float))
    count = 0
    while any(n < min_val for n in triple) and count < max_iter:
        count += 1
        amin, amax = triple.argmin(), triple.argmax()
        factor = min_val / triple[amin]
        if triple[amax] >= factor * min_val:
            triple /= factor**0.5
            triple[amin] *= factor**1.5
    return triple


def cells_from_dimension(dim, target_elems, min_val=1, strict_max=False, parity=None):
    """meshing_utils.py:26-110"""
    elems = np.prod(target_elems)
    if parity == "odd":
        funcs = [
            lambda x, add=add, fn=fn: int(2 * fn(x / 2) + add)
            for add in (-1, 1)
            for fn in (np.ceil, np.floor)
        ]
    elif parity == "even":
        funcs = [lambda x, fn=fn: int(2 * fn(x / 2)) for fn in (np.ceil, np.floor)]
    else:
        funcs = [np.ceil, np.floor]
    elems = max(min_val**3, elems)
    x, y, z = np.abs(dim)
    a = x ** (2 / 3) * (elems / y / z) ** (1 / 3)
    b = y ** (2 / 3) * (elems / x / z) ** (1 / 3)
    c = z ** (2 / 3) * (elems / x / y) ** (1 / 3)
    a, b, c = apportion_triple((a, b, c), min_val=min_val)
    epsilon = elems
    result = [funcs[0](k) for k in (a, b, c)]
    for fns in product(*[funcs] * 3):
        res = [f(k) for f, k in zip(fns, (a, b, c), strict=False)]
        epsilon_new = elems - np.prod(res)
        if (
            np.abs(epsilon_new) <= epsilon
            and all(r >= min_val for r in res)
            and (not strict_max or epsilon_new >= 0)
        ):
            epsilon = np.abs(epsilon_new)
            result = res
    return np.array(result).astype(int)


def mesh_cuboid_arrays(dimension, target_elems, position=(0, 0, 0), quat=(0, 0, 0, 1)):
    """Cell arrays of ``mesh_Cuboid`` (meshing.py:60-95 + _collection_from_obj_and_cells
    :19-29): cell index = (ix*ny + iy)*nz + iz, absolute positions pos + R grid,
    every cell carries the parent's orientation.  Returns (pos, quat, dim)."""
    dim0 = np.asarray(dimension, dtype=float)
    if np.isscalar(target_elems):
        nnn = cells_from_dimension(dim0, target_elems)
    else:
        nnn = target_elems
    nnn = np.array(nnn, dtype=int)
    new_dim = dim0 / nnn
    xs, ys, zs = (
        np.linspace(d / 2 * (1 / n - 1), d / 2 * (1 - 1 / n), n)
        for d, n in zip(dim0, nnn, strict=False)
    )
    grid = np.array([(x, y, z) for x in xs for y in ys for z in zs])
    Rm = quat_to_matrix(quat)[0]
    pos = np.asarray(position, dtype=float) + grid @ Rm.T
    n = len(grid)
    return pos, np.tile(np.asarray(quat, dtype=float), (n, 1)), np.tile(new_dim, (n, 1))


# ----------------------------------------------------------------------------------
# demag tensor (demag.py:124-321)
# ----------------------------------------------------------------------------------
def filter_distance_mask(pos, dim, max_dist):
    """demag.py:245-249: flat index i*n+j (source i, observer j)."""
    n = len(pos)
    pos2 = np.tile(pos, (n, 1)) - np.repeat(pos, n, axis=0)
    dist2 = np.linalg.norm(pos2, axis=1)
    dim2 = np.tile(dim, (n, 1)), np.repeat(dim, n, axis=0)
    maxdim2 = np.concatenate(dim2, axis=1).max(axis=1)
    return (dist2 / maxdim2) < max_dist


def match_pairs_indices(pos, quat, dim):
    """demag.py:291-308: unique ordered pairs keyed on
    [p_j - p_i, quat_j, quat_i, dim_j - dim_i] + 1e-9 rounded to 8 decimals."""
    n = len(pos)
    pos2 = np.tile(pos, (n, 1)) - np.repeat(pos, n, axis=0)
    rotQ2a = np.tile(quat, (n, 1)).reshape((n * n, -1))
    rotQ2b = np.repeat(quat, n, axis=0).reshape((n * n, -1))
    dim2 = np.tile(dim, (n, 1)) - np.repeat(dim, n, axis=0)
    prop = (np.concatenate([pos2, rotQ2a, rotQ2b, dim2], axis=1) + 1e-9).round(8)
    _, unique_inds, unique_inv_inds = np.unique(
        prop, return_index=True, return_inverse=True, axis=0
    )
    return unique_inds, np.asarray(unique_inv_inds).reshape(-1)


def demag_tensor_ref(pos, quat, dim, pairs_matching=False, split=1, max_dist=0, threads=1):
    """(3,n,n,3) array [k unit-pol][i source][j observer][l comp], A/m per tesla.

    ``threads`` > 1 evaluates the ``split`` source chunks on a thread pool (numpy ufuncs release the
    GIL): the same arithmetic, used by bench.py so that the CPU arm gets every host core -- the reference
    itself runs this part on one thread.

    Restates demag.py:124-224 including the three passes (one per unit polarization,
    :182), ``split`` chunking of the source list (:202-218), the pairs_matching
    gather (:198) and the max_dist scatter (:193-196).  The reference's max_dist path
    raises as written (4-way unpack of a 3-tuple, :171-173); this follows its evident
    intent (return_params=True) -- SURVEY Appendix A Q2.
    """
    pos = np.atleast_2d(np.asarray(pos, dtype=float))
    quat = np.atleast_2d(np.asarray(quat, dtype=float))
    dim = np.atleast_2d(np.asarray(dim, dtype=float))
    n = len(pos)
    if pairs_matching and split != 1:
        msg = "Pairs matching does not support splitting"
        raise ValueError(msg)
    Rm = quat_to_matrix(quat)

    mask_inds = None
    unique_inv_inds = None
    if max_dist != 0:
        mask_inds = filter_distance_mask(pos, dim, max_dist)
    elif pairs_matching:
        mask_inds, unique_inv_inds = match_pairs_indices(pos, quat, dim)

    if mask_inds is not None:
        obs_p = np.tile(pos, (n, 1))[mask_inds]
        pos_p = np.repeat(pos, n, axis=0)[mask_inds]
        quat_p = np.repeat(quat, n, axis=0)[mask_inds]
        dim_p = np.repeat(dim, n, axis=0)[mask_inds]

    H_point = []
    for unit_pol in [(1, 0, 0), (0, 1, 0), (0, 0, 1)]:
        # rot0.inv().apply(unit_pol): global unit vector in every cell's local frame
        pol_all = np.einsum("pji,j->pi", Rm, np.asarray(unit_pol, dtype=float))
        if mask_inds is not None:
            polarization = np.repeat(pol_all, n, axis=0)[mask_inds]
            H_unique = getBH_cuboid_pairs("H", obs_p, pos_p, quat_p, dim_p, polarization)
            if max_dist != 0:
                H_temp = np.zeros((n * n, 3))
                H_temp[mask_inds] = H_unique
                H_unit_pol = H_temp
            else:
                H_unit_pol = H_unique[unique_inv_inds]
        elif split > 1:
            chunks = [idx for idx in np.array_split(np.arange(n), split) if idx.size > 0]

            def one(idx, pol_all=pol_all):
                return getBH_cuboid_sources("H", pos, pos[idx], quat[idx], dim[idx], pol_all[idx])

            if threads > 1:
                from concurrent.futures import ThreadPoolExecutor

                with ThreadPoolExecutor(max_workers=threads) as pool:
                    parts = list(pool.map(one, chunks))
            else:
                parts = [one(idx) for idx in chunks]
            H_unit_pol = np.concatenate(parts, axis=0)
        else:
            H_unit_pol = getBH_cuboid_sources("H", pos, pos, quat, dim, pol_all)
        H_point.append(H_unit_pol)
    return np.array(H_point).reshape((3, n, n, 3))


def tensor_to_matrix(T4):
    """demag.py:451-452: Tmat[l*n+j, k*n+i] = mu0 * T4[k,i,j,l] (dimensionless)."""
    n = T4.shape[1]
    T = T4 * MU0
    return T.swapaxes(2, 3).reshape((3 * n, 3 * n)).T


# ----------------------------------------------------------------------------------
# system build + solve (demag.py:420-476)
# ----------------------------------------------------------------------------------
def susceptibility_vector(chi, n):
    """(3n,) component-major vector from scalar / 3-vector / (n,) / (n,3) input
    (subset of demag.py:17-102 needed at array level)."""
    chi = np.asarray(chi, dtype=float)
    if chi.ndim == 0:
        susis = np.ones((n, 3)) * chi
    elif chi.shape == (3,) and n != 3:
        susis = np.tile(chi, (n, 1))
    elif chi.shape == (n,):
        susis = np.repeat(chi[:, None], 3, axis=1)
    elif chi.shape == (n, 3):
        susis = chi
    else:
        msg = "bad susceptibility shape"
        raise ValueError(msg)
    return np.reshape(susis, 3 * n, order="F")


def apply_demag_ref(
    pos,
    quat,
    dim,
    pol,
    chi,
    H_ext=None,
    pairs_matching=False,
    max_dist=0,
    split=1,
    return_parts=False,
    timings=None,
    threads=1,
):
    """Solved LOCAL-frame polarizations (n,3).  ``timings``: optional dict that receives the wall seconds of the
    three stages SURVEY 8d names (assembly, building Q with the dense S@T, LAPACK solve).

    Restates demag.py:420-476: J0 = R_i J_i stacked order="F" (:420-428), S = diag(chi)
    dense (:432), T (:444-452), Q = I - S@T via dense matmul (:466), np.linalg.solve
    (:470), write-back R_i^T x_i (:472-476).  Current sources are not restated.
    """
    pos = np.atleast_2d(np.asarray(pos, dtype=float))
    pol = np.atleast_2d(np.asarray(pol, dtype=float))
    n = len(pos)
    Rm = quat_to_matrix(np.broadcast_to(np.atleast_2d(quat), (n, 4)))
    pol_glob = np.einsum("pij,pj->pi", Rm, pol)
    pol_magnets = np.reshape(pol_glob, (3 * n, 1), order="F")
    sus = susceptibility_vector(chi, n)
    S = np.diag(sus)
    if H_ext is None:
        H_ext = np.zeros((n, 3))
    H_ext = np.broadcast_to(np.asarray(H_ext, dtype=float), (n, 3))
    H_ext_v = np.reshape(H_ext, (3 * n, 1), order="F")

    import time as _time

    t0 = _time.perf_counter()
    T4 = demag_tensor_ref(
        pos, quat, dim, pairs_matching=pairs_matching, split=split, max_dist=max_dist, threads=threads
    )
    T = tensor_to_matrix(T4)
    t1 = _time.perf_counter()
    Q = np.eye(3 * n) - np.matmul(S, T)
    rhs = pol_magnets + np.matmul(S, H_ext_v)
    t2 = _time.perf_counter()
    pol_new = np.linalg.solve(Q, rhs)
    t3 = _time.perf_counter()
    if timings is not None:
        timings.update(assembly_s=t1 - t0, build_Q_s=t2 - t1, solve_s=t3 - t2)
    pol_new = np.reshape(pol_new, (n, 3), order="F")
    pol_loc = np.einsum("pji,pj->pi", Rm, pol_new)
    if return_parts:
        return pol_loc, {"T": T, "Q": Q, "rhs": rhs, "x_global": pol_new}
    return pol_loc
