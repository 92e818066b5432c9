"""Meshing front end of the demag path, API and cell order unchanged from the reference:
``mesh_Cuboid`` (/root/reference/src/magpylib_material_response/meshing.py:32-95; helpers
meshing_utils.py:8-110), ``slice_Cuboid`` (meshing.py:247-304) and ``mesh_all`` with its volume
apportioning (meshing.py:375-454, meshing_utils.py:113-137).  The cell order (x outer, z fastest,
meshing.py:87) defines the row order of the demag tensor.  The cylinder meshers and ``voxelize``
produce cells the B200 kernels do not evaluate (Cylinder / CylinderSegment closed forms) and are
out of scope.
"""

from __future__ import annotations

from collections import Counter
from itertools import product

import numpy as np

from ._compat import BaseCurrent, Collection, Cuboid
from .utils import logger


def apportion_triple(triple, min_val=1, max_iter=30):
    """Rebalance a triple so every entry is >= min_val with the product kept
    (meshing_utils.py:8-23)."""
    vals = np.abs(np.array(triple, dtype=float))
    for _ in range(max_iter):
        if not np.any(vals < min_val):
            break
        lo, hi = vals.argmin(), vals.argmax()
        factor = min_val / vals[lo]
        if vals[hi] >= factor * min_val:
            vals /= factor**0.5
            vals[lo] *= factor**1.5
    return vals


def cells_from_dimension(dim, target_elems, min_val=1, strict_max=False, parity=None):
    """Number of divisions per axis closest to ``target_elems`` cubes-like cells
    (meshing_utils.py:26-110)."""
    elems = np.prod(target_elems)
    if parity == "odd":
        rounders = [
            (lambda v, add=add, fn=fn: int(2 * fn(v / 2) + add))
            for add in (-1, 1)
            for fn in (np.ceil, np.floor)
        ]
    elif parity == "even":
        rounders = [(lambda v, fn=fn: int(2 * fn(v / 2))) for fn in (np.ceil, np.floor)]
    else:
        rounders = [np.ceil, np.floor]
    elems = max(min_val**3, elems)
    x, y, z = np.abs(dim)
    est = (
        x ** (2 / 3) * (elems / y / z) ** (1 / 3),
        y ** (2 / 3) * (elems / x / z) ** (1 / 3),
        z ** (2 / 3) * (elems / x / y) ** (1 / 3),
    )
    est = apportion_triple(est, min_val=min_val)
    best_err = elems
    best = [rounders[0](k) for k in est]
    for combo in product(rounders, repeat=3):
        cand = [f(k) for f, k in zip(combo, est, strict=True)]
        err = elems - np.prod(cand)
        if abs(err) <= best_err and min(cand) >= min_val and (not strict_max or err >= 0):
            best_err = abs(err)
            best = cand
    return np.array(best).astype(int)


def _collection_from_obj_and_cells(obj, cells, **style_kwargs):
    """meshing.py:19-29: cells inherit susceptibility, collection takes the parent's pose."""
    susceptibility = getattr(obj, "susceptibility", None)
    if susceptibility is not None:
        for cell in cells:
            cell.susceptibility = susceptibility
    coll = Collection(cells)
    coll.style.update(obj.style.as_dict(), _match_properties=False)
    coll.style.update(coll._process_style_kwargs(**style_kwargs))
    coll.position = obj.position
    coll.orientation = obj.orientation
    return coll


def mesh_Cuboid(cuboid, target_elems, verbose=False, **kwargs):
    """Split a Cuboid into a regular grid of equal cuboid cells (meshing.py:32-95).

    target_elems: int (apportioned to near-cubic cells) or an (nx, ny, nz) triple.
    Returns a Collection of Cuboid cells in x-outer / z-fastest order.
    """
    if not isinstance(cuboid, Cuboid):
        msg = (
            "Object to be meshed must be a Cuboid, "
            f"received instead {cuboid.__class__.__name__!r}"
        )
        raise TypeError(msg)
    dim0 = cuboid.dimension
    pol0 = cuboid.polarization
    nnn = cells_from_dimension(dim0, target_elems) if np.isscalar(target_elems) else target_elems
    if verbose and logger is not None:
        logger.info(
            "Meshing Cuboid",
            dimensions=f"{nnn[0]}x{nnn[1]}x{nnn[2]}",
            elements=int(np.prod(nnn)),
            target=target_elems,
        )
    nnn = np.array(nnn, dtype=int)
    new_dim = dim0 / nnn
    axes = [
        np.linspace(d / 2 * (1 / n - 1), d / 2 * (1 - 1 / n), n)
        for d, n in zip(dim0, nnn, strict=True)
    ]
    grid = np.stack(np.meshgrid(*axes, indexing="ij"), axis=-1).reshape(-1, 3)
    cells = [Cuboid(polarization=pol0, dimension=new_dim, position=pp) for pp in grid]
    return _collection_from_obj_and_cells(cuboid, cells, **kwargs)


def get_volume(obj, return_containing_cube_edge=False):
    """Object volume and the edge of the smallest origin-centred unrotated cube containing it
    (meshing_utils.py:113-137; only the Cuboid branch exists on the B200 path)."""
    if obj.__class__.__name__ != "Cuboid":
        msg = "Unsupported object type for volume calculation"
        raise TypeError(msg)
    dim = obj.dimension
    vol = dim[0] * dim[1] * dim[2]
    if return_containing_cube_edge:
        return vol, max(dim)
    return vol


def slice_Cuboid(cuboid, shift=0.5, axis="z", **kwargs):
    """Slice a Cuboid in two along ``axis`` at the relative position ``shift`` in (0, 1)
    (meshing.py:247-304); returns a Collection of the two parts."""
    if not isinstance(cuboid, Cuboid):
        msg = (
            "Object to be sliced must be a Cuboid, "
            f"received instead {cuboid.__class__.__name__!r}"
        )
        raise TypeError(msg)
    if not 0 < shift < 1:
        msg = "Shift must be between 0 and 1 (exclusive)"
        raise ValueError(msg)
    dim0 = cuboid.dimension
    pol0 = cuboid.polarization
    ind = "xyz".index(axis)
    dim_k = dim0[ind]
    dims_k = dim_k * (1 - shift), dim_k * shift
    shift_k = (dim_k - dims_k[0]) / 2, -(dim_k - dims_k[1]) / 2
    cells = []
    for d, sh in zip(dims_k, shift_k, strict=True):
        dimension = np.array(dim0, dtype=float)
        dimension[ind] = d
        position = np.zeros(3)
        position[ind] = sh
        cells.append(Cuboid(polarization=pol0, dimension=dimension, position=position))
    return _collection_from_obj_and_cells(cuboid, cells[::-1], **kwargs)


def mesh_all(obj, target_elems, min_elems=8, per_child_elems=False, inplace=False, **kwargs):
    """Replace every Cuboid of ``obj`` (an object or a Collection tree) by its meshed Collection;
    ``target_elems`` is shared among the children in proportion to their volumes unless
    ``per_child_elems`` (meshing.py:375-454).  Current sources pass through unchanged; any other
    object raises the reference's TypeError (supported here: Cuboid)."""
    supported = (Cuboid,)
    allowed = (*supported, BaseCurrent)
    if not inplace:
        obj = obj.copy(**kwargs)
    children = [obj]
    if isinstance(obj, Collection):
        children = list(obj.sources_all)
    incompatible_objs = [c for c in children if not isinstance(c, allowed)]
    supported_objs = [c for c in children if isinstance(c, supported)]
    if not per_child_elems:
        volumes = np.array([get_volume(c) for c in supported_objs], dtype=float)
        volumes /= volumes.sum()
        target_elems_by_child = (volumes * target_elems).astype(int)
        target_elems_by_child[target_elems_by_child < min_elems] = min_elems
    else:
        target_elems_by_child = [max(min_elems, target_elems)] * len(supported_objs)
    if incompatible_objs:
        counts_incompatible = Counter(s.__class__.__name__ for s in incompatible_objs)
        counts_str = ", ".join(f"{count} {name}" for name, count in counts_incompatible.items())
        msg = (
            "Incompatible objects found: "
            f"{counts_str}"
            f"\nSupported: {[s.__name__ for s in supported]}."
        )
        raise TypeError(msg)
    for child, targ_elems in zip(supported_objs, target_elems_by_child, strict=True):
        parent = child.parent
        kw = kwargs if parent is None else {}
        child_meshed = mesh_Cuboid(child, int(targ_elems), **kw)
        child.parent = None
        if parent is not None:
            parent.add(child_meshed)
        else:
            obj = child_meshed
    return obj
