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

from ._compat import HAVE_MAGPYLIB, BaseCurrent, Collection, Cuboid
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
    """meshing.py:19-29: cells inherit susceptibility, collection takes the parent's pose.  ``cells``
    is a list of objects or a ``CellBlock`` (structure-of-arrays cells of the local object model)."""
    susceptibility = getattr(obj, "susceptibility", None)
    if isinstance(cells, list):
        if susceptibility is not None:
            for cell in cells:
                cell.susceptibility = susceptibility
        coll = Collection(cells)
    else:
        if susceptibility is not None:
            cells.attrs["susceptibility"] = susceptibility
        coll = Collection._from_block(cells)
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
    if HAVE_MAGPYLIB:  # pragma: no cover - real magpylib objects: one Cuboid per cell, as the reference
        cells = [Cuboid(polarization=pol0, dimension=new_dim, position=pp) for pp in grid]
    else:
        from .objects import CellBlock

        cells = CellBlock(grid, new_dim, pol=pol0)  # SoA cells; objects are views made on demand
    return _collection_from_obj_and_cells(cuboid, cells, **kwargs)


def get_volume(obj, return_containing_cube_edge=False):
    """Volume of a Cuboid and, on request, the edge of the smallest origin-centred axis-aligned cube
    that contains it (meshing_utils.py:113-137; the B200 path has the Cuboid branch only)."""
    if not isinstance(obj, Cuboid):
        msg = "Unsupported object type for volume calculation"
        raise TypeError(msg)
    edges = np.asarray(obj.dimension, dtype=float)
    volume = float(np.prod(edges))
    return (volume, float(edges.max())) if return_containing_cube_edge else volume


def slice_Cuboid(cuboid, shift=0.5, axis="z", **kwargs):
    """Cut a Cuboid in two with a plane normal to ``axis`` at the relative position ``shift`` in (0, 1),
    measured from the low side (meshing.py:247-304).  Returns a Collection [low part, high part]."""
    if not isinstance(cuboid, Cuboid):
        msg = (
            "Object to be sliced must be a Cuboid, "
            f"received instead {cuboid.__class__.__name__!r}"
        )
        raise TypeError(msg)
    if not 0 < shift < 1:
        msg = "Shift must be between 0 and 1 (exclusive)"
        raise ValueError(msg)
    k = "xyz".index(axis)
    full = np.asarray(cuboid.dimension, dtype=float)
    lo_edge = -full[k] / 2
    parts = []
    for frac0, frac1 in ((0.0, shift), (shift, 1.0)):  # low part first
        size = full.copy()
        size[k] = full[k] * (frac1 - frac0)
        centre = np.zeros(3)
        centre[k] = lo_edge + full[k] * (frac0 + frac1) / 2
        parts.append(Cuboid(polarization=cuboid.polarization, dimension=size, position=centre))
    return _collection_from_obj_and_cells(cuboid, parts, **kwargs)


def _share_elements(cuboids, target_elems, min_elems, per_child_elems):
    """Cells per cuboid: the same number for all, or ``target_elems`` split by volume fraction
    (truncated), never below ``min_elems`` (meshing.py:418-424)."""
    if per_child_elems:
        return [max(min_elems, target_elems)] * len(cuboids)
    vol = np.array([get_volume(c) for c in cuboids], dtype=float)
    return [max(min_elems, int(v / vol.sum() * target_elems)) for v in vol]


def mesh_all(obj, target_elems, min_elems=8, per_child_elems=False, inplace=False, **kwargs):
    """Mesh every Cuboid below ``obj`` (a single object or a Collection tree) and put the meshed
    Collection in its place (meshing.py:375-454).  ``target_elems`` is shared among the cuboids by
    volume unless ``per_child_elems``; current sources stay as they are; anything else raises the
    reference's TypeError (the B200 path supports Cuboid magnets)."""
    meshable = (Cuboid,)
    root = obj if inplace else obj.copy(**kwargs)
    leaves = list(root.sources_all) if isinstance(root, Collection) else [root]
    cuboids = [c for c in leaves if isinstance(c, meshable)]
    quota = _share_elements(cuboids, target_elems, min_elems, per_child_elems) if cuboids else []
    rejected = Counter(type(c).__name__ for c in leaves if not isinstance(c, (*meshable, BaseCurrent)))
    if rejected:
        listing = ", ".join(f"{cnt} {name}" for name, cnt in rejected.items())
        msg = f"Incompatible objects found: {listing}\nSupported: {[c.__name__ for c in meshable]}."
        raise TypeError(msg)
    for cub, n_cells in zip(cuboids, quota, strict=True):
        owner = cub.parent
        meshed = mesh_Cuboid(cub, n_cells, **({} if owner is not None else kwargs))
        cub.parent = None  # detach the original; its meshed stand-in takes its place
        if owner is None:
            root = meshed
        else:
            owner.add(meshed)
    return root
