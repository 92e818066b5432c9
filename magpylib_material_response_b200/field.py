"""``getB`` / ``getH`` of cuboid cells at observers through the CUDA field kernel
(``mmr_field``) -- stands in for the magpylib call the reference's tests make after the
solve (/root/reference/tests/test_isotropic_anisotropic.py:23).
Cuboid magnets only; no CPU fallback.
"""

from __future__ import annotations

import numpy as np

from . import _native
from ._compat import BaseCurrent, BaseMagnet, Collection, Cuboid, Sensor


def _flatten_sources(sources):
    out = []
    for s in sources:
        if isinstance(s, Collection):
            out.extend(s.sources_all)
        else:
            out.append(s)
    return out


def _observer_array(observers):
    """(points (m,3), output shape prefix)."""
    if isinstance(observers, Sensor):
        pts = observers.global_pixels()
        return pts, np.asarray(observers.pixel).shape[:-1]
    if isinstance(observers, list | tuple) and observers and all(isinstance(o, Sensor) for o in observers):
        parts = [o.global_pixels() for o in observers]
        shp = np.asarray(observers[0].pixel).shape[:-1]
        return np.concatenate(parts, axis=0), (len(observers), *shp)
    arr = np.asarray(observers, dtype=float)
    if arr.shape[-1] != 3:
        msg = f"observers must have shape (..., 3), got {arr.shape}"
        raise ValueError(msg)
    return arr.reshape(-1, 3), arr.shape[:-1]


def cell_arrays(cells):
    """SoA view of cuboid cells: pos (n,3), quat (n,4) scalar-last, dim (n,3), pol (n,3)."""
    n = len(cells)
    pos = np.empty((n, 3))
    quat = np.empty((n, 4))
    dim = np.empty((n, 3))
    pol = np.zeros((n, 3))
    for i, c in enumerate(cells):
        p = getattr(c, "barycenter", c.position)
        if np.ndim(p) != 1:
            msg = "objects with paths are not supported by the B200 field kernel"
            raise ValueError(msg)
        pos[i] = p
        quat[i] = c.orientation.as_quat()
        dim[i] = c.dimension
        if c.polarization is not None:
            pol[i] = c.polarization
    return pos, quat, dim, pol


def getBH(field, sources, observers, sumup=True, squeeze=True, device=None):
    """Field of cuboid sources at the observers, summed over sources (what
    ``Collection.getB`` returns).  ``field`` is "B" (tesla) or "H" (A/m)."""
    if field not in ("B", "H"):
        msg = "field must be 'B' or 'H'"
        raise ValueError(msg)
    srcs = _flatten_sources(sources)
    bad = [s for s in srcs if not isinstance(s, Cuboid)]
    if bad:
        kinds = sorted({type(s).__name__ for s in bad})
        if any(isinstance(s, BaseCurrent) for s in bad) or any(isinstance(s, BaseMagnet) for s in bad):
            msg = f"the B200 field kernel evaluates Cuboid cells only; found {kinds}"
            raise NotImplementedError(msg)
        msg = f"unsupported source objects {kinds}"
        raise TypeError(msg)
    pts, shape = _observer_array(observers)
    ctx = _native.default_context(device)
    pos, quat, dim, pol = cell_arrays(srcs)
    d_out = ctx.empty(len(pts), 3)
    ctx.field(
        ctx.to_device(pos), ctx.to_device(quat), ctx.to_device(dim), ctx.to_device(pol),
        ctx.to_device(pts), d_out, _native.FIELD_B if field == "B" else _native.FIELD_H,
    )
    out = d_out.cpu().numpy().reshape(*shape, 3)
    return out
