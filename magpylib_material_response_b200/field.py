"""``getB`` / ``getH`` at observers through the CUDA field kernels -- stands in for the magpylib
calls the reference makes around the solve: ``collection.getB(observers)`` of the solved cells
(/root/reference/tests/test_isotropic_anisotropic.py:23; ``mmr_field``) and
``magpy.getB(currents_list, pos, sumup=True)`` (demag.py:456-463, SURVEY.md 8f rank 3;
``mmr_field_currents``).  Cuboid magnets, Polyline and Circle currents; no CPU fallback.
"""

from __future__ import annotations

import numpy as np

from . import _native
from ._compat import BaseCurrent, BaseMagnet, Circle, Collection, Cuboid, Polyline, Sensor


def _is_block(g):
    return getattr(g, "_block", None) is not None


def _flatten_sources(sources):
    """Leaf sources in order; a meshed collection of the local object model stays ONE entry (its
    structure-of-arrays cell block is read as arrays, no per-cell objects)."""
    out = []
    for s in sources:
        if isinstance(s, Collection):
            out.extend(s._source_groups() if hasattr(s, "_source_groups") else s.sources_all)
        else:
            out.append(s)
    return out


def _sensor_points(sensor):
    """Global pixel positions of a sensor, (path length, *pixel shape, 3)."""
    pix = np.asarray(sensor.pixel, dtype=float)
    pos = np.asarray(sensor.position, dtype=float)
    quat = np.asarray(sensor.orientation.as_quat(), dtype=float)
    if pos.ndim == 1 and quat.ndim == 1:
        return (sensor.orientation.apply(pix.reshape(-1, 3)) + pos).reshape(1, *pix.shape)
    from scipy.spatial.transform import Rotation as R

    p = max(len(np.atleast_2d(pos)), len(np.atleast_2d(quat)))
    pos = np.broadcast_to(np.atleast_2d(pos), (p, 3)) if pos.ndim == 1 else pos
    quat = np.broadcast_to(np.atleast_2d(quat), (p, 4)) if quat.ndim == 1 else quat
    flat = pix.reshape(-1, 3)
    pts = np.stack([R.from_quat(q).apply(flat) + t for q, t in zip(quat, pos, strict=True)])
    return pts.reshape(p, *pix.shape)


def _observer_array(observers):
    """(points (m,3), output shape prefix).  Sensor observers follow magpylib's axis order
    (path, sensor, *pixel shape) with length-1 path / sensor axes squeezed."""
    if isinstance(observers, Sensor):
        pts = _sensor_points(observers)  # (p, *pix, 3)
        shape = pts.shape[:-1] if pts.shape[0] > 1 else pts.shape[1:-1]
        return pts.reshape(-1, 3), shape
    if isinstance(observers, list | tuple) and observers and all(isinstance(o, Sensor) for o in observers):
        parts = [_sensor_points(o) for o in observers]
        if len({a.shape[1:] for a in parts}) != 1:
            msg = "all sensors must have the same pixel shape"
            raise ValueError(msg)
        p = max(a.shape[0] for a in parts)
        parts = [np.broadcast_to(a[-1:], (p, *a.shape[1:])) if a.shape[0] == 1 else a for a in parts]
        if len({a.shape[0] for a in parts}) != 1:
            msg = "sensor paths must have the same length"
            raise ValueError(msg)
        pts = np.stack(parts, axis=1)  # (p, k, *pix, 3)
        shape = pts.shape[:-1] if p > 1 else pts.shape[1:-1]
        return pts.reshape(-1, 3), shape
    arr = np.asarray(observers, dtype=float)
    if arr.shape[-1] != 3:
        msg = f"observers must have shape (..., 3), got {arr.shape}"
        raise ValueError(msg)
    return arr.reshape(-1, 3), arr.shape[:-1]


def cell_arrays(cells):
    """SoA view of cuboid cells: pos (n,3), quat (n,4) scalar-last, dim (n,3), pol (n,3)."""
    n = sum(len(c._block) if _is_block(c) else 1 for c in cells)
    pos = np.empty((n, 3))
    quat = np.empty((n, 4))
    dim = np.empty((n, 3))
    pol = np.zeros((n, 3))
    i = 0
    for c in cells:
        if _is_block(c):  # a meshed collection: array slices
            blk = c._block
            m = len(blk)
            pos[i:i + m], quat[i:i + m], dim[i:i + m], pol[i:i + m] = blk.pos, blk.rot.as_quat(), blk.dim, blk.pol
            i += m
            continue
        p = getattr(c, "barycenter", c.position)
        if np.ndim(p) != 1:
            msg = "objects with paths are not supported by the B200 field kernel"
            raise ValueError(msg)
        pos[i] = p
        quat[i] = c.orientation.as_quat()
        dim[i] = c.dimension
        if c.polarization is not None:
            pol[i] = c.polarization
        i += 1
    return pos, quat, dim, pol


def current_arrays(currents):
    """Current sources as the arrays ``mmr_field_currents`` takes: Polyline vertices moved to the
    global frame and split into segments (nseg,7) = (p1, p2, I); Circles (nloop,9) = (centre, quat
    scalar-last, diameter, I).  Sources without a current / vertices / diameter contribute nothing
    (magpylib evaluates them as zero)."""
    segs, loops = [], []
    for c in currents:
        if np.ndim(c.position) != 1:
            msg = "objects with paths are not supported by the B200 field kernel"
            raise ValueError(msg)
        cur = 0.0 if c.current is None else float(c.current)
        if isinstance(c, Polyline):
            if c.vertices is None:
                continue
            v = c.orientation.apply(np.asarray(c.vertices, dtype=float)) + np.asarray(c.position, dtype=float)
            segs.extend([*a, *b, cur] for a, b in zip(v[:-1], v[1:], strict=True))
        elif isinstance(c, Circle):
            if c.diameter is None:
                continue
            loops.append([*np.asarray(c.position, dtype=float), *c.orientation.as_quat(), float(c.diameter), cur])
        else:
            msg = f"the B200 current-field kernel evaluates Polyline and Circle sources; found {type(c).__name__}"
            raise NotImplementedError(msg)
    return np.array(segs, dtype=float).reshape(-1, 7), np.array(loops, dtype=float).reshape(-1, 9)


def current_field(field, currents, points, device=None, ctx=None):
    """Summed B (tesla) or H (A/m) of the current sources at ``points`` (m,3) -> (m,3) numpy."""
    seg, loop = current_arrays(currents)
    pts = np.ascontiguousarray(points, dtype=float).reshape(-1, 3)
    if len(pts) == 0 or (len(seg) == 0 and len(loop) == 0):
        return np.zeros((len(pts), 3))
    ctx = _native.default_context(device) if ctx is None else ctx
    out = np.zeros((len(pts), 3))
    d_pts = ctx.to_device(pts)
    # the kernel stages its sources in shared memory: feed long polylines in chunks
    chunk = 3000
    for s0 in range(0, max(len(seg), 1), chunk):
        sl = seg[s0:s0 + chunk]
        lp = loop if s0 == 0 else loop[:0]
        if len(sl) == 0 and len(lp) == 0:
            continue
        d_out = ctx.empty(len(pts), 3)
        ctx.field_currents(d_pts, ctx.to_device(sl) if len(sl) else None, ctx.to_device(lp) if len(lp) else None,
                           d_out, _native.FIELD_B if field == "B" else _native.FIELD_H)
        out += d_out.cpu().numpy()
    return out


def getBH(field, sources, observers, sumup=True, squeeze=True, device=None):
    """Field of the sources at the observers, summed over sources (what ``Collection.getB``
    returns).  ``field`` is "B" (tesla) or "H" (A/m)."""
    if field not in ("B", "H"):
        msg = "field must be 'B' or 'H'"
        raise ValueError(msg)
    if not sumup or not squeeze:
        # the kernel reduces over the sources on the device; per-source output is not produced
        msg = "the B200 field path returns the summed, squeezed field only (sumup=True, squeeze=True)"
        raise NotImplementedError(msg)
    srcs = _flatten_sources(sources)
    cuboids = [s for s in srcs if _is_block(s) or isinstance(s, Cuboid)]
    currents = [s for s in srcs if isinstance(s, BaseCurrent)]
    bad = [s for s in srcs if not _is_block(s) and not isinstance(s, Cuboid | BaseCurrent)]
    if bad:
        kinds = sorted({type(s).__name__ for s in bad})
        if any(isinstance(s, BaseMagnet) for s in bad):
            msg = f"the B200 field kernel evaluates Cuboid cells only; found {kinds}"
            raise NotImplementedError(msg)
        msg = f"unsupported source objects {kinds}"
        raise TypeError(msg)
    pts, shape = _observer_array(observers)
    ctx = _native.default_context(device)
    out = np.zeros((len(pts), 3))
    if cuboids:
        pos, quat, dim, pol = cell_arrays(cuboids)
        d_out = ctx.empty(len(pts), 3)
        ctx.field(
            ctx.to_device(pos), ctx.to_device(quat), ctx.to_device(dim), ctx.to_device(pol),
            ctx.to_device(pts), d_out, _native.FIELD_B if field == "B" else _native.FIELD_H,
        )
        out += d_out.cpu().numpy()
    if currents:
        out += current_field(field, currents, pts, ctx=ctx)
    return out.reshape(*shape, 3)
