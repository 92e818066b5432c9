"""JSON persistence of (solved) collections -- the reference's answer to "do not recompute the
demagnetization" (/root/reference/src/magpylib_material_response/utils.py:110-271, format in
docs/serialization.md).  Same document format, same type tags, same
errors and warnings; the classes are whatever ``_compat`` provides (real magpylib objects when
importable, the local stand-ins otherwise).  Cylinder / CylinderSegment tags are accepted only
when magpylib supplies those classes.
"""

from __future__ import annotations

import json
import warnings

import numpy as np
from scipy.spatial.transform import Rotation

from . import _compat
from ._compat import BaseCurrent, BaseMagnet, Circle, Collection, Cuboid, Polyline, Sensor

# Type-tag <-> class mapping (utils.py:244-256).
_CLASS_BY_TYPE = {
    "magnet.Cuboid": Cuboid,
    "current.Polyline": Polyline,
    "current.Circle": Circle,
    "Sensor": Sensor,
    "Collection": Collection,
}
if _compat.HAVE_MAGPYLIB:  # pragma: no cover - magpylib is absent in the build/GPU image
    import magpylib as _magpy

    _CLASS_BY_TYPE["magnet.Cylinder"] = _magpy.magnet.Cylinder
    _CLASS_BY_TYPE["magnet.CylinderSegment"] = _magpy.magnet.CylinderSegment
_TYPE_BY_CLASS = {cls: tag for tag, cls in _CLASS_BY_TYPE.items()}


def _rotation_matrix(obj):
    return obj.orientation.as_matrix().tolist()


def _serialize_recursive(obj, parent="warn"):
    """Object -> JSON-compatible dict (utils.py:110-173)."""
    typ = _TYPE_BY_CLASS.get(type(obj))
    if typ is None:
        msg = (
            f"Unsupported object type: {obj.__class__.__name__!r}. "
            f"Supported types: {sorted(_CLASS_BY_TYPE)}"
        )
        raise TypeError(msg)
    dd = {
        "type": typ,
        "position": {"value": np.asarray(obj.position).tolist(), "unit": "m"},
        "orientation": {"value": _rotation_matrix(obj), "representation": "matrix"},
    }
    style = obj.style.as_dict()
    if style:
        dd["style"] = style
    if parent == "warn" and obj.parent is not None:
        warnings.warn(f"object parent ({obj.parent}) not included in serialization", stacklevel=2)
    if isinstance(obj, BaseMagnet):
        pol = (0.0, 0.0, 0.0) if obj.polarization is None else obj.polarization
        dd["polarization"] = {"value": np.asarray(pol, dtype=float).tolist(), "unit": "T"}
        susceptibility = getattr(obj, "susceptibility", None)
        if susceptibility is not None:
            sus = np.asarray(susceptibility)
            dd["susceptibility"] = {"value": sus.tolist() if sus.ndim else float(sus)}
    if isinstance(obj, BaseCurrent):
        dd["current"] = {"value": float(0.0 if obj.current is None else obj.current), "unit": "A"}
    if typ in ("magnet.Cuboid", "magnet.Cylinder"):
        dd["dimension"] = {"value": np.asarray(obj.dimension).tolist(), "unit": "m"}
    elif typ == "magnet.CylinderSegment":
        dd["dimension"] = {"value": np.asarray(obj.dimension).tolist(), "unit": "m,m,m,deg,deg"}
    elif typ == "current.Polyline":
        dd["vertices"] = {"value": np.asarray(obj.vertices).tolist(), "unit": "m"}
    elif typ == "current.Circle":
        dd["diameter"] = {"value": float(obj.diameter), "unit": "m"}
    elif typ == "Sensor":
        if obj.pixel is not None:
            dd["pixel"] = {"value": np.asarray(obj.pixel).tolist(), "unit": "m"}
    elif typ == "Collection":
        dd["children"] = [_serialize_recursive(child, parent="ignore") for child in obj.children]
    return dd


def _check_unit(field_name, inp, expected):
    got = inp.get("unit")
    if got != expected:
        msg = f"{field_name} unit must be {expected!r}, got {got!r}"
        raise ValueError(msg)


def _deserialize_recursive(inp):
    """Inverse of :func:`_serialize_recursive` (utils.py:183-241)."""
    typ = inp.get("type")
    constr = _CLASS_BY_TYPE.get(typ)
    if constr is None:
        msg = f"Unknown type tag {typ!r}. Supported tags: {sorted(_CLASS_BY_TYPE)}"
        raise TypeError(msg)
    kw = {}
    _check_unit("Position", inp["position"], "m")
    kw["position"] = inp["position"]["value"]
    rep = inp["orientation"].get("representation")
    if rep != "matrix":
        msg = f"Orientation representation must be 'matrix', got {rep!r}"
        raise ValueError(msg)
    kw["orientation"] = Rotation.from_matrix(inp["orientation"]["value"])
    style = inp.get("style")
    if style is not None:
        kw["style"] = style
    if inp.get("parent") is not None:
        warnings.warn(f"object parent ({inp['parent']}) ignored", stacklevel=2)
    if issubclass(constr, BaseMagnet):
        _check_unit("Polarization", inp["polarization"], "T")
        kw["polarization"] = inp["polarization"]["value"]
    if issubclass(constr, BaseCurrent):
        _check_unit("Current", inp["current"], "A")
        kw["current"] = inp["current"]["value"]
    if typ in ("magnet.Cuboid", "magnet.Cylinder"):
        _check_unit("Dimension", inp["dimension"], "m")
        kw["dimension"] = inp["dimension"]["value"]
    elif typ == "magnet.CylinderSegment":
        _check_unit("Dimension", inp["dimension"], "m,m,m,deg,deg")
        kw["dimension"] = inp["dimension"]["value"]
    elif typ == "current.Polyline":
        _check_unit("Vertices", inp["vertices"], "m")
        kw["vertices"] = inp["vertices"]["value"]
    elif typ == "current.Circle":
        _check_unit("Diameter", inp["diameter"], "m")
        kw["diameter"] = inp["diameter"]["value"]
    elif typ == "Sensor" and "pixel" in inp:
        _check_unit("Pixel", inp["pixel"], "m")
        kw["pixel"] = inp["pixel"]["value"]
    obj = constr(**kw)
    if inp.get("susceptibility") is not None:
        sus = inp["susceptibility"]["value"]
        obj.susceptibility = tuple(sus) if isinstance(sus, list) else sus
    if typ == "Collection":
        obj.add(*[_deserialize_recursive(child) for child in inp["children"]])
    return obj


def to_json(*objs, **json_kwargs):
    """Serialize one or more objects to a JSON string (``json_kwargs`` go to ``json.dumps``)."""
    return json.dumps([_serialize_recursive(obj) for obj in objs], **json_kwargs)


def from_json(s):
    """Deserialize a JSON string produced by :func:`to_json`; returns a list of objects."""
    return [_deserialize_recursive(obj) for obj in json.loads(s)]
