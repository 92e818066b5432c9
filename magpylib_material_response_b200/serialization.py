"""JSON persistence of (solved) collections -- the reference's answer to "do not recompute the
demagnetization" (/root/reference/src/magpylib_material_response/utils.py:110-271, document format in
docs/serialization.md).

The document format, type tags, units, error messages and warnings are the reference's; the
implementation is table-driven: every type tag maps to its class and to the list of
``(json key, attribute, unit)`` fields it carries besides the common pose/style block.  The classes
are whatever ``_compat`` provides (real magpylib objects when importable, the local stand-ins
otherwise); the Cylinder / CylinderSegment tags exist only when magpylib supplies those classes.
"""

from __future__ import annotations

import json
import warnings
from typing import NamedTuple

import numpy as np
from scipy.spatial.transform import Rotation

from . import _compat
from ._compat import BaseCurrent, BaseMagnet, Circle, Collection, Cuboid, Polyline, Sensor


class _Field(NamedTuple):
    key: str        # key in the JSON object and constructor keyword
    label: str      # name used in the unit error message
    unit: str
    scalar: bool = False
    optional: bool = False


class _Kind(NamedTuple):
    cls: type
    fields: tuple


_POLARIZATION = _Field("polarization", "Polarization", "T")
_CURRENT = _Field("current", "Current", "A", scalar=True)

# type tag -> class and type-specific fields (utils.py:244-256 for the tags)
_KINDS = {
    "magnet.Cuboid": _Kind(Cuboid, (_POLARIZATION, _Field("dimension", "Dimension", "m"))),
    "current.Polyline": _Kind(Polyline, (_CURRENT, _Field("vertices", "Vertices", "m"))),
    "current.Circle": _Kind(Circle, (_CURRENT, _Field("diameter", "Diameter", "m", scalar=True))),
    "Sensor": _Kind(Sensor, (_Field("pixel", "Pixel", "m", optional=True),)),
    "Collection": _Kind(Collection, ()),
}
if _compat.HAVE_MAGPYLIB:  # pragma: no cover - magpylib is absent in the build/GPU image
    import magpylib as _magpy

    _KINDS["magnet.Cylinder"] = _Kind(_magpy.magnet.Cylinder, (_POLARIZATION, _Field("dimension", "Dimension", "m")))
    _KINDS["magnet.CylinderSegment"] = _Kind(
        _magpy.magnet.CylinderSegment, (_POLARIZATION, _Field("dimension", "Dimension", "m,m,m,deg,deg")))
_TAG_OF = {kind.cls: tag for tag, kind in _KINDS.items()}
# the reference's private names, kept for callers that import them
_CLASS_BY_TYPE = {tag: kind.cls for tag, kind in _KINDS.items()}
_TYPE_BY_CLASS = _TAG_OF


def _plain(value, scalar):
    """numpy / python value -> JSON primitive(s); missing excitations serialize as zero."""
    if scalar:
        return float(0.0 if value is None else value)
    return np.asarray((0.0, 0.0, 0.0) if value is None else value, dtype=float).tolist()


def _serialize_recursive(obj, parent="warn"):
    """Object -> JSON-compatible dict with a namespaced ``type`` tag (utils.py:110-173)."""
    tag = _TAG_OF.get(type(obj))
    if tag is None and type(obj).__name__ == "CellView":  # a row of a meshed collection's cell block
        tag = "magnet.Cuboid"
    if tag is None:
        msg = (
            f"Unsupported object type: {obj.__class__.__name__!r}. "
            f"Supported types: {sorted(_KINDS)}"
        )
        raise TypeError(msg)
    if parent == "warn" and obj.parent is not None:
        warnings.warn(f"object parent ({obj.parent}) not included in serialization", stacklevel=2)
    doc = {
        "type": tag,
        "position": {"value": np.asarray(obj.position).tolist(), "unit": "m"},
        "orientation": {"value": obj.orientation.as_matrix().tolist(), "representation": "matrix"},
    }
    style = obj.style.as_dict()
    if style:
        doc["style"] = style
    for fld in _KINDS[tag].fields:
        value = getattr(obj, fld.key)
        if fld.optional and value is None:
            continue
        doc[fld.key] = {"value": _plain(value, fld.scalar), "unit": fld.unit}
    if isinstance(obj, BaseMagnet):
        chi = getattr(obj, "susceptibility", None)
        if chi is not None:
            chi = np.asarray(chi)
            doc["susceptibility"] = {"value": chi.tolist() if chi.ndim else float(chi)}
    if tag == "Collection":
        doc["children"] = [_serialize_recursive(child, parent="ignore") for child in obj.children]
    return doc


def _check_unit(field_name, inp, expected):
    got = inp.get("unit")
    if got != expected:
        msg = f"{field_name} unit must be {expected!r}, got {got!r}"
        raise ValueError(msg)


def _deserialize_recursive(inp):
    """Inverse of :func:`_serialize_recursive` (utils.py:183-241)."""
    tag = inp.get("type")
    kind = _KINDS.get(tag)
    if kind is None:
        msg = f"Unknown type tag {tag!r}. Supported tags: {sorted(_KINDS)}"
        raise TypeError(msg)
    _check_unit("Position", inp["position"], "m")
    representation = inp["orientation"].get("representation")
    if representation != "matrix":
        msg = f"Orientation representation must be 'matrix', got {representation!r}"
        raise ValueError(msg)
    kwargs = {
        "position": inp["position"]["value"],
        "orientation": Rotation.from_matrix(inp["orientation"]["value"]),
    }
    if inp.get("style") is not None:
        kwargs["style"] = inp["style"]
    if inp.get("parent") is not None:
        warnings.warn(f"object parent ({inp['parent']}) ignored", stacklevel=2)
    for fld in kind.fields:
        if fld.optional and fld.key not in inp:
            continue
        _check_unit(fld.label, inp[fld.key], fld.unit)
        kwargs[fld.key] = inp[fld.key]["value"]
    obj = kind.cls(**kwargs)
    if inp.get("susceptibility") is not None:
        chi = inp["susceptibility"]["value"]
        obj.susceptibility = tuple(chi) if isinstance(chi, list) else chi
    if tag == "Collection":
        obj.add(*[_deserialize_recursive(child) for child in inp["children"]])
    return obj


def to_json(*objs, **json_kwargs):
    """Serialize one or more objects to a JSON string (``json_kwargs`` go to ``json.dumps``)."""
    return json.dumps([_serialize_recursive(obj) for obj in objs], **json_kwargs)


def from_json(s):
    """Deserialize a JSON string produced by :func:`to_json`; returns a list of objects."""
    return [_deserialize_recursive(obj) for obj in json.loads(s)]
