"""Object-model provider: real magpylib when importable, the local shim otherwise.

The reference imports ``magpylib`` for its object classes (demag.py:7-11, meshing.py:6-9).
magpylib is not installed in this image; everything on the demag path is duck-typed, so the
shim in ``objects.py`` is sufficient and real magpylib objects are accepted unchanged.
"""

from __future__ import annotations

try:  # pragma: no cover - magpylib is absent in the build/GPU image
    import magpylib as _magpy
    from magpylib import Collection, Sensor
    from magpylib._src.obj_classes.class_BaseExcitations import BaseCurrent, BaseMagnet
    from magpylib.current import Circle, Polyline
    from magpylib.magnet import Cuboid

    HAVE_MAGPYLIB = True
    mu_0 = _magpy.mu_0
except ImportError:
    from .objects import MU0 as mu_0
    from .objects import BaseCurrent, BaseMagnet, Circle, Collection, Cuboid, Polyline, Sensor

    HAVE_MAGPYLIB = False

__all__ = ["HAVE_MAGPYLIB", "BaseCurrent", "BaseMagnet", "Circle", "Collection", "Cuboid", "Polyline", "Sensor", "mu_0"]
