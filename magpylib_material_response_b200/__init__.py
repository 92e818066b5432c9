"""B200-native demagnetization hot path of magpylib-material-response.

Drop-in for the reference's ``demag`` / ``meshing`` modules on the path named in
BASELINE.json: ``mesh_Cuboid`` -> ``apply_demag`` (tensor assembly + dense solve) -> ``getB``.
Compute runs in hand-written CUDA for sm_100a behind the C ABI in include/mmr_b200.h.
"""

from __future__ import annotations

from . import demag, meshing
from ._compat import HAVE_MAGPYLIB, BaseCurrent, BaseMagnet, Circle, Collection, Cuboid, Polyline, Sensor, mu_0
from .demag import apply_demag, demag_tensor, get_H_ext, get_susceptibilities
from .meshing import mesh_all, mesh_Cuboid, slice_Cuboid
from .serialization import from_json, to_json

__version__ = "0.1.0"

__all__ = [
    "HAVE_MAGPYLIB",
    "BaseCurrent",
    "BaseMagnet",
    "Circle",
    "Collection",
    "Cuboid",
    "Polyline",
    "Sensor",
    "__version__",
    "apply_demag",
    "demag",
    "demag_tensor",
    "from_json",
    "get_H_ext",
    "get_susceptibilities",
    "mesh_Cuboid",
    "mesh_all",
    "meshing",
    "mu_0",
    "slice_Cuboid",
    "to_json",
]
