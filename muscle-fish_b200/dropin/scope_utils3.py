"""``import scope_utils3 as su`` drop-in (see ``dropin/muscle_fish.py``): ``normalize_image``,
``optimize_threshold_blob_log`` and the ``gauss_*`` formulas are defined here, every other name
(``animate_zstacks``, colormaps, CZI metadata helpers ...) falls through to the original module."""
import os as _os
import sys as _sys

_here = _os.path.dirname(_os.path.abspath(__file__))
_root = _os.path.dirname(_os.path.dirname(_here))
if _root not in _sys.path:
    _sys.path.insert(0, _root)

from muscle_fish_b200.scope_utils3 import (  # noqa: E402,F401
    normalize_image, optimize_threshold_blob_log, gauss_1d, gauss_2d, gauss_3d,
)
from muscle_fish_b200.dropin import _passthrough  # noqa: E402

HOT_PATH = ("normalize_image", "optimize_threshold_blob_log", "gauss_1d", "gauss_2d", "gauss_3d")


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return _passthrough.fallthrough("scope_utils3", _here, name)


def __dir__():
    return sorted(set(globals()) | set(_passthrough.names("scope_utils3", _here)))
