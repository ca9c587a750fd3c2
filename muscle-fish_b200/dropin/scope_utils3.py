"""``import scope_utils3 as su`` drop-in (see ``dropin/muscle_fish.py``)."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)

from muscle_fish_b200.scope_utils3 import (  # noqa: E402,F401
    normalize_image, optimize_threshold_blob_log, gauss_1d, gauss_2d, gauss_3d,
)
