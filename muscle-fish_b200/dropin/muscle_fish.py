"""``import muscle_fish as mf`` drop-in: put this directory first on PYTHONPATH (the reference's
``modules/`` directory after it) and the reference's scripts (general_pipeline/fish_analysis.py:21-22
and friends) run unchanged: ``find_spots``, ``find_spots_snrfilter``, ``fix_bleaching`` (and ``threshold_fiber``
/ ``threshold_nuclei``) are the B200 implementations, every other name (``open_image`` ...) falls through to
the original module (see ``_passthrough.py``)."""
import os as _os
import sys as _sys

_here = _os.path.dirname(_os.path.abspath(__file__))
_root = _os.path.dirname(_os.path.dirname(_here))
if _root not in _sys.path:
    _sys.path.insert(0, _root)

from muscle_fish_b200.muscle_fish import find_spots, find_spots_snrfilter, fix_bleaching  # noqa: E402,F401
from muscle_fish_b200.muscle_fish import find_spots_snrfilter_with_distances  # noqa: E402,F401  (fused extra)
from muscle_fish_b200.dropin import _passthrough  # noqa: E402

HOT_PATH = ("find_spots", "find_spots_snrfilter", "fix_bleaching")

# The segmentation functions (blurs + Li / Otsu on the GPU, labelling on the host) replace the originals too
# unless MFISH_DROPIN_SEGMENTATION=0 leaves them to the fall-through below.
if _os.environ.get("MFISH_DROPIN_SEGMENTATION", "1") != "0":
    from muscle_fish_b200.segmentation import threshold_fiber, threshold_nuclei  # noqa: E402,F401
    HOT_PATH = HOT_PATH + ("threshold_fiber", "threshold_nuclei")


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return _passthrough.fallthrough("muscle_fish", _here, name)


def __dir__():
    return sorted(set(globals()) | set(_passthrough.names("muscle_fish", _here)))
