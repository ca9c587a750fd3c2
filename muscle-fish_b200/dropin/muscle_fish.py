"""``import muscle_fish as mf`` drop-in: put this directory first on PYTHONPATH and the
reference's scripts (general_pipeline/fish_analysis.py:21-22 and friends) pick up the
B200 spot path unchanged."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)

from muscle_fish_b200.muscle_fish import find_spots, find_spots_snrfilter, fix_bleaching  # noqa: E402,F401
