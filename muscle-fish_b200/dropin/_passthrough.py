"""Fall-through of the drop-in modules to the ORIGINAL reference modules.

``dropin/muscle_fish.py`` and ``dropin/scope_utils3.py`` define only the hot-path functions.  Every
other name a reference script uses (``mf.open_image``, ``mf.threshold_fiber``, ``mf.threshold_nuclei``,
``su.animate_zstacks``, colormaps ... -- general_pipeline/fish_analysis.py:111-165) is resolved here:
the original file of the same name is looked up in ``$MUSCLE_FISH_MODULES`` and then on the remaining
``sys.path`` entries (the drop-in's own directory skipped), executed once under a private module name
and consulted by the drop-in's module-level ``__getattr__``.  Python itself never merges two files of
the same name: the first one on the path wins, which is why this exists.
"""
from __future__ import annotations

import importlib.util
import os
import sys

_ORIGINALS = {}


def find_original(name: str, own_dir: str):
    """Path of the reference's ``<name>.py`` or None."""
    own = os.path.realpath(own_dir)
    dirs = []
    env = os.environ.get("MUSCLE_FISH_MODULES")
    if env:
        dirs.extend(env.split(os.pathsep))
    dirs.extend(p if p else os.getcwd() for p in sys.path)
    for d in dirs:
        try:
            if os.path.realpath(d) == own:
                continue
            f = os.path.join(d, name + ".py")
            if os.path.isfile(f):
                return f
        except OSError:
            continue
    return None


def original(name: str, own_dir: str):
    """The reference's module ``name`` (loaded on first use), or None when it is not installed."""
    if name in _ORIGINALS:
        return _ORIGINALS[name]
    path = find_original(name, own_dir)
    mod = None
    if path is not None:
        spec = importlib.util.spec_from_file_location(f"_{name}_original", path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = mod
        try:
            spec.loader.exec_module(mod)
        except BaseException:
            del sys.modules[spec.name]
            raise
    _ORIGINALS[name] = mod
    return mod


def fallthrough(name: str, own_dir: str, attr: str):
    mod = original(name, own_dir)
    if mod is None:
        raise AttributeError(
            f"module {name!r} (B200 drop-in) defines only the spot-analysis hot path and the original "
            f"{name}.py was not found for {attr!r}: put the reference's modules/ directory after the drop-in "
            f"on PYTHONPATH or set MUSCLE_FISH_MODULES")
    try:
        return getattr(mod, attr)
    except AttributeError:
        raise AttributeError(f"module {name!r} has no attribute {attr!r}") from None


def names(name: str, own_dir: str):
    mod = _ORIGINALS.get(name)
    return [] if mod is None else [n for n in vars(mod) if not n.startswith("__")]
