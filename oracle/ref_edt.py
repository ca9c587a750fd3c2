"""Oracle (TEST INFRASTRUCTURE ONLY): nucleus distance transform + gather.

Follows ``nocodazole/fish_analysis_noco.py:218-221,297-298`` and
``simulation/analyze_simulation.py:101,137`` (paths relative to the upstream
repository root).  ``scipy.ndimage.distance_transform_edt`` is the real
dependency the reference calls.
"""
from __future__ import annotations

import numpy as np
from scipy import ndimage as ndi
from scipy.interpolate import interpn


def distance_transform_edt(input, sampling=None):
    """Distance (float64) from every voxel to the nearest ZERO voxel."""
    return ndi.distance_transform_edt(np.asarray(input), sampling=sampling)


def edt_gather(grassfire, spots, dims_xyz):
    """nocodazole/fish_analysis_noco.py:219-221,297-298.

    grid_? = [dims*i ...]; spots_xyz = spot[:3]*dims; interpn(...).  Queries
    fall exactly on grid nodes, so this equals ``grassfire[ix, iy, iz]``.
    """
    dims_xyz = np.asarray(dims_xyz, dtype=np.float64)
    grid = tuple([dims_xyz[a] * i for i in range(grassfire.shape[a])] for a in range(3))
    spots = np.asarray(spots, dtype=np.float64)
    spots_xyz = np.array([s[:3] * dims_xyz for s in spots if not np.isnan(s[0])])
    if len(spots_xyz) == 0:
        return np.empty((0,))
    return interpn(grid, grassfire, spots_xyz)


def edt_bruteforce_sq(input, sampling=None):
    """O(N*M) brute force squared distance to the nearest zero voxel.

    Independent of scipy's algorithm; used to pin the oracle on small
    volumes.  Returns float64 squared distances (``inf`` if no zero voxel).
    """
    a = np.asarray(input) != 0
    nd = a.ndim
    w = np.ones(nd) if sampling is None else np.asarray(sampling, dtype=np.float64)
    zeros = np.argwhere(~a).astype(np.float64) * w
    out = np.full(a.shape, np.inf)
    if len(zeros) == 0:
        return out
    pts = np.argwhere(np.ones(a.shape, dtype=bool)).astype(np.float64) * w
    chunk = max(1, 2_000_000 // max(1, len(zeros)))
    res = np.empty(len(pts))
    for s in range(0, len(pts), chunk):
        d = pts[s:s + chunk, None, :] - zeros[None, :, :]
        res[s:s + chunk] = np.min(np.sum(d * d, axis=-1), axis=1)
    return res.reshape(a.shape)
