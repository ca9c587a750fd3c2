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


# ---------------------------------------------------------------------------------------------------------
# Model of the two-threads-per-line envelope pass (csrc/edt.cu, BANDS = 2).  Not a restatement of anything in the
# reference -- scipy's transform has no such step -- but the executable specification the CUDA kernel was written
# against: plain Python on one line, checked against the brute-force minimum in tests/test_oracle.py.
# ---------------------------------------------------------------------------------------------------------
_NONE = -1


def _hidden(Ga, a, Gb, b, Gq, q):
    """site b is hidden by (a, q), a < b < q, with G(u) = F(u) + u^2: point b on or above the segment a-q"""
    return (Gb - Ga) * (q - b) >= (Gq - Gb) * (b - a)


def _build_chain(F):
    """hull of one band as an in-place linked stack: link[q] = entry below q when q was pushed -> (top, link)"""
    link = [_NONE] * len(F)
    t0 = t1 = _NONE
    for q, f in enumerate(F):
        if f is None:
            continue
        while t1 != _NONE and _hidden(F[t1] + t1 * t1, t1, F[t0] + t0 * t0, t0, f + q * q, q):
            t0, t1 = t1, link[t1]
        link[q] = t0
        t1, t0 = t0, q
    return t0, link


def _fill_down(F, link, cur, pstart, out, to_line):
    """walk the chain downwards from ``cur`` while p runs from ``pstart`` to 0 (the band's own frame)"""
    prev = link[cur]
    for p in range(pstart, -1, -1):
        while prev != _NONE and 2 * p * (cur - prev) <= (F[cur] + cur * cur) - (F[prev] + prev * prev):
            cur, prev = prev, link[prev]
        out[to_line(p)] = F[cur] + (p - cur) ** 2


def two_band_envelope(F):
    """min over u of F[u] + (p - u)^2 for every p of one line (``None`` = no site), computed the way the banded
    kernel does: the lower band scans upwards, the upper band scans downwards in mirrored coordinates, the two
    chains are joined by their lower common tangent (entries dropped from the two tops only), and each band fills
    outwards from the tangent's crossing point.  Returns a list (``None`` where the line has no site)."""
    L = len(F)
    Lh = L // 2
    out = [None] * L
    top = L - 1
    mirror = lambda t: top - t  # noqa: E731
    Fm = F[::-1]  # the line in the upper band's frame
    t_lo, link_lo = _build_chain(F[:Lh])
    t_hi, link_hi = _build_chain(Fm[:L - Lh])
    link_lo = link_lo + [_NONE] * (L - Lh)
    link_hi = link_hi + [_NONE] * Lh
    if t_lo == _NONE and t_hi == _NONE:
        return out
    if t_lo == _NONE:
        _fill_down(Fm, link_hi, t_hi, top, out, mirror)
        return out
    if t_hi == _NONE:
        _fill_down(F, link_lo, t_lo, top, out, lambda p: p)
        return out
    G = lambda u: F[u] + u * u  # noqa: E731
    l, lb = t_lo, link_lo[t_lo]
    r = mirror(t_hi)
    rb = link_hi[t_hi]
    rb = _NONE if rb == _NONE else mirror(rb)
    while True:
        moved = False
        while lb != _NONE and _hidden(G(lb), lb, G(l), l, G(r), r):
            l, lb = lb, link_lo[lb]
            moved = True
        while rb != _NONE and _hidden(G(l), l, G(r), r, G(rb), rb):
            r = rb
            t = link_hi[mirror(r)]
            rb = _NONE if t == _NONE else mirror(t)
            moved = True
        if not moved:
            break
    m = (G(r) - G(l)) // (2 * (r - l))  # l is at least as good as r for p <= m (floor division)
    m = max(-1, min(m, top))
    if m >= 0:
        _fill_down(F, link_lo, l, m, out, lambda p: p)
    if top - (m + 1) >= 0:
        _fill_down(Fm, link_hi, mirror(r), top - (m + 1), out, mirror)
    return out
