"""Host-side filter taps and small exact helpers (float64, numpy only).

The taps reproduce what ``scipy.ndimage.gaussian_filter1d`` builds
(``truncate=4.0``): radius ``int(4*sigma + 0.5)``, ``phi = exp(-x^2/(2 s^2))``
normalised by its discrete sum; the order-2 taps are
``(x^2/s^4 - 1/s^2) * phi`` with the *same* normalised ``phi`` (their sum is
deliberately NOT forced to zero -- SURVEY 2a).  They are computed here in
float64 and handed to the kernels, which hold them in constant memory.
"""
from __future__ import annotations

import math

import numpy as np

TRUNCATE = 4.0


def gaussian_radius(sigma: float, truncate: float = TRUNCATE) -> int:
    return int(truncate * float(sigma) + 0.5)


def gaussian_taps(sigma: float, order: int = 0, truncate: float = TRUNCATE):
    """Return (half_taps, radius): half_taps[k] is the weight at offsets +-k.

    Orders 0 and 2 are symmetric, so only k = 0..radius is returned.
    """
    if order not in (0, 2):
        raise ValueError("only orders 0 and 2 are on the spot path")
    sigma = float(sigma)
    r = gaussian_radius(sigma, truncate)
    if sigma <= 0.0:
        w = np.zeros(1)
        w[0] = 1.0 if order == 0 else 0.0
        return w, 0
    x = np.arange(-r, r + 1)
    s2 = sigma * sigma
    phi = np.exp(-0.5 / s2 * x ** 2)
    phi = phi / phi.sum()
    if order == 0:
        full = phi
    else:
        # q(x) = x^2/s^4 - 1/s^2  (second derivative polynomial)
        full = (x.astype(np.float64) ** 2 / (s2 * s2) - 1.0 / s2) * phi
    return np.ascontiguousarray(full[r:]), r


def sigma_list(min_sigma, max_sigma, num_sigma):
    """skimage.blob_log's linear sigma ladder (log_scale=False)."""
    scale = np.linspace(0, 1, int(num_sigma))
    return scale * (float(max_sigma) - float(min_sigma)) + float(min_sigma)


# --------------------------------------------------------------------------- #
# blob overlap (equal or unequal sigma), the closed form blob_log's pruning
# step uses: spheres of radius sigma*sqrt(3) in 3-D.
# --------------------------------------------------------------------------- #
def sphere_overlap_fraction(dist_vox: float, sigma1: float, sigma2: float, ndim: int = 3) -> float:
    if sigma1 == 0 and sigma2 == 0:
        return 0.0
    smax = max(sigma1, sigma2)
    if sigma1 > sigma2:
        r1, r2 = 1.0, sigma2 / sigma1
    else:
        r2, r1 = 1.0, sigma1 / sigma2
    d = dist_vox / (smax * math.sqrt(ndim))
    if d > r1 + r2:
        return 0.0
    if d <= abs(r1 - r2):
        return 1.0
    vol = (math.pi / (12 * d) * (r1 + r2 - d) ** 2
           * (d * d + 2 * d * (r1 + r2) - 3 * (r1 - r2) ** 2))
    return vol / (4.0 / 3 * math.pi * min(r1, r2) ** 3)


def prune_cutoff_sq(sigma1: float, sigma2: float | None = None, overlap: float = 0.5) -> int:
    """Largest integer squared voxel distance at which two blobs still overlap
    by more than ``overlap`` (-1 if none).  Coordinates are integers, so the
    kernels compare integer squared distances against this table.
    """
    if sigma2 is None:
        sigma2 = sigma1
    smax = max(sigma1, sigma2)
    lim = int((2.0 * smax * math.sqrt(3.0)) ** 2) + 2
    best = -1
    for n in range(1, lim + 1):
        if sphere_overlap_fraction(math.sqrt(n), sigma1, sigma2) > overlap:
            best = n
    return best


# --------------------------------------------------------------------------- #
# np.interp / np.percentile scalar arithmetic, restated so a pair of order
# statistics selected on the GPU can be finished exactly in float64.
# --------------------------------------------------------------------------- #
def interp01(x, lo: float, hi: float):
    """np.interp(x, (lo, hi), (0, 1)) for lo <= x <= hi (float64)."""
    x = np.asarray(x, dtype=np.float64)
    lo = float(lo)
    hi = float(hi)
    if hi == lo:
        return np.zeros_like(x) + 1.0  # np.interp returns fp[-1] for x >= xp[-1]
    slope = (1.0 - 0.0) / (hi - lo)
    out = slope * (x - lo) + 0.0
    out = np.where(x >= hi, 1.0, out)
    out = np.where(x <= lo, 0.0, out)
    return out


def percentile_virtual_index(n: int, q_percent: float):
    """numpy 'linear' method (virtual index (n - 1) * q): virtual index, floor index, gamma."""
    q = np.true_divide(q_percent, 100.0)
    vi = (n - 1) * q
    k = math.floor(vi)
    k = min(max(k, 0), n - 1)
    gamma = vi - math.floor(vi)
    return vi, int(k), float(gamma)


def lerp(a: float, b: float, t: float) -> float:
    """numpy.lib._function_base_impl._lerp for scalars."""
    d = b - a
    out = a + d * t
    if t >= 0.5:
        out = b - d * (1.0 - t)
    if d == 0:
        out = a
    return float(out)
